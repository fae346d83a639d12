/* pbwt_gpu_shim.c -- host side of the drop-in: the reference's own entry points for the hot path,
 * re-implemented in C on top of the C ABI in include/pbwtgpu.h.
 *
 * Compiled against the reference's pbwt.h (via -I$PBWT_REF, nothing is copied) and linked with the
 * reference's unmodified objects.  The reference's pbwtMatch.c is compiled with
 *   -DmatchMaximalWithin=ref_matchMaximalWithin -DmatchSequencesSweep=ref_matchSequencesSweep
 *   -DpbwtLongMatches=ref_pbwtLongMatches -DmatchSequencesDynamic=ref_matchSequencesDynamic
 * so that every caller in the program (pbwtMain.c:404-413, pbwtPaint.c:87,237, pbwtImpute.c:1142-1146)
 * reaches the functions below instead, with unchanged signatures and callback contract
 * (pbwt.h:208-213: report(ai, bi, start, end), synchronous, in the reference's order).
 * Errors follow the reference's convention: die() (utils.c:31-44).  There is no CPU fallback for the
 * GPU-backed paths; options this shim does not accelerate (-check, -stats) are passed to
 * the reference's renamed originals and say so on the log.
 */
#include "pbwt.h"
#include "pbwtgpu.h"

/* the reference's originals, renamed at compile time (see above) */
void ref_matchMaximalWithin (PBWT *p, void (*report)(int, int, int, int)) ;
void ref_matchSequencesSweep (PBWT *p, PBWT *q, void (*report)(int, int, int, int)) ;
void ref_pbwtLongMatches (PBWT *p, int L) ;
void ref_matchSequencesDynamic (PBWT *p, FILE *fp) ;

static pbwtgpu_ctx *gpu = 0 ;

static pbwtgpu_ctx *gpuContext (void)
{
  if (!gpu)
    { int rc = pbwtgpu_init (0, &gpu) ;
      if (rc) die ("pbwtgpu_init failed: %s", pbwtgpu_strerror (rc)) ;
    }
  return gpu ;
}

static void gpuCheck (int rc, const char *what)
{ if (rc) die ("%s failed: %s: %s", what, pbwtgpu_strerror (rc), pbwtgpu_last_error (gpu)) ; }

/* ---- matchMaximalWithin (pbwtMatch.c:115-142) ---- */

void matchMaximalWithin (PBWT *p, void (*report)(int ai, int bi, int start, int end))
{
  pbwtgpu_match *rec = 0 ;
  int64_t n = 0, i ;

  if (!p || !p->yz) die ("matchMaximalWithin called without a PBWT") ;
  gpuCheck (pbwtgpu_max_within (gpuContext (), arrp(p->yz,0,uchar), arrayMax(p->yz), p->M, p->N,
				p->aFstart, 0, 0, &rec, &n), "pbwtgpu_max_within") ;
  for (i = 0 ; i < n ; ++i) (*report) (rec[i].ai, rec[i].bi, rec[i].start, rec[i].end) ;
  pbwtgpu_free (rec) ;
}

/* ---- reportMatch (pbwtMatch.c:46-58): the CLI printer; zero-length matches are dropped ---- */

static void reportMatchLine (int ai, int bi, int start, int end)
{
  if (start == end) return ;
  printf ("MATCH\t%d\t%d\t%d\t%d\t%d\n", ai, bi, start, end, end-start) ;
}

/* ---- pbwtLongMatches (pbwtMatch.c:148-183): -maxWithin is L == 0 ---- */

void pbwtLongMatches (PBWT *p, int L)
{
  if (!p || !p->yz) die ("option -longWithin called without a PBWT") ;
  if (L < 0) die ("L %d for longWithin must be >= 0", L) ;
  if (isCheck || isStats)
    { fprintf (logFile, "pbwt_gpu: -check / -stats run the reference CPU path\n") ;
      ref_pbwtLongMatches (p, L) ;
      return ;
    }
  if (L)			/* matchLongWithin2 (pbwtMatch.c:85-113) */
    { pbwtgpu_match *rec = 0 ;
      int64_t n = 0, i ;
      gpuCheck (pbwtgpu_long_within (gpuContext (), arrp(p->yz,0,uchar), arrayMax(p->yz), p->M, p->N,
				     p->aFstart, L, 0, 0, &rec, &n), "pbwtgpu_long_within") ;
      for (i = 0 ; i < n ; ++i) reportMatchLine (rec[i].ai, rec[i].bi, rec[i].start, rec[i].end) ;
      pbwtgpu_free (rec) ;
    }
  else
    matchMaximalWithin (p, reportMatchLine) ;
}

/* ---- matchSequencesSweep (pbwtMatch.c:363-443) ---- */

void matchSequencesSweep (PBWT *p, PBWT *q, void (*report)(int ai, int bi, int start, int end))
{
  pbwtgpu_match *rec = 0 ;
  pbwtgpu_nomatch *nom = 0 ;
  int64_t n = 0, nn = 0, i, j = 0 ;
  long totLen = 0 ;

  if (q->N != p->N) die ("query length in matchSequences %d != PBWT length %d", q->N, p->N) ;
  if (isCheck)
    { fprintf (logFile, "pbwt_gpu: -check runs the reference CPU path\n") ;
      ref_matchSequencesSweep (p, q, report) ;
      return ;
    }
  gpuCheck (pbwtgpu_match_sweep (gpuContext (), arrp(p->yz,0,uchar), arrayMax(p->yz), p->M, p->N, p->aFstart,
				 arrp(q->yz,0,uchar), arrayMax(q->yz), q->M, q->aFstart,
				 &rec, &n, &nom, &nn), "pbwtgpu_match_sweep") ;
  /* records and "no match" log lines interleave in site order exactly as the reference emits them:
     a query with no match at site k logs before any record that ends at a later site */
  for (i = 0 ; i < n ; ++i)
    { while (j < nn && nom[j].site < rec[i].end)
	{ fprintf (logFile, "no match to query %d value %d at site %d\n", nom[j].query, nom[j].allele, nom[j].site) ; ++j ; }
      (*report) (rec[i].ai, rec[i].bi, rec[i].start, rec[i].end) ;
      totLen += rec[i].end - rec[i].start ;
    }
  for ( ; j < nn ; ++j)
    fprintf (logFile, "no match to query %d value %d at site %d\n", nom[j].query, nom[j].allele, nom[j].site) ;
  /* summary line, pbwtMatch.c:438-439 */
  fprintf (logFile, "Average number of best matches including alternates %.1f, Average length %.1f, Av number per position %.1f\n",
	   n/(double)q->M, totLen/(double)n, totLen/(double)(q->M*q->N)) ;
  pbwtgpu_free (rec) ; pbwtgpu_free (nom) ;
}

void matchSequencesDynamic (PBWT *p, FILE *fp)	/* pbwtMatch.c:352-357 */
{
  PBWT *q = pbwtRead (fp) ;
  matchSequencesSweep (p, q, reportMatchLine) ;
  pbwtDestroy (q) ;
}

/* ---- batched build: what the readers' per-site loop (pbwtIO.c:558-586) becomes when a batch of
   sites is handed over at once.  x is site-major, one byte per allele (the readers' x[] per line);
   returns a PBWT laid out exactly as pbwtReadLineFile leaves it: yz a genuine Array, aFstart identity,
   aFend malloc'd (pbwt.h:35-53, array.h:44-51). ---- */

PBWT *pbwtGpuBuildFromSites (const uchar *x, int M, int N)
{
  int wpc = pbwtgpu_words_per_col (M) ;
  uint32_t *bits = mycalloc ((long)wpc * (N > 0 ? N : 1), uint32_t) ;
  uint8_t *yz = 0 ; int64_t nyz = 0 ;
  PBWT *p = pbwtCreate (M, N) ;
  long k ; int h ;

  for (k = 0 ; k < N ; ++k)
    for (h = 0 ; h < M ; ++h)
      if (x[k*M + h]) bits[k*wpc + (h >> 5)] |= 1u << (h & 31) ;
  p->aFend = myalloc (M, int) ;
  gpuCheck (pbwtgpu_build (gpuContext (), bits, M, N, p->aFstart, PBWTGPU_WANT_YZ, &yz, &nyz, p->aFend, 0, 0, 0),
	    "pbwtgpu_build") ;
  p->yz = arrayCreate (nyz > 0 ? nyz : 1, uchar) ;
  if (nyz > 0)
    { array(p->yz, nyz-1, uchar) = 0 ;	/* sets arrayMax, as pbwtRead does (pbwtIO.c:209) */
      memcpy (arrp(p->yz,0,uchar), yz, nyz) ;
    }
  pbwtgpu_free (yz) ; free (bits) ;
  return p ;
}
