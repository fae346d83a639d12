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
 * The readers of pbwtIO.c are batched onto the GPU by two more renames applied to pbwtIO.c only (see
 * gpuReaderWriteForwards below).
 * Errors follow the reference's convention: die() (utils.c:31-44).  There is no CPU fallback for the
 * GPU-backed paths: under -check the matches still come from the GPU and are verified one by one against the decoded
 * haplotypes, as the reference's reportMatch does (pbwtMatch.c:33-58), which makes -check a self-test of the device
 * results.  The reference's renamed originals stay linked but are not called.
 */
#include <time.h>
#include "pbwt.h"
#include "pbwtgpu.h"

/* the reference's originals, renamed at compile time (see above) */
void ref_matchMaximalWithin (PBWT *p, void (*report)(int, int, int, int)) ;
void ref_matchSequencesSweep (PBWT *p, PBWT *q, void (*report)(int, int, int, int)) ;
void ref_pbwtLongMatches (PBWT *p, int L) ;
void ref_matchSequencesDynamic (PBWT *p, FILE *fp) ;

static pbwtgpu_ctx *gpu = 0 ;

/* PBWTGPU_SHIM_TIMING=1: where the drop-in's wall time goes (device context, device call, callbacks), on stderr */
static double shimNow (void)
{ struct timespec ts ; clock_gettime (CLOCK_MONOTONIC, &ts) ; return ts.tv_sec + 1e-9 * ts.tv_nsec ; }

static void shimTiming (const char *what, double t0)
{
  static int on = -1 ;
  if (on < 0) on = getenv ("PBWTGPU_SHIM_TIMING") != 0 ;
  if (on) fprintf (stderr, "pbwtgpu shim timing: %s %.3f s\n", what, shimNow () - t0) ;
}

static pbwtgpu_ctx *gpuContext (void)
{
  if (!gpu)
    { double t0 = shimNow () ;
      int rc = pbwtgpu_init (0, &gpu) ;
      if (rc) die ("pbwtgpu_init failed: %s", pbwtgpu_strerror (rc)) ;
      shimTiming ("pbwtgpu_init (device context)", t0) ;
    }
  return gpu ;
}

static void gpuCheck (int rc, const char *what)
{ if (rc) die ("%s failed: %s: %s", what, pbwtgpu_strerror (rc), pbwtgpu_last_error (gpu)) ; }

/* ---- matchMaximalWithin (pbwtMatch.c:115-142) ---- */

static void reportMatchLine (int ai, int bi, int start, int end) ;	/* the CLI printer, below */
static void reportMatchGroup (const pbwtgpu_group *h, const int32_t *bi) ;	/* the same lines, a whole group at a time */

void matchMaximalWithin (PBWT *p, void (*report)(int ai, int bi, int start, int end))
{
  int32_t *g = 0 ;		/* grouped records: {ai, start, end, n} + n x bi per run of callbacks that share ai, start, end */
  int64_t nw = 0, i = 0 ;

  pbwtgpu_ctx *ctx ;
  double t0 ;

  if (!p || !p->yz) die ("matchMaximalWithin called without a PBWT") ;
  ctx = gpuContext () ;
  t0 = shimNow () ;
  gpuCheck (pbwtgpu_max_within_grouped (ctx, arrp(p->yz,0,uchar), arrayMax(p->yz), p->M, p->N,
					p->aFstart, 0, 0, &g, &nw), "pbwtgpu_max_within_grouped") ;
  shimTiming ("pbwtgpu_max_within_grouped (H2D, sweep, D2H)", t0) ;
  t0 = shimNow () ;
  while (i < nw)		/* one callback per record, in the reference's order */
    { const pbwtgpu_group *h = (const pbwtgpu_group*) (g + i) ;
      const int32_t *bi = g + i + 4 ;
      int j ;
      if (report == reportMatchLine && !isCheck)	/* the printer: what a group's lines share is formatted once */
	reportMatchGroup (h, bi) ;
      else
	for (j = 0 ; j < h->n ; ++j) (*report) (h->ai, bi[j], h->start, h->end) ;
      i += 4 + h->n ;
    }
  shimTiming ("report() callbacks", t0) ;
  pbwtgpu_free (g) ;
}

/* ---- reportMatch (pbwtMatch.c:46-58): the CLI printer; zero-length matches are dropped ---- */

static uchar **checkHapsA, **checkHapsB ;	/* -check only: decoded haplotypes of the query side and the panel side */
static int Ncheck ;

static void checkMatchMaximal (uchar *x, uchar *y, int start, int end, int N)	/* pbwtMatch.c:33-44 */
{
  int i ;

  if (start && (x[start-1] == y[start-1]))
    die ("match not maximal - can extend backwards\n") ;
  if (end < N && (x[end] == y[end]))
    die ("match not maximal - can extend forwards\n") ;
  for (i = start ; i < end ; ++i)
    if (x[i] != y[i])
      die ("match not a match at %d\n", i) ;
}

static void freeHaps (uchar **h, int M)
{ int j ; for (j = 0 ; j < M ; ++j) free (h[j]) ; free (h) ; }

/* the printer's lines are assembled in a buffer of our own (18.5 M printf calls are most of the drop-in's wall time on the
   10,000 x 100,000 panel); matchLinesFlush () hands it to stdout before anyone else writes there */
static char matchLineBuf[1 << 16] ;
static int matchLineFill = 0 ;

static void matchLinesFlush (void)
{ if (matchLineFill) { fwrite (matchLineBuf, 1, matchLineFill, stdout) ; matchLineFill = 0 ; } }

static inline char *putInt (char *c, int v)	/* decimal, as %d prints it */
{
  char t[12] ; int n = 0 ;
  unsigned u = v < 0 ? 0u - (unsigned) v : (unsigned) v ;
  if (v < 0) *c++ = '-' ;
  do { t[n++] = (char) ('0' + u % 10) ; u /= 10 ; } while (u) ;
  while (n) *c++ = t[--n] ;
  return c ;
}

static void reportMatchLine (int ai, int bi, int start, int end)
{
  char *c ;
  if (start == end) return ;
  if (matchLineFill > (int) sizeof (matchLineBuf) - 80) matchLinesFlush () ;
  c = matchLineBuf + matchLineFill ;	/* "MATCH\t%d\t%d\t%d\t%d\t%d\n", pbwtMatch.c:46-49 */
  memcpy (c, "MATCH\t", 6) ; c += 6 ;
  c = putInt (c, ai) ; *c++ = '\t' ;
  c = putInt (c, bi) ; *c++ = '\t' ;
  c = putInt (c, start) ; *c++ = '\t' ;
  c = putInt (c, end) ; *c++ = '\t' ;
  c = putInt (c, end-start) ; *c++ = '\n' ;
  matchLineFill = (int) (c - matchLineBuf) ;
  if (isCheck)			/* the GPU's record is a real match and maximal */
    { matchLinesFlush () ;	/* (a failed check dies: what was reported so far must be out, as with the reference's printf) */
      checkMatchMaximal (checkHapsA[ai], checkHapsB[bi], start, end, Ncheck) ;
    }
}

/* the lines of one group -- n callbacks that share ai, start and end -- exactly as reportMatchLine would print them one
   by one: "MATCH\t<ai>\t" and "\t<start>\t<end>\t<end-start>\n" are formatted once, only bi changes from line to line */
static void reportMatchGroup (const pbwtgpu_group *h, const int32_t *bi)
{
  char head[24], tail[48], *c ;
  int nh, nt, j ;
  if (h->start == h->end) return ;	/* zero-length matches are dropped (pbwtMatch.c:48) */
  c = head ; memcpy (c, "MATCH\t", 6) ; c += 6 ; c = putInt (c, h->ai) ; *c++ = '\t' ; nh = (int) (c - head) ;
  c = tail ; *c++ = '\t' ; c = putInt (c, h->start) ; *c++ = '\t' ; c = putInt (c, h->end) ; *c++ = '\t' ;
  c = putInt (c, h->end - h->start) ; *c++ = '\n' ; nt = (int) (c - tail) ;
  for (j = 0 ; j < h->n ; ++j)
    { if (matchLineFill > (int) sizeof (matchLineBuf) - 96) matchLinesFlush () ;
      c = matchLineBuf + matchLineFill ;
      memcpy (c, head, nh) ; c += nh ;
      c = putInt (c, bi[j]) ;
      memcpy (c, tail, nt) ; c += nt ;
      matchLineFill = (int) (c - matchLineBuf) ;
    }
}

/* ---- pbwtLongMatches (pbwtMatch.c:148-183): -maxWithin is L == 0 ---- */

void pbwtLongMatches (PBWT *p, int L)
{
  if (!p || !p->yz) die ("option -longWithin called without a PBWT") ;
  if (L < 0) die ("L %d for longWithin must be >= 0", L) ;
  if (isCheck) { checkHapsA = checkHapsB = pbwtHaplotypes (p) ; Ncheck = p->N ; }	/* pbwtMatch.c:156 */
  if (isStats)			/* -stats (pbwtMatch.c:130-131, 161-178): with -maxWithin a histogram of the reporters' match
				   lengths replaces the MATCH lines (accumulated on the GPU, one count per reporter at the length
				   of the longer of its two blocks); matchLongWithin2 never touches the histogram, so with
				   -longWithin L the MATCH lines are printed and the summary is that of an empty histogram
				   (the reference's "Average 0.0 matches per sample" and 0/0 length) */
    { pbwtgpu_match *rec = 0 ;
      int64_t n = 0, i = 0 ;
      Array hist = arrayCreate (1000000, int) ;
      int nTot = 0, len ;
      long hTot = 0 ;
      if (L)
	{ gpuCheck (pbwtgpu_long_within (gpuContext (), arrp(p->yz,0,uchar), arrayMax(p->yz), p->M, p->N,
					 p->aFstart, L, 0, 0, &rec, &n), "pbwtgpu_long_within") ;
	  for (i = 0 ; i < n ; ++i) reportMatchLine (rec[i].ai, rec[i].bi, rec[i].start, rec[i].end) ;
	}
      else			/* the matches are consumed on the device: only the N+1 counters come back */
	{ int64_t *h = myalloc (p->N + 1, int64_t) ;
	  gpuCheck (pbwtgpu_max_within_stats (gpuContext (), arrp(p->yz,0,uchar), arrayMax(p->yz), p->M, p->N,
					      p->aFstart, h), "pbwtgpu_max_within_stats") ;
	  for (len = 0 ; len <= p->N ; ++len) if (h[len]) array(hist, len, int) = (int) h[len] ;
	  free (h) ;
	}
      matchLinesFlush () ;
      for (len = 0 ; len < arrayMax(hist) ; ++len)
	if (arr(hist, len, int))
	  { nTot += arr(hist, len, int) ;
	    hTot += arr(hist, len, int) * len ;
	    printf ("%d\t%d\n", len, arr(hist, len, int)) ;
	  }
      fprintf (logFile, "Average %.1f matches per sample\n", nTot/(double)p->M) ;
      fprintf (logFile, "Average length %.1f\n", hTot/(double)nTot) ;
      arrayDestroy (hist) ;
      pbwtgpu_free (rec) ;
    }
  else if (L)			/* matchLongWithin2 (pbwtMatch.c:85-113) */
    { pbwtgpu_match *rec = 0 ;
      int64_t n = 0, i ;
      gpuCheck (pbwtgpu_long_within (gpuContext (), arrp(p->yz,0,uchar), arrayMax(p->yz), p->M, p->N,
				     p->aFstart, L, 0, 0, &rec, &n), "pbwtgpu_long_within") ;
      for (i = 0 ; i < n ; ++i) reportMatchLine (rec[i].ai, rec[i].bi, rec[i].start, rec[i].end) ;
      pbwtgpu_free (rec) ;
    }
  else
    matchMaximalWithin (p, reportMatchLine) ;
  matchLinesFlush () ;
  if (isCheck) { freeHaps (checkHapsA, p->M) ; checkHapsA = checkHapsB = 0 ; }
}

/* ---- pbwtBuildReverse (pbwtCore.c:151-191); pbwtCore.c is compiled with -DpbwtBuildReverse=ref_pbwtBuildReverse ---- */

void ref_pbwtBuildReverse (PBWT *p) ;

void pbwtBuildReverse (PBWT *p)
{
  uint8_t *zz = 0 ; int64_t nzz = 0 ;
  int *aFend, *aRend ;

  aFend = myalloc (p->M, int) ; aRend = myalloc (p->M, int) ;
  gpuCheck (pbwtgpu_build_reverse (gpuContext (), arrp(p->yz,0,uchar), arrayMax(p->yz), p->M, p->N, p->aFstart,
				   aFend, &zz, &nzz, aRend), "pbwtgpu_build_reverse") ;
  if (!p->aFend) p->aFend = aFend ; else free (aFend) ;	/* an aFend read from the file is kept, :157-164 */
  if (!p->aRstart) p->aRstart = myalloc (p->M, int) ;
  memcpy (p->aRstart, p->aFend, p->M * sizeof(int)) ;
  p->zz = arrayReCreate (p->zz, nzz > 0 ? nzz : 1, uchar) ;
  if (nzz > 0)
    { array(p->zz, nzz-1, uchar) = 0 ;
      memcpy (arrp(p->zz,0,uchar), zz, nzz) ;
    }
  if (p->aRend) free (p->aRend) ;
  p->aRend = aRend ;
  fprintf (logFile, "built reverse PBWT - size %ld\n", arrayMax(p->zz)) ;
  pbwtgpu_free (zz) ;
  if (isCheck)			/* print out the reversed haplotypes, pbwtCore.c:181-187 */
    { FILE *fp = fopen ("rev.haps","w") ;
      Array tz = p->yz ; int *ta = p->aFstart ;
      p->yz = p->zz ; p->aFstart = p->aRstart ;
      pbwtWriteHaplotypes (fp, p) ;
      p->yz = tz ; p->aFstart = ta ;
      fclose (fp) ;
    }
}

/* ---- matchSequencesSweep (pbwtMatch.c:363-443) ---- */

static void logNoMatch (const pbwtgpu_nomatch *e)
{ fprintf (logFile, "no match to query %d value %d at site %d\n", e->query, e->allele, e->site) ; }

/* More queries than one device call takes (8,192; PBWTGPU_QUERY_CHUNK overrides, for tests).  A query's matches depend
   on the panel and on its own alleles only; the other queries decide nothing but the ORDER of the callbacks (queries
   are visited in the order of the query PBWT at each site, pbwtMatch.c:377-378,431-432).  So the queries are cut into
   chunks of consecutive haplotype indices; one walk over q's stream (the reference's run-wise cursor,
   pbwtCursorForwardsRead) un-sorts every column into per-chunk bit columns; each chunk is built and swept against the
   panel on the GPU; a second walk supplies every query's position in the query order at each site, and the chunks'
   record streams -- each already in that order, restricted to its chunk -- are merged site by site. */

typedef struct { int rank, chunk, len ; int64_t start ; } QGroup ;

static int qGroupOrder (const void *a, const void *b)
{ return ((const QGroup*)a)->rank - ((const QGroup*)b)->rank ; }

static void matchSweepChunked (PBWT *p, PBWT *q, int chunk, void (*report)(int ai, int bi, int start, int end),
			       int64_t *nOut, long *totLenOut)
{
  int Q = q->M, N = q->N, nC = (Q + chunk - 1) / chunk, c, j, k ;	/* chunk: the most one device call takes */
  uint32_t **bits = myalloc (nC, uint32_t*) ;
  pbwtgpu_match **rec = mycalloc (nC, pbwtgpu_match*) ;
  pbwtgpu_nomatch **nom = mycalloc (nC, pbwtgpu_nomatch*) ;
  int64_t *n = mycalloc (nC, int64_t), *nn = mycalloc (nC, int64_t), *ir = mycalloc (nC, int64_t), *in = mycalloc (nC, int64_t) ;
  int64_t i, total = 0 ;
  long totLen = 0 ;
  int *rank = myalloc (Q, int), *wpc = myalloc (nC, int), *cstart = myalloc (nC+1, int), *chunkOf = myalloc (Q, int) ;
  QGroup *grp = myalloc (Q, QGroup), *ngrp = myalloc (Q, QGroup) ;
  PbwtCursor *uq ;

  for (c = 0 ; c <= nC ; ++c) cstart[c] = (int) ((long)c * Q / nC) ;	/* chunk c = queries [cstart[c], cstart[c+1]): sizes differ by at most one */
  for (c = 0 ; c < nC ; ++c) for (j = cstart[c] ; j < cstart[c+1] ; ++j) chunkOf[j] = c ;

  /* 1. un-sort the query columns into per-chunk bit columns */
  for (c = 0 ; c < nC ; ++c)
    { wpc[c] = pbwtgpu_words_per_col (cstart[c+1] - cstart[c]) ;
      bits[c] = mycalloc ((size_t)wpc[c] * (N > 0 ? N : 1), uint32_t) ;
    }
  uq = pbwtCursorCreate (q, TRUE, TRUE) ;
  for (k = 0 ; k < N ; ++k)
    { for (j = 0 ; j < Q ; ++j)
	if (uq->y[j])
	  { int h = uq->a[j], l ;
	    c = chunkOf[h] ; l = h - cstart[c] ;
	    bits[c][(size_t)k * wpc[c] + (l >> 5)] |= 1u << (l & 31) ;
	  }
      pbwtCursorForwardsRead (uq) ;
    }
  pbwtCursorDestroy (uq) ;
  free (wpc) ; free (chunkOf) ;

  /* 2. every chunk: its own PBWT (started from q's start order restricted to the chunk), swept against the panel */
  for (c = 0 ; c < nC ; ++c)
    { int c0 = cstart[c], cq = cstart[c+1] - c0, m = 0 ;
      int *aStart = myalloc (cq, int), *aEnd = myalloc (cq, int), *aOrder ;
      uint8_t *yz = 0 ; int64_t nyz = 0 ;
      for (j = 0 ; j < Q ; ++j)
	{ int h = q->aFstart ? q->aFstart[j] : j ;
	  if (h >= c0 && h < c0 + cq) aStart[m++] = h - c0 ;
	}
      for (j = 0 ; j < cq && aStart[j] == j ; ++j) ;
      aOrder = j < cq ? aStart : 0 ;	/* the usual case: q starts from the identity, and so does every chunk */
      gpuCheck (pbwtgpu_build (gpuContext (), bits[c], cq, N, aOrder, PBWTGPU_WANT_YZ, &yz, &nyz, aEnd, 0, 0, 0),
		"pbwtgpu_build (query chunk)") ;
      free (bits[c]) ;
      gpuCheck (pbwtgpu_match_sweep (gpuContext (), arrp(p->yz,0,uchar), arrayMax(p->yz), p->M, p->N, p->aFstart,
				     yz, nyz, cq, aOrder, &rec[c], &n[c], &nom[c], &nn[c]), "pbwtgpu_match_sweep (query chunk)") ;
      for (i = 0 ; i < n[c] ; ++i) rec[c][i].ai += c0 ;
      for (i = 0 ; i < nn[c] ; ++i) nom[c][i].query += c0 ;
      pbwtgpu_free (yz) ; free (aStart) ; free (aEnd) ;
    }
  free (bits) ; free (cstart) ;

  /* 3. merge site by site in the order of the whole query PBWT */
  uq = pbwtCursorCreate (q, TRUE, TRUE) ;
  for (k = 0 ; k <= N ; ++k)
    { int ng = 0, nng = 0, g ;
      BOOL any = FALSE ;
      for (c = 0 ; c < nC && !any ; ++c)
	any = (ir[c] < n[c] && rec[c][ir[c]].end == k) || (in[c] < nn[c] && nom[c][in[c]].site == k) ;
      if (any)
	{ for (j = 0 ; j < Q ; ++j) rank[uq->a[j]] = j ;
	  for (c = 0 ; c < nC ; ++c)
	    { while (ir[c] < n[c] && rec[c][ir[c]].end == k)	/* one group = one query's records at this site */
		{ int ai = rec[c][ir[c]].ai ;
		  if (ng >= Q) die ("matchSweepChunked: more record groups than queries at site %d", k) ;
		  grp[ng].rank = rank[ai] ; grp[ng].chunk = c ; grp[ng].start = ir[c] ; grp[ng].len = 0 ;
		  while (ir[c] < n[c] && rec[c][ir[c]].end == k && rec[c][ir[c]].ai == ai) { ++ir[c] ; ++grp[ng].len ; }
		  ++ng ;
		}
	      while (in[c] < nn[c] && nom[c][in[c]].site == k)
		{ if (nng >= Q) die ("matchSweepChunked: more no-match events than queries at site %d", k) ;
		  ngrp[nng].rank = rank[nom[c][in[c]].query] ; ngrp[nng].chunk = c ; ngrp[nng].start = in[c]++ ; ngrp[nng].len = 1 ;
		  ++nng ;
		}
	    }
	  qsort (grp, ng, sizeof(QGroup), qGroupOrder) ;
	  qsort (ngrp, nng, sizeof(QGroup), qGroupOrder) ;
	  for (g = 0 ; g < ng ; ++g)
	    { const pbwtgpu_match *r = rec[grp[g].chunk] + grp[g].start ;
	      for (j = 0 ; j < grp[g].len ; ++j)
		{ (*report) (r[j].ai, r[j].bi, r[j].start, r[j].end) ;
		  totLen += r[j].end - r[j].start ;
		}
	      total += grp[g].len ;
	    }
	  for (g = 0 ; g < nng ; ++g) logNoMatch (nom[ngrp[g].chunk] + ngrp[g].start) ;
	}
      if (k < N) pbwtCursorForwardsRead (uq) ;
    }
  pbwtCursorDestroy (uq) ;
  for (c = 0 ; c < nC ; ++c)
    { if (ir[c] != n[c] || in[c] != nn[c]) die ("matchSweepChunked: chunk %d left records behind", c) ;
      pbwtgpu_free (rec[c]) ; pbwtgpu_free (nom[c]) ;
    }
  free (rec) ; free (nom) ; free (n) ; free (nn) ; free (ir) ; free (in) ; free (rank) ; free (grp) ; free (ngrp) ;
  *nOut = total ; *totLenOut = totLen ;
}

void matchSequencesSweep (PBWT *p, PBWT *q, void (*report)(int ai, int bi, int start, int end))
{
  pbwtgpu_match *rec = 0 ;
  pbwtgpu_nomatch *nom = 0 ;
  int64_t n = 0, nn = 0, i, j = 0 ;
  long totLen = 0 ;
  int chunk = 8192 ;		/* queries per device call (include/pbwtgpu.h: pbwtgpu_match_sweep) */

  if (q->N != p->N) die ("query length in matchSequences %d != PBWT length %d", q->N, p->N) ;
  if (isCheck) { checkHapsA = pbwtHaplotypes (q) ; checkHapsB = pbwtHaplotypes (p) ; Ncheck = p->N ; }	/* pbwtMatch.c:373 */
  if (getenv ("PBWTGPU_QUERY_CHUNK") && atoi (getenv ("PBWTGPU_QUERY_CHUNK")) > 0)
    chunk = atoi (getenv ("PBWTGPU_QUERY_CHUNK")) ;
  if (q->M > chunk)
    matchSweepChunked (p, q, chunk, report, &n, &totLen) ;
  else
    { gpuCheck (pbwtgpu_match_sweep (gpuContext (), arrp(p->yz,0,uchar), arrayMax(p->yz), p->M, p->N, p->aFstart,
				     arrp(q->yz,0,uchar), arrayMax(q->yz), q->M, q->aFstart,
				     &rec, &n, &nom, &nn), "pbwtgpu_match_sweep") ;
      /* records and "no match" log lines interleave in site order exactly as the reference emits them:
	 a query with no match at site k logs before any record that ends at a later site */
      for (i = 0 ; i < n ; ++i)
	{ while (j < nn && nom[j].site < rec[i].end) logNoMatch (&nom[j++]) ;
	  (*report) (rec[i].ai, rec[i].bi, rec[i].start, rec[i].end) ;
	  totLen += rec[i].end - rec[i].start ;
	}
      for ( ; j < nn ; ++j) logNoMatch (&nom[j]) ;
      pbwtgpu_free (rec) ; pbwtgpu_free (nom) ;
    }
  /* summary line, pbwtMatch.c:438-439 */
  fprintf (logFile, "Average number of best matches including alternates %.1f, Average length %.1f, Av number per position %.1f\n",
	   n/(double)q->M, totLen/(double)n, totLen/(double)(q->M*q->N)) ;
  if (isCheck) { freeHaps (checkHapsA, q->M) ; freeHaps (checkHapsB, p->M) ; checkHapsA = checkHapsB = 0 ; }
}

void matchSequencesDynamic (PBWT *p, FILE *fp)	/* pbwtMatch.c:352-357 */
{
  PBWT *q = pbwtRead (fp) ;
  matchSequencesSweep (p, q, reportMatchLine) ;
  matchLinesFlush () ;
  pbwtDestroy (q) ;
}

/* ---- batched ingest: the readers of pbwtIO.c (pbwtReadLineFile :558-586, pbwtReadLineTwoFile :595-623,
   pbwtReadMacs :460-492, the phase reader :820-826, pbwtCheckPoint :158-168) all run, per parsed site,
       for (j = 0 ; j < M ; ++j) u->y[j] = x[u->a[j]] ;  pbwtCursorWriteForwards (u) ;
   and finish with pbwtCursorToAFend (u, p).  pbwtIO.c is compiled with
       -DpbwtCursorWriteForwards=gpuReaderWriteForwards -DpbwtCursorToAFend=gpuReaderToAFend
   so those calls land here ("lazy cursor"): u->a is left at the identity, which makes the gathered u->y the
   UNSORTED column x; columns are bit-packed into a batch, and a full batch (or the final partial one) is swept
   on the GPU by pbwtgpu_build from the true order carried between batches, the pack3 bytes are appended to
   u->z exactly where pack3arrayAdd would have put them, and ToAFend hands the true order to p->aFend.
   The readers themselves are untouched. ---- */

void pbwtCursorToAFend (PbwtCursor *u, PBWT *p) ;	/* the reference's own (pbwtCore.c:587-591) */

static struct
{ PbwtCursor *u ;		/* the reader's cursor being batched */
  int M, wpc, cap, n ;		/* batch capacity and fill, in sites */
  long written ;		/* arrayMax(u->z) as this batcher left it: anything else means another cursor at a recycled address */
  uint32_t *bits ;		/* page-locked (pbwtgpu_host_alloc) */
  int *aTrue ;			/* the real a[] entering the batch */
  BOOL startIsIdentity ;	/* u->a, frozen at adoption, is the identity: u->y is the column as parsed */
} rb ;

static void readerFlush (PbwtCursor *u)
{
  uint8_t *yz = 0 ; int64_t nyz = 0 ;
  int *aNew ;
  long max ;

  if (!rb.n) return ;
  aNew = myalloc (rb.M, int) ;
  gpuCheck (pbwtgpu_build (gpuContext (), rb.bits, rb.M, rb.n, rb.aTrue, PBWTGPU_WANT_YZ, &yz, &nyz, aNew, 0, 0, 0),
	    "pbwtgpu_build") ;
  max = arrayMax(u->z) ;
  arrayExtend (u->z, max + nyz) ;	/* as pack3arrayAdd does, pbwtCore.c:270-277 */
  memcpy (arrp(u->z,max,uchar), yz, nyz) ;
  arrayMax(u->z) = max + nyz ;
  u->n += nyz ;
  rb.written = max + nyz ;
  memcpy (rb.aTrue, aNew, rb.M*sizeof(int)) ;
  memset (rb.bits, 0, (size_t)rb.wpc * rb.cap * sizeof(uint32_t)) ;
  rb.n = 0 ;
  pbwtgpu_free (yz) ; free (aNew) ;
}

/* Any cursor is adopted, whatever order it starts from and whether or not its stream is empty (there is no CPU path):
   u->a is frozen at adoption, so what the reader gathers is y[j] = x[a0[j]] with a0 that frozen order, and the parsed
   column is recovered as x[a0[j]] = y[j]; the true order starts as a copy of a0 and lives in rb.aTrue. */
void gpuReaderWriteForwards (PbwtCursor *u)
{
  int j ;
  uint32_t *col ;

  /* a cursor is recognised by its address, which the allocator recycles: it is the one being batched only if its
     width and the size of its stream are what this batcher left behind; an untouched pair (nothing written, nothing
     pending) is always (re)adopted, which costs nothing */
  if (u != rb.u || u->M != rb.M || arrayMax(u->z) != rb.written || (!rb.n && !rb.written))
    { if (rb.u && rb.n) die ("gpuReaderWriteForwards: a batch of another cursor is still pending") ;
      pbwtgpu_free (rb.bits) ; free (rb.aTrue) ;
      rb.u = u ; rb.M = u->M ; rb.wpc = pbwtgpu_words_per_col (u->M) ; rb.n = 0 ; rb.written = arrayMax(u->z) ;
      { const char *e = getenv ("PBWTGPU_READER_BATCH") ;	/* sites per batch; default bounded by 256 MB of columns */
	long cap = e ? atol (e) : 65536 ;
	long most = (256L << 20) / (4L * rb.wpc) ;
	if (cap > most) cap = most ;
	if (cap < 1) cap = 1 ;
	rb.cap = (int) cap ;
      }
      /* page-locked, so that the batch goes up at the full PCIe rate (pbwtgpu_host_alloc) */
      rb.bits = (uint32_t*) pbwtgpu_host_alloc ((size_t)rb.wpc * rb.cap * sizeof(uint32_t)) ;
      if (!rb.bits) die ("gpuReaderWriteForwards: no page-locked memory for a batch of %d sites", rb.cap) ;
      memset (rb.bits, 0, (size_t)rb.wpc * rb.cap * sizeof(uint32_t)) ;
      rb.aTrue = myalloc (rb.M, int) ;
      memcpy (rb.aTrue, u->a, rb.M*sizeof(int)) ;
      rb.startIsIdentity = TRUE ;
      for (j = 0 ; j < rb.M ; ++j) if (u->a[j] != j) { rb.startIsIdentity = FALSE ; break ; }
    }
  col = rb.bits + (size_t)rb.n * rb.wpc ;
  if (rb.startIsIdentity)
    { for (j = 0 ; j < rb.M ; ++j) if (u->y[j]) col[j >> 5] |= 1u << (j & 31) ; }
  else
    for (j = 0 ; j < rb.M ; ++j) if (u->y[j]) { int h = u->a[j] ; col[h >> 5] |= 1u << (h & 31) ; }
  u->isBlockEnd = FALSE ;
  if (++rb.n == rb.cap) readerFlush (u) ;
}

void gpuReaderToAFend (PbwtCursor *u, PBWT *p)
{
  /* only the cursor this batch belongs to: same address, same width, and a stream exactly as this batcher left it (a
     later reader that landed on the recycled address and wrote nothing must not inherit the previous order) */
  if (u != rb.u || u->M != rb.M || p->M != rb.M || arrayMax(u->z) != rb.written)
    { pbwtCursorToAFend (u, p) ; return ; }
  readerFlush (u) ;
  if (!p->aFend) p->aFend = myalloc (p->M, int) ;
  memcpy (p->aFend, rb.aTrue, p->M*sizeof(int)) ;
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
