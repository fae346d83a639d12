/* vcf_text.c -- `-readVcfGT` for plain-text VCF without htslib, feeding the batched device build.
 *
 * The reference reads VCF/BCF through htslib (pbwtHtslib.c:52-172), an un-vendored third-party library that is not
 * part of this image.  This file provides the same entry point, PBWT *pbwtReadVcfGT (char *filename), for what the
 * reference's own fixtures use (test/OMNI.vcf, BASELINE config 1: `pbwt -readVcfGT test/OMNI.vcf -maxWithin`):
 * text VCF with a GT field, plain or gzip/bgzip-compressed (zlib, which the reference links anyway).  It follows pbwtReadVcfGT step by step -- samples (:30-35), one chromosome
 * only (:78-79), upper-cased REF/ALT (:81-83,141-143), haploid lines as homozygous diploid (:94-110), missing alleles
 * as 0 plus the zMissing / missingOffset arrays (:111-125,150-161), multi-allelic records split into one site per
 * ALT allele with y = (x == i) (:127-168), ALT "." kept as a site with no carriers, the checkpoint hook (:170) -- and
 * hands every column to the shim's lazy cursor (gpuReaderWriteForwards / gpuReaderToAFend, pbwt_gpu_shim.c), so the
 * per-site gather + pbwtCursorWriteForwards of :145-146 runs batched on the GPU like the other readers.
 * BCF, -readVcfPL and -writeVcf still need htslib and die() saying so.
 */
#define _GNU_SOURCE	/* getline, strndup */
#include <ctype.h>
#include <sys/types.h>
#include <zlib.h>
#include "pbwt.h"

void gpuReaderWriteForwards (PbwtCursor *u) ;	/* pbwt_gpu_shim.c */
void gpuReaderToAFend (PbwtCursor *u, PBWT *p) ;

const char *pbwtHtslibVersionString (void) { return "none (text VCF reader of pbwt-b200)" ; }

PBWT *pbwtReadVcfPL (char *filename)
{ die ("pbwtReadVcfPL(%s): htslib is not available in this build", filename) ; return 0 ; }

void pbwtWriteVcf (PBWT *p, char *filename, char *reference_fname, char *mode)
{ (void)p ; (void)reference_fname ; (void)mode ; die ("pbwtWriteVcf(%s): htslib is not available in this build", filename) ; }

static int variation (const char *ref, const char *alt)	/* pbwtHtslib.c:37-50 */
{
  char *buf = myalloc (strlen (ref) + strlen (alt) + 2, char) ;
  int var ;
  sprintf (buf, "%s\t%s", ref, alt) ;
  dictAdd (variationDict, buf, &var) ;
  free (buf) ;
  return var ;
}

static char *upperCopy (const char *s, int n)
{
  char *t = myalloc (n+1, char) ;
  int i ;
  for (i = 0 ; i < n ; ++i) t[i] = toupper ((unsigned char) s[i]) ;
  t[n] = 0 ;
  return t ;
}

/* next tab-separated field of a line: returns its start, sets *len, advances *cp past the separator */
static char *nextField (char **cp, int *len)
{
  char *s = *cp, *e = s ;
  if (!s) { *len = 0 ; return 0 ; }
  while (*e && *e != '\t' && *e != '\n' && *e != '\r') ++e ;
  *len = e - s ;
  *cp = (*e == '\t') ? e+1 : 0 ;
  return s ;
}

/* getline() over a gzFile: the whole line including its newline, buffer grown as needed; 0 at end of file */
static ssize_t gzGetLine (char **line, size_t *cap, gzFile fp)
{
  size_t n = 0 ;
  if (!*line) { *cap = 1 << 16 ; *line = myalloc (*cap, char) ; }
  for (;;)
    { if (!gzgets (fp, *line + n, (int)(*cap - n))) break ;
      n += strlen (*line + n) ;
      if (n && (*line)[n-1] == '\n') break ;
      if (n + 1 >= *cap) { char *t = myalloc (*cap * 2, char) ; memcpy (t, *line, n + 1) ; free (*line) ; *line = t ; *cap *= 2 ; }
      else if (gzeof (fp)) break ;	/* last line without a newline */
    }
  return (ssize_t) n ;
}

PBWT *pbwtReadVcfGT (char *filename)
{
  gzFile fp = strcmp (filename, "-") ? gzopen (filename, "r") : gzdopen (0, "r") ;	/* reads plain text as well */
  char *line = 0 ; size_t cap = 0 ; ssize_t got ;
  PBWT *p = 0 ;
  PbwtCursor *u = 0 ;
  uchar *x = 0, *xMissing = 0 ;
  int *gt = 0 ;			/* alleles of the record, -1 = missing, -2 = absent (lower ploidy) */
  long nMissing = 0 ;
  int nMissingSites = 0, nSamples = 0, i, j ;

  if (!fp) die ("failed to open good vcf file\n") ;
  gzbuffer (fp, 1 << 20) ;
  if (strcmp (filename, "-"))		/* (a pipe cannot be rewound) */
    { int c0 = gzgetc (fp), c1 = gzgetc (fp), c2 = gzgetc (fp) ;
      if (c0 == 'B' && c1 == 'C' && c2 == 'F') die ("pbwtReadVcfGT(%s): BCF needs htslib, which is not available in this build", filename) ;
      gzrewind (fp) ;
    }

  while ((got = gzGetLine (&line, &cap, fp)) > 0)
    { char *cp = line, *f ; int len ;
      if (line[0] == '#' && line[1] == '#') continue ;
      if (line[0] == '#')			/* #CHROM POS ID REF ALT QUAL FILTER INFO FORMAT samples... */
	{ Array names = arrayCreate (1024, char*) ;
	  for (i = 0 ; (f = nextField (&cp, &len)) ; ++i)
	    if (i >= 9 && len) array(names, arrayMax(names), char*) = strndup (f, len) ;
	  nSamples = arrayMax(names) ;
	  if (!nSamples) die ("no samples in the header of %s", filename) ;
	  p = pbwtCreate (nSamples*2, 0) ;	/* assume diploid! (:60) */
	  p->samples = arrayCreate (p->M, int) ;
	  for (i = 0 ; i < nSamples ; ++i)
	    { int k = sampleAdd (arr(names, i, char*), 0, 0, 0) ;
	      array(p->samples, 2*i, int) = k ;
	      array(p->samples, 2*i+1, int) = k ;
	      free (arr(names, i, char*)) ;
	    }
	  arrayDestroy (names) ;
	  p->sites = arrayCreate (10000, Site) ;
	  u = pbwtCursorCreate (p, TRUE, TRUE) ;
	  x = myalloc (p->M, uchar) ;
	  xMissing = myalloc (p->M+1, uchar) ;
	  xMissing[p->M] = Y_SENTINEL ;
	  gt = myalloc (p->M, int) ;
	  continue ;
	}
      if (!p) die ("%s: a record before the #CHROM header line", filename) ;
      if (line[0] == '\n' || line[0] == '\r') continue ;

      /* fixed fields */
      char *chromF, *posF, *refF, *altF, *fmtF ; int chromL, posL, refL, altL, fmtL ;
      chromF = nextField (&cp, &chromL) ; posF = nextField (&cp, &posL) ; nextField (&cp, &len) ;
      refF = nextField (&cp, &refL) ; altF = nextField (&cp, &altL) ;
      nextField (&cp, &len) ; nextField (&cp, &len) ; nextField (&cp, &len) ;
      fmtF = nextField (&cp, &fmtL) ;
      if (!fmtF || !cp) die ("%s: short record near site %d", filename, p->N) ;
      if (!p->chrom) p->chrom = strndup (chromF, chromL) ;
      else if (strlen (p->chrom) != (size_t)chromL || strncmp (p->chrom, chromF, chromL)) break ;
      int pos = atoi (posF) ; (void)posL ;

      /* where is GT in FORMAT? */
      int gtIndex = -1, k = 0 ;
      for (i = 0 ; i <= fmtL ; )
	{ int e = i ; while (e < fmtL && fmtF[e] != ':') ++e ;
	  if (e - i == 2 && fmtF[i] == 'G' && fmtF[i+1] == 'T') { gtIndex = k ; break ; }
	  ++k ; i = e + 1 ;
	}
      if (gtIndex < 0) continue ;		/* no GT in the FORMAT (:87) */

      /* genotypes: up to two alleles per sample */
      int maxPloidy = 0 ;
      for (i = 0 ; i < nSamples ; ++i)
	{ f = nextField (&cp, &len) ;
	  if (!f) die ("%s: %d samples expected at %.*s:%d", filename, nSamples, chromL, chromF, pos) ;
	  int s = 0, e, sub = 0 ;
	  while (sub < gtIndex && s < len) { if (f[s] == ':') ++sub ; ++s ; }
	  e = s ; while (e < len && f[e] != ':') ++e ;
	  int ploidy = 0 ;
	  gt[2*i] = gt[2*i+1] = -2 ;
	  while (s < e)
	    { int a ;
	      if (f[s] == '.') { a = -1 ; ++s ; }
	      else { a = 0 ; if (!isdigit ((unsigned char) f[s])) die ("bad genotype at %.*s:%d", chromL, chromF, pos) ;
		     while (s < e && isdigit ((unsigned char) f[s])) a = 10*a + (f[s++] - '0') ; }
	      if (ploidy >= 2) die ("%d != %d GT values at %.*s:%d - not haploid or diploid?", 3*nSamples, p->M, chromL, chromF, pos) ;
	      gt[2*i + ploidy++] = a ;
	      if (s < e && (f[s] == '|' || f[s] == '/')) ++s ;
	    }
	  if (ploidy > maxPloidy) maxPloidy = ploidy ;
	}
      if (maxPloidy == 0) continue ;

      memset (xMissing, 0, p->M) ;
      long wasMissing = nMissing ;
      if (maxPloidy == 1)			/* all haploid: homozygous diploid (:94-110) */
	for (i = 0 ; i < nSamples ; ++i)
	  if (gt[2*i] < 0)
	    { x[2*i] = x[2*i+1] = 0 ; xMissing[2*i] = xMissing[2*i+1] = 1 ; nMissing += 2 ; }
	  else
	    x[2*i] = x[2*i+1] = gt[2*i] ;
      else
	for (i = 0 ; i < p->M ; ++i)
	  { if (gt[i] == -2) die ("unexpected end of genotype vector in VCF") ;
	    if (gt[i] == -1) { x[i] = 0 ; xMissing[i] = 1 ; ++nMissing ; }
	    else x[i] = gt[i] ;
	  }

      /* split into biallelic sites (:127-168) */
      BOOL noAlt = (altL == 1 && altF[0] == '.') ;
      char *REF = upperCopy (refF, refL) ;
      int a0 = 0 ;
      for (i = 1 ; a0 <= altL ; ++i)
	{ int a1 = a0 ; while (a1 < altL && altF[a1] != ',') ++a1 ;
	  char *ALT = noAlt ? strdup (".") : upperCopy (altF + a0, a1 - a0) ;
	  a0 = a1 + 1 ;

	  for (j = 0 ; j < p->M ; ++j) u->y[j] = x[u->a[j]] == i ? 1 : 0 ;
	  gpuReaderWriteForwards (u) ;

	  if (nMissing > wasMissing)
	    { if (!p->zMissing)
		{ p->zMissing = arrayCreate (10000, uchar) ;
		  array(p->zMissing, 0, uchar) = 0 ;	/* needed so missing[] has offset > 0 */
		  p->missingOffset = arrayCreate (1024, long) ;
		}
	      array(p->missingOffset, p->N, long) = arrayMax(p->zMissing) ;
	      pack3arrayAdd (xMissing, p->M, p->zMissing) ;	/* NB original order, not pbwt sort */
	      nMissingSites++ ;
	    }
	  else if (nMissing)
	    array(p->missingOffset, p->N, long) = 0 ;

	  Site *s = arrayp(p->sites, p->N++, Site) ;
	  s->x = pos ;
	  s->varD = variation (REF, ALT) ;
	  free (ALT) ;
	}
      free (REF) ;

      if (nCheckPoint && !(p->N % nCheckPoint)) pbwtCheckPoint (u, p) ;
    }
  { int zerr = Z_OK ;			/* a NULL from gzgets is the end of the file only when zlib agrees: a truncated
					   stream or a bad CRC / length trailer must not yield a silently shorter panel */
    const char *zmsg = gzerror (fp, &zerr) ;
    if (zerr == Z_BUF_ERROR && !gzeof (fp)) die ("pbwtReadVcfGT(%s): compressed input is truncated", filename) ;
    if (zerr != Z_OK && zerr != Z_STREAM_END && zerr != Z_BUF_ERROR)
      die ("pbwtReadVcfGT(%s): error reading compressed input: %s", filename, zmsg ? zmsg : "zlib error") ;
  }
  if (!p) die ("%s: no #CHROM header line", filename) ;
  gpuReaderToAFend (u, p) ;

  free (line) ; free (gt) ; free (x) ; free (xMissing) ; pbwtCursorDestroy (u) ;
  { int zrc = gzclose (fp) ;
    if (zrc != Z_OK && zrc != Z_BUF_ERROR) die ("pbwtReadVcfGT(%s): error closing input (zlib %d)", filename, zrc) ;
  }

  fprintf (logFile, "read genotypes from %s with %ld sample names and %ld sites on chromosome %s: M, N are %d, %d\n",
	   filename, arrayMax(p->samples)/2, arrayMax(p->sites), p->chrom, p->M, p->N) ;
  if (p->missingOffset) fprintf (logFile, "%ld missing values at %d sites\n", nMissing, nMissingSites) ;
  return p ;
}
