/* shim_selftest.c -- exercises the shim's build paths through the reference's own writer:
 *   shim_selftest in.hapbytes M N out.pbwt             pbwtGpuBuildFromSites (whole batch)
 *   shim_selftest in.hapbytes M N out.pbwt start.i32   the readers' per-site loop (pbwtIO.c:572-574) on a cursor that starts
 *                                                      from the order in start.i32 (M int32) -- the lazy cursor adopts it
 * in.hapbytes: N*M bytes, site-major.  Writes a .pbwt with the reference's pbwtWrite (pbwtIO.c:33-57). */
#include "pbwt.h"

PBWT *pbwtGpuBuildFromSites (const uchar *x, int M, int N) ;
void gpuReaderWriteForwards (PbwtCursor *u) ;
void gpuReaderToAFend (PbwtCursor *u, PBWT *p) ;

char *commandLine = "shim_selftest" ;
FILE *logFile ;
const char *pbwtCommitHash (void) { return "shim" ; }

int main (int argc, char **argv)
{
  if (argc != 5 && argc != 6) { fprintf (stderr, "usage: %s in.hapbytes M N out.pbwt [start.i32]\n", argv[0]) ; return 2 ; }
  int M = atoi (argv[2]), N = atoi (argv[3]) ;
  logFile = stderr ;
  pbwtInit () ;
  uchar *x = myalloc ((long)M*N + 1, uchar) ;
  FILE *f = fopen (argv[1], "rb") ;
  if (!f || fread (x, 1, (long)M*N, f) != (size_t)((long)M*N)) die ("cannot read %s", argv[1]) ;
  fclose (f) ;
  PBWT *p ;
  if (argc == 6)
    { int j, k ;
      p = pbwtCreate (M, 0) ;
      f = fopen (argv[5], "rb") ;
      if (!f || fread (p->aFstart, sizeof(int), M, f) != (size_t)M) die ("cannot read %s", argv[5]) ;
      fclose (f) ;
      PbwtCursor *u = pbwtCursorCreate (p, TRUE, TRUE) ;	/* picks up aFstart (pbwtCore.c:430-432) */
      for (k = 0 ; k < N ; ++k)
	{ for (j = 0 ; j < M ; ++j) u->y[j] = x[(long)k*M + u->a[j]] ;
	  gpuReaderWriteForwards (u) ;
	  ++p->N ;
	}
      gpuReaderToAFend (u, p) ;
      pbwtCursorDestroy (u) ;
    }
  else
    p = pbwtGpuBuildFromSites (x, M, N) ;
  f = fopen (argv[4], "wb") ;
  pbwtWrite (p, f) ;
  fclose (f) ;
  pbwtDestroy (p) ;
  return 0 ;
}
