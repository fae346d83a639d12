/* shim_selftest.c -- exercises pbwtGpuBuildFromSites through the reference's own writer:
 *   shim_selftest in.hapbytes M N out.pbwt
 * in.hapbytes: N*M bytes, site-major.  Writes a .pbwt with the reference's pbwtWrite (pbwtIO.c:33-57). */
#include "pbwt.h"

PBWT *pbwtGpuBuildFromSites (const uchar *x, int M, int N) ;

char *commandLine = "shim_selftest" ;
FILE *logFile ;
const char *pbwtCommitHash (void) { return "shim" ; }

int main (int argc, char **argv)
{
  if (argc != 5) { fprintf (stderr, "usage: %s in.hapbytes M N out.pbwt\n", argv[0]) ; return 2 ; }
  int M = atoi (argv[2]), N = atoi (argv[3]) ;
  logFile = stderr ;
  pbwtInit () ;
  uchar *x = myalloc ((long)M*N + 1, uchar) ;
  FILE *f = fopen (argv[1], "rb") ;
  if (!f || fread (x, 1, (long)M*N, f) != (size_t)((long)M*N)) die ("cannot read %s", argv[1]) ;
  fclose (f) ;
  PBWT *p = pbwtGpuBuildFromSites (x, M, N) ;
  f = fopen (argv[4], "wb") ;
  pbwtWrite (p, f) ;
  fclose (f) ;
  pbwtDestroy (p) ;
  return 0 ;
}
