/* htslib_stub.c -- TEST INFRASTRUCTURE ONLY.
 * htslib (the reference's un-vendored third-party dependency, Makefile:2-5) is not installed in
 * this image, so the four symbols pbwtHtslib.c would provide (pbwt.h:29,201-204) are stubbed.
 * None of the hot-path arithmetic lives behind them; they only parse VCF/BCF text. */
#include "pbwt.h"

const char *pbwtHtslibVersionString(void) { return "stub"; }
PBWT *pbwtReadVcfGT(char *filename) { die("pbwtReadVcfGT(%s): htslib not available in this build", filename); return 0; }
PBWT *pbwtReadVcfPL(char *filename) { die("pbwtReadVcfPL(%s): htslib not available in this build", filename); return 0; }
void pbwtWriteVcf(PBWT *p, char *filename, char *reference_fname, char *mode)
{ (void)p; (void)reference_fname; (void)mode; die("pbwtWriteVcf(%s): htslib not available in this build", filename); }
