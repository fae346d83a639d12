/* ref_driver.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Marshals flat buffers in and out of the UNMODIFIED reference (compiled in place from
 * $PBWT_REF by oracle/Makefile into oracle/_ref/libpbwtref.so).  No algorithm lives here:
 * every function below calls the reference's own API in the same sequence the reference's
 * readers and CLI commands do (file:line given per function).
 */
#define _GNU_SOURCE
#include "pbwt.h" /* from $PBWT_REF via -I, never copied into this repo */
#include <time.h>

#define ORACLE_PREFIX ref_
#include "oracle_api.h"

/* globals normally defined in pbwtMain.c (pbwt.h:28,91-92) */
char *commandLine = "oracle/ref_driver";
FILE *logFile = 0;
const char *pbwtCommitHash(void) { return "oracle"; }

static int initialised = 0;
static void ensure_init(void) {
  if (!initialised) {
    pbwtInit(); /* pbwtCore.c:34 -- sets up the pack3 decode LUT */
    initialised = 1;
  }
  if (!logFile) logFile = fopen("/dev/null", "w");
}

static inline int xbit(const uint32_t *col, int h) { return (col[h >> 5] >> (h & 31)) & 1; }

/* PBWT object around a stored stream, laid out as pbwtRead does (pbwtIO.c:172-217) */
static PBWT *pbwt_from_yz(const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *aStart) {
  PBWT *p = pbwtCreate(M, N);
  if (aStart) memcpy(p->aFstart, aStart, M * sizeof(int));
  p->yz = arrayCreate(nyz > 0 ? nyz : 1, uchar);
  if (nyz > 0) {
    array(p->yz, nyz - 1, uchar) = 0;
    memcpy(arrp(p->yz, 0, uchar), yz, nyz);
  }
  return p;
}

int ref_build(const uint32_t *xbits, int wpc, int M, int N, const int32_t *aStart, int with_d,
              uint8_t *yz_out, int64_t yz_cap, int64_t *nyz, int32_t *aEnd, int32_t *dEnd) {
  ensure_init();
  PBWT *p = pbwtCreate(M, 0);
  if (aStart) memcpy(p->aFstart, aStart, M * sizeof(int));
  PbwtCursor *u = pbwtCursorCreate(p, TRUE, TRUE);
  for (int k = 0; k < N; ++k) { /* the readers' per-site loop, pbwtIO.c:572-574 */
    const uint32_t *col = xbits + (size_t)k * wpc;
    for (int j = 0; j < M; ++j) u->y[j] = xbit(col, u->a[j]);
    if (with_d) pbwtCursorWriteForwardsAD(u, k);
    else pbwtCursorWriteForwards(u);
    ++p->N;
  }
  pbwtCursorToAFend(u, p);
  int rc = 0;
  *nyz = arrayMax(p->yz);
  if (*nyz > yz_cap) rc = -1;
  else if (*nyz) memcpy(yz_out, arrp(p->yz, 0, uchar), *nyz);
  memcpy(aEnd, u->a, M * sizeof(int));
  if (with_d && dEnd) memcpy(dEnd, u->d, (M + 1) * sizeof(int));
  pbwtCursorDestroy(u);
  pbwtDestroy(p);
  return rc;
}

int ref_trace(const uint32_t *xbits, int wpc, int M, int N, const int32_t *aStart,
              int32_t *a_trace, int32_t *d_trace, uint8_t *y_trace, int32_t *c_trace) {
  ensure_init();
  PBWT *p = pbwtCreate(M, 0);
  if (aStart) memcpy(p->aFstart, aStart, M * sizeof(int));
  PbwtCursor *u = pbwtCursorCreate(p, TRUE, TRUE);
  for (int k = 0; k < N; ++k) {
    const uint32_t *col = xbits + (size_t)k * wpc;
    memcpy(a_trace + (size_t)k * M, u->a, M * sizeof(int));
    memcpy(d_trace + (size_t)k * (M + 1), u->d, (M + 1) * sizeof(int));
    int c = 0;
    for (int j = 0; j < M; ++j) { u->y[j] = xbit(col, u->a[j]); c += !u->y[j]; }
    memcpy(y_trace + (size_t)k * M, u->y, M);
    c_trace[k] = c;
    pbwtCursorWriteForwardsAD(u, k);
    ++p->N;
  }
  memcpy(a_trace + (size_t)N * M, u->a, M * sizeof(int));
  memcpy(d_trace + (size_t)N * (M + 1), u->d, (M + 1) * sizeof(int));
  pbwtCursorDestroy(u);
  pbwtDestroy(p);
  return 0;
}

/* recording callback handed to the reference's match functions */
static oracle_match *g_rec;
static int64_t g_cap, g_n;
static void record_cb(int ai, int bi, int start, int end) {
  if (g_n < g_cap) { g_rec[g_n].ai = ai; g_rec[g_n].bi = bi; g_rec[g_n].start = start; g_rec[g_n].end = end; }
  ++g_n;
}
static void count_cb(int ai, int bi, int start, int end) { (void)ai; (void)bi; (void)start; (void)end; ++g_n; }

int ref_max_within(const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *aStart,
                   oracle_match *rec, int64_t cap, int64_t *nrec) {
  ensure_init();
  PBWT *p = pbwt_from_yz(yz, nyz, M, N, aStart);
  g_rec = rec; g_cap = cap; g_n = 0;
  matchMaximalWithin(p, record_cb); /* pbwtMatch.c:115 */
  *nrec = g_n;
  pbwtDestroy(p);
  return 0;
}

/* matchLongWithin2 is static in the reference and only reachable through pbwtLongMatches (pbwtMatch.c:148-183),
 * whose printer writes MATCH lines to stdout (pbwtMatch.c:46-49): capture stdout and parse the lines back. */
int ref_long_within(const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *aStart, int T,
                    oracle_match *rec, int64_t cap, int64_t *nrec) {
  ensure_init();
  PBWT *p = pbwt_from_yz(yz, nyz, M, N, aStart);
  char *buf = 0; size_t len = 0;
  fflush(stdout);
  FILE *saved = stdout;
  stdout = open_memstream(&buf, &len); /* glibc: stdout is an assignable FILE* */
  pbwtLongMatches(p, T);
  fclose(stdout);
  stdout = saved;
  int64_t n = 0;
  for (char *line = buf; line && *line; ) {
    int ai, bi, st, en, ln;
    if (sscanf(line, "MATCH\t%d\t%d\t%d\t%d\t%d", &ai, &bi, &st, &en, &ln) == 5) {
      if (n < cap && rec) { rec[n].ai = ai; rec[n].bi = bi; rec[n].start = st; rec[n].end = en; }
      ++n;
    }
    line = strchr(line, '\n');
    if (line) ++line;
  }
  *nrec = n;
  free(buf);
  pbwtDestroy(p);
  return 0;
}

int ref_build_reverse(const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *aStart,
                      int32_t *aFend, uint8_t *zz_out, int64_t zz_cap, int64_t *nzz, int32_t *aRend) {
  ensure_init();
  PBWT *p = pbwt_from_yz(yz, nyz, M, N, aStart);
  pbwtBuildReverse(p); /* pbwtCore.c:151 */
  int rc = 0;
  *nzz = arrayMax(p->zz);
  if (*nzz > zz_cap) rc = -1;
  else if (*nzz) memcpy(zz_out, arrp(p->zz, 0, uchar), *nzz);
  memcpy(aFend, p->aFend, M * sizeof(int));
  memcpy(aRend, p->aRend, M * sizeof(int));
  pbwtDestroy(p);
  return rc;
}

int ref_sweep_ad(const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *aStart,
                 int32_t *aEnd, int32_t *dEnd) {
  ensure_init();
  PBWT *p = pbwt_from_yz(yz, nyz, M, N, aStart);
  PbwtCursor *u = pbwtCursorCreate(p, TRUE, TRUE);
  for (int k = 0; k < N; ++k) pbwtCursorForwardsReadAD(u, k); /* pbwtCore.c:543 */
  memcpy(aEnd, u->a, M * sizeof(int));
  memcpy(dEnd, u->d, (M + 1) * sizeof(int));
  pbwtCursorDestroy(u);
  pbwtDestroy(p);
  return 0;
}

int ref_sweep_a(const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *aStart, int32_t *aEnd) {
  ensure_init();
  PBWT *p = pbwt_from_yz(yz, nyz, M, N, aStart);
  PbwtCursor *u = pbwtCursorCreate(p, TRUE, TRUE);
  for (int k = 0; k < N; ++k) pbwtCursorForwardsRead(u); /* pbwtCore.c:527 */
  memcpy(aEnd, u->a, M * sizeof(int));
  pbwtCursorDestroy(u);
  pbwtDestroy(p);
  return 0;
}

int ref_match_sweep(const uint8_t *pyz, int64_t npyz, int M, int N, const int32_t *paStart,
                    const uint8_t *qyz, int64_t nqyz, int Q, const int32_t *qaStart,
                    oracle_match *rec, int64_t cap, int64_t *nrec,
                    oracle_nomatch *nomatch, int64_t nmcap, int64_t *nnomatch) {
  ensure_init();
  PBWT *p = pbwt_from_yz(pyz, npyz, M, N, paStart);
  PBWT *q = pbwt_from_yz(qyz, nqyz, Q, N, qaStart);
  /* the reference reports "no match" only as text on logFile (pbwtMatch.c:406-407) */
  char *logbuf = 0; size_t loglen = 0;
  FILE *saved = logFile;
  logFile = open_memstream(&logbuf, &loglen);
  g_rec = rec; g_cap = cap; g_n = 0;
  matchSequencesSweep(p, q, record_cb); /* pbwtMatch.c:363 */
  fclose(logFile);
  logFile = saved;
  *nrec = g_n;
  int64_t nn = 0;
  for (char *line = logbuf; line && *line; ) {
    int jj, x, k;
    if (sscanf(line, "no match to query %d value %d at site %d", &jj, &x, &k) == 3) {
      if (nn < nmcap) { nomatch[nn].query = jj; nomatch[nn].allele = x; nomatch[nn].site = k; }
      ++nn;
    }
    line = strchr(line, '\n');
    if (line) ++line;
  }
  *nnomatch = nn;
  free(logbuf);
  pbwtDestroy(p);
  pbwtDestroy(q);
  return 0;
}

int ref_pack3(uint8_t *y, int M, uint8_t *out) {
  ensure_init();
  y[M] = Y_SENTINEL;
  return pack3(y, M, out); /* pbwtCore.c:254 */
}

int ref_unpack3(const uint8_t *yz, int M, uint8_t *y, int32_t *n0) {
  ensure_init();
  return unpack3((uchar *)yz, M, y, n0); /* pbwtCore.c:279 */
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int ref_time_build_maxwithin(const uint32_t *xbits, int wpc, int M, int N,
                             double *t_build, double *t_match, int64_t *nyz, int64_t *nmatch) {
  ensure_init();
  /* unpack the bit columns to the byte-per-allele x[] the reference's readers hold, outside
   * the timed region (parsing excluded on both sides) */
  uint8_t *x = malloc((size_t)M + 1);
  PBWT *p = pbwtCreate(M, 0);
  PbwtCursor *u = pbwtCursorCreate(p, TRUE, TRUE);
  double tb = 0;
  for (int k = 0; k < N; ++k) {
    const uint32_t *col = xbits + (size_t)k * wpc;
    for (int h = 0; h < M; ++h) x[h] = xbit(col, h);
    x[M] = Y_SENTINEL;
    double t0 = now_s();
    for (int j = 0; j < M; ++j) u->y[j] = x[u->a[j]];
    pbwtCursorWriteForwards(u);
    ++p->N;
    tb += now_s() - t0;
  }
  pbwtCursorToAFend(u, p);
  pbwtCursorDestroy(u);
  *t_build = tb;
  *nyz = arrayMax(p->yz);
  g_n = 0;
  double t0 = now_s();
  matchMaximalWithin(p, count_cb);
  *t_match = now_s() - t0;
  *nmatch = g_n;
  pbwtDestroy(p);
  free(x);
  return 0;
}

/* ---- the same timers with checksums of the outputs (bench.py's parity gate; oracle_api.h) ---- */
static uint64_t g_S, g_T;
static void hash_cb(int ai, int bi, int start, int end) {
  const uint64_t m = oracle_mix_record(ai, bi, start, end);
  g_S += m;
  g_T += (uint64_t)(++g_n) * m;
}

static PBWT *timed_build(const uint32_t *xbits, int wpc, int M, int N, int with_d, double *t_build) {
  uint8_t *x = malloc((size_t)M + 1);
  PBWT *p = pbwtCreate(M, 0);
  PbwtCursor *u = pbwtCursorCreate(p, TRUE, TRUE);
  double tb = 0;
  for (int k = 0; k < N; ++k) { /* the readers' per-site loop, pbwtIO.c:572-574 */
    const uint32_t *col = xbits + (size_t)k * wpc;
    for (int h = 0; h < M; ++h) x[h] = xbit(col, h);
    x[M] = Y_SENTINEL;
    double t0 = now_s();
    for (int j = 0; j < M; ++j) u->y[j] = x[u->a[j]];
    if (with_d) pbwtCursorWriteForwardsAD(u, k);
    else pbwtCursorWriteForwards(u);
    ++p->N;
    tb += now_s() - t0;
  }
  pbwtCursorToAFend(u, p);
  pbwtCursorDestroy(u);
  free(x);
  *t_build = tb;
  return p;
}

int ref_time_build(const uint32_t *xbits, int wpc, int M, int N, int with_d, double *t_build, int64_t *nyz, uint64_t *ck) {
  ensure_init();
  PBWT *p = timed_build(xbits, wpc, M, N, with_d, t_build);
  *nyz = arrayMax(p->yz);
  oracle_ck_bytes(arrp(p->yz, 0, uchar), *nyz, ck);
  pbwtDestroy(p);
  return 0;
}

int ref_time_build_maxwithin_ck(const uint32_t *xbits, int wpc, int M, int N, double *t_build, double *t_match,
                                int64_t *nyz, int64_t *nmatch, uint64_t *ck) {
  ensure_init();
  PBWT *p = timed_build(xbits, wpc, M, N, 0, t_build);
  *nyz = arrayMax(p->yz);
  oracle_ck_bytes(arrp(p->yz, 0, uchar), *nyz, ck);
  g_n = 0; g_S = g_T = 0;
  double t0 = now_s();
  matchMaximalWithin(p, hash_cb); /* pbwtMatch.c:115 */
  *t_match = now_s() - t0;
  *nmatch = g_n;
  ck[2] = g_S; ck[3] = g_T;
  pbwtDestroy(p);
  return 0;
}
