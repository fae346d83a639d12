/* mock_pbwtgpu.c -- TEST INFRASTRUCTURE ONLY: a stand-in for libpbwtgpu.so on machines without a GPU.
 *
 * It exports the entry points of include/pbwtgpu.h that the host shim (pbwt_b200/host/pbwt_gpu_shim.c) binds,
 * answered by the CPU oracle port (oracle/libpbwt_oracle.so, orc_*).  Linked ONLY into
 * tests/host_mock/_build/pbwt_hostcheck, which the `-m "not gpu"` suite uses to check the shim's HOST logic --
 * callback order, log lines, Array / aFend ownership, the readers' lazy cursor and its batching -- against the
 * reference's golden outputs.  Nothing under pbwt_b200/ links or loads this file; the product library has no CPU path
 * (pbwtgpu_init fails with PBWTGPU_ERR_NODEVICE without a B200), and parity of the device code is established by the
 * `-m gpu` suite through the real library.
 */
#include <stdlib.h>
#include <string.h>

#include "pbwtgpu.h"
#define ORACLE_PREFIX orc_
#include "oracle_api.h"

struct pbwtgpu_ctx { int unused; };
static struct pbwtgpu_ctx theCtx;

int pbwtgpu_words_per_col(int M) { return 4 * ((M + 127) / 128); }
int pbwtgpu_init(int device, pbwtgpu_ctx **ctx) { (void)device; *ctx = &theCtx; return PBWTGPU_OK; }
void pbwtgpu_destroy(pbwtgpu_ctx *ctx) { (void)ctx; }
const char *pbwtgpu_strerror(int code) { return code == PBWTGPU_OK ? "ok" : "mock device error"; }
const char *pbwtgpu_last_error(pbwtgpu_ctx *ctx) { (void)ctx; return "mock device (CPU oracle port)"; }
void pbwtgpu_free(void *p) { free(p); }
void *pbwtgpu_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void pbwtgpu_host_trim(void) {}

int pbwtgpu_build(pbwtgpu_ctx *ctx, const uint32_t *xbits, int M, int N, const int32_t *a_start, unsigned flags,
                  uint8_t **yz, int64_t *nyz, int32_t *a_end, int32_t *d_end, pbwtgpu_match **matches, int64_t *nmatch) {
  (void)ctx;
  if ((flags & ~(unsigned)PBWTGPU_WANT_YZ) || d_end || matches || nmatch) return PBWTGPU_ERR_UNSUPPORTED; /* the shim never asks */
  const int64_t cap = (int64_t)N * (M + 8) + 16;
  uint8_t *buf = (uint8_t *)malloc((size_t)cap);
  int32_t *ae = a_end ? a_end : (int32_t *)malloc(sizeof(int32_t) * (size_t)M);
  int32_t *de = (int32_t *)malloc(sizeof(int32_t) * ((size_t)M + 1));
  if (!buf || !ae || !de) return PBWTGPU_ERR_NOMEM;
  const int rc = orc_build(xbits, pbwtgpu_words_per_col(M), M, N, a_start, 0, buf, cap, nyz, ae, de);
  if (!a_end) free(ae);
  free(de);
  if (rc) { free(buf); return PBWTGPU_ERR_ARG; }
  *yz = buf;
  return PBWTGPU_OK;
}

static int ends(const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *a_start, int32_t *a_end, int32_t *d_end) {
  if (!a_end && !d_end) return 0;
  int32_t *ae = (int32_t *)malloc(sizeof(int32_t) * (size_t)M), *de = (int32_t *)malloc(sizeof(int32_t) * ((size_t)M + 1));
  if (!ae || !de) return PBWTGPU_ERR_NOMEM;
  orc_sweep_ad(yz, nyz, M, N, a_start, ae, de);
  if (a_end) memcpy(a_end, ae, sizeof(int32_t) * (size_t)M);
  if (d_end) memcpy(d_end, de, sizeof(int32_t) * ((size_t)M + 1));
  free(ae); free(de);
  return 0;
}

int pbwtgpu_max_within(pbwtgpu_ctx *ctx, const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *a_start,
                       int32_t *a_end, int32_t *d_end, pbwtgpu_match **matches, int64_t *nmatch) {
  (void)ctx;
  int64_t n = 0;
  orc_max_within(yz, nyz, M, N, a_start, NULL, 0, &n);
  oracle_match *rec = (oracle_match *)malloc(sizeof(oracle_match) * (size_t)(n > 0 ? n : 1));
  if (!rec) return PBWTGPU_ERR_NOMEM;
  orc_max_within(yz, nyz, M, N, a_start, rec, n, &n);
  if (getenv("PBWTGPU_MOCK_CORRUPT")) /* tests: a wrong device result that -check must catch */
    for (int64_t i = 0; i < n; ++i)
      if (rec[i].end - rec[i].start > 1) { rec[i].start -= rec[i].start > 0 ? 1 : -1; break; }
  *matches = (pbwtgpu_match *)rec;
  *nmatch = n;
  return ends(yz, nyz, M, N, a_start, a_end, d_end);
}

/* grouped stream: runs of records that share ai, start, end -> {ai, start, end, n} + n x bi */
int pbwtgpu_max_within_grouped(pbwtgpu_ctx *ctx, const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *a_start, int32_t *a_end,
                               int32_t *d_end, int32_t **groups, int64_t *nwords) {
  pbwtgpu_match *rec = NULL;
  int64_t n = 0, i = 0, w = 0;
  const int rc = pbwtgpu_max_within(ctx, yz, nyz, M, N, a_start, a_end, d_end, &rec, &n);
  if (rc) return rc;
  int32_t *g = (int32_t *)malloc(sizeof(int32_t) * (size_t)(5 * n + 4));
  if (!g) { free(rec); return PBWTGPU_ERR_NOMEM; }
  while (i < n) {
    int64_t j = i;
    while (j < n && rec[j].ai == rec[i].ai && rec[j].start == rec[i].start && rec[j].end == rec[i].end) ++j;
    g[w] = rec[i].ai; g[w + 1] = rec[i].start; g[w + 2] = rec[i].end; g[w + 3] = (int32_t)(j - i);
    for (int64_t t = i; t < j; ++t) g[w + 4 + (t - i)] = rec[t].bi;
    w += 4 + (j - i);
    i = j;
  }
  free(rec);
  *groups = g;
  *nwords = w;
  return PBWTGPU_OK;
}

/* the device consumer of the matches behind -stats: one count per reporter (a run of records with the same ai and end)
   at the length of the longer of its blocks, derived here from the oracle's records */
int pbwtgpu_max_within_stats(pbwtgpu_ctx *ctx, const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *a_start, int64_t *hist) {
  (void)ctx;
  int64_t n = 0, i = 0;
  orc_max_within(yz, nyz, M, N, a_start, NULL, 0, &n);
  oracle_match *rec = (oracle_match *)malloc(sizeof(oracle_match) * (size_t)(n > 0 ? n : 1));
  if (!rec) return PBWTGPU_ERR_NOMEM;
  orc_max_within(yz, nyz, M, N, a_start, rec, n, &n);
  memset(hist, 0, sizeof(int64_t) * ((size_t)N + 1));
  while (i < n) {
    const int ai = rec[i].ai, end = rec[i].end;
    int best = 0;
    for (; i < n && rec[i].ai == ai && rec[i].end == end; ++i)
      if (rec[i].end - rec[i].start > best) best = rec[i].end - rec[i].start;
    ++hist[best];
  }
  free(rec);
  return PBWTGPU_OK;
}

int pbwtgpu_long_within(pbwtgpu_ctx *ctx, const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *a_start, int L,
                        int32_t *a_end, int32_t *d_end, pbwtgpu_match **matches, int64_t *nmatch) {
  (void)ctx;
  if (L <= 0) return PBWTGPU_ERR_ARG;
  int64_t n = 0;
  orc_long_within(yz, nyz, M, N, a_start, L, NULL, 0, &n);
  oracle_match *rec = (oracle_match *)malloc(sizeof(oracle_match) * (size_t)(n > 0 ? n : 1));
  if (!rec) return PBWTGPU_ERR_NOMEM;
  orc_long_within(yz, nyz, M, N, a_start, L, rec, n, &n);
  *matches = (pbwtgpu_match *)rec;
  *nmatch = n;
  return ends(yz, nyz, M, N, a_start, a_end, d_end);
}

int pbwtgpu_build_reverse(pbwtgpu_ctx *ctx, const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *a_start,
                          int32_t *a_fend, uint8_t **zz, int64_t *nzz, int32_t *a_rend) {
  (void)ctx;
  const int64_t cap = (int64_t)M * (N > 0 ? N : 1) + 16;
  uint8_t *buf = (uint8_t *)malloc((size_t)cap);
  int32_t *af = a_fend ? a_fend : (int32_t *)malloc(sizeof(int32_t) * (size_t)M);
  if (!buf || !af) return PBWTGPU_ERR_NOMEM;
  const int rc = orc_build_reverse(yz, nyz, M, N, a_start, af, buf, cap, nzz, a_rend);
  if (!a_fend) free(af);
  if (rc) { free(buf); return PBWTGPU_ERR_ARG; }
  *zz = buf;
  return PBWTGPU_OK;
}

int pbwtgpu_match_sweep(pbwtgpu_ctx *ctx, const uint8_t *pyz, int64_t npyz, int M, int N, const int32_t *pa_start,
                        const uint8_t *qyz, int64_t nqyz, int Q, const int32_t *qa_start, pbwtgpu_match **matches,
                        int64_t *nmatch, pbwtgpu_nomatch **nomatch, int64_t *nnomatch) {
  (void)ctx;
  const char *qmax = getenv("PBWTGPU_MOCK_QMAX"); /* tests: refuse more queries per call, as the device does beyond 8,192 */
  if (qmax && Q > atoi(qmax)) return PBWTGPU_ERR_UNSUPPORTED;
  int64_t n = 0, nn = 0;
  orc_match_sweep(pyz, npyz, M, N, pa_start, qyz, nqyz, Q, qa_start, NULL, 0, &n, NULL, 0, &nn);
  oracle_match *rec = (oracle_match *)malloc(sizeof(oracle_match) * (size_t)(n > 0 ? n : 1));
  oracle_nomatch *nom = (oracle_nomatch *)malloc(sizeof(oracle_nomatch) * (size_t)(nn > 0 ? nn : 1));
  if (!rec || !nom) return PBWTGPU_ERR_NOMEM;
  orc_match_sweep(pyz, npyz, M, N, pa_start, qyz, nqyz, Q, qa_start, rec, n, &n, nom, nn, &nn);
  *matches = (pbwtgpu_match *)rec;
  *nmatch = n;
  *nomatch = (pbwtgpu_nomatch *)nom;
  *nnomatch = nn;
  return PBWTGPU_OK;
}
