/* oracle_api.h -- TEST INFRASTRUCTURE ONLY (never linked into the product path).
 *
 * One flat C interface that two different CPU implementations export:
 *
 *   ref_*  (oracle/_ref/libpbwtref.so)   = the UNMODIFIED reference sources compiled in
 *           place from $PBWT_REF (default /root/reference) plus oracle/ref_driver.c,
 *           which only marshals buffers in and out of the reference's own functions.
 *   orc_*  (oracle/libpbwt_oracle.so)    = oracle/pbwt_oracle.c, a from-scratch CPU
 *           restatement of the same algorithms (kind "port"), pinned against ref_* and
 *           against the golden vectors in tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load either library.
 *
 * Panel input format (shared with the GPU library, include/pbwtgpu.h): site-major bit
 * columns.  Column k starts at word k*wpc of a uint32 array; bit (h & 31) of word
 * (h >> 5) is the allele of haplotype h at site k.  wpc = 4*ceil(M/128) so that every
 * column is 16-byte aligned.
 *
 * Match records are {ai, bi, start, end} exactly as handed to the reference's
 * report(ai,bi,start,end) callback (pbwtMatch.c:115-142, :363-443), in callback order,
 * zero-length records included.
 */
#ifndef PBWT_ORACLE_API_H
#define PBWT_ORACLE_API_H

#include <stdint.h>

#ifndef ORACLE_PREFIX
#error "define ORACLE_PREFIX to ref_ or orc_ before including oracle_api.h"
#endif
#define ORC_CAT2(a, b) a##b
#define ORC_CAT(a, b) ORC_CAT2(a, b)
#define ORC(name) ORC_CAT(ORACLE_PREFIX, name)

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int32_t ai, bi, start, end; } oracle_match;
typedef struct { int32_t query, allele, site; } oracle_nomatch;

/* Build: per site gather y[j]=x[a[j]] then pbwtCursorWriteForwards (with_d=0) or
 * pbwtCursorWriteForwardsAD(u,k) (with_d=1).  pbwtIO.c:558-586, pbwtCore.c:573-585.
 * yz_out must hold yz_cap bytes; *nyz gets the byte count (returns -1 if yz_cap too small).
 * aEnd int[M]; dEnd int[M+1] (only written when with_d). aStart may be NULL (identity). */
int ORC(build)(const uint32_t *xbits, int wpc, int M, int N, const int32_t *aStart, int with_d,
               uint8_t *yz_out, int64_t yz_cap, int64_t *nyz, int32_t *aEnd, int32_t *dEnd);

/* Per-site trace of the AD build (state ENTERING site k, k = 0..N):
 * a_trace int[(N+1)*M], d_trace int[(N+1)*(M+1)], y_trace uchar[N*M], c_trace int[N]. */
int ORC(trace)(const uint32_t *xbits, int wpc, int M, int N, const int32_t *aStart,
               int32_t *a_trace, int32_t *d_trace, uint8_t *y_trace, int32_t *c_trace);

/* matchMaximalWithin over a stored pack3 stream. pbwtMatch.c:115-142.
 * Records beyond cap are counted but not stored. */
int ORC(max_within)(const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *aStart,
                    oracle_match *rec, int64_t cap, int64_t *nrec);

/* pbwtLongMatches(p, T) with T > 0 = matchLongWithin2 over a stored pack3 stream. pbwtMatch.c:85-113. */
int ORC(long_within)(const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *aStart, int T,
                     oracle_match *rec, int64_t cap, int64_t *nrec);

/* pbwtBuildReverse (pbwtCore.c:151-191) on a stored stream: aFend int[M] (= aRstart), the reverse stream zz
 * (returns -1 if zz_cap is too small; *nzz is the byte count either way) and aRend int[M]. */
int ORC(build_reverse)(const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *aStart,
                       int32_t *aFend, uint8_t *zz_out, int64_t zz_cap, int64_t *nzz, int32_t *aRend);

/* N x pbwtCursorForwardsReadAD from a fresh cursor: final a int[M], d int[M+1].
 * pbwtCore.c:543-557. */
int ORC(sweep_ad)(const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *aStart,
                  int32_t *aEnd, int32_t *dEnd);

/* N x pbwtCursorForwardsRead (run-wise ForwardsAPacked path): final a. pbwtCore.c:527-541,595-619. */
int ORC(sweep_a)(const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *aStart, int32_t *aEnd);

/* matchSequencesSweep(panel, query). pbwtMatch.c:363-443. "no match" log lines are returned
 * as triples in nomatch[]. */
int ORC(match_sweep)(const uint8_t *pyz, int64_t npyz, int M, int N, const int32_t *paStart,
                     const uint8_t *qyz, int64_t nqyz, int Q, const int32_t *qaStart,
                     oracle_match *rec, int64_t cap, int64_t *nrec,
                     oracle_nomatch *nomatch, int64_t nmcap, int64_t *nnomatch);

/* pack3 / unpack3 of one column. pbwtCore.c:254-305.  y must have M+1 bytes (sentinel slot). */
int ORC(pack3)(uint8_t *y, int M, uint8_t *out);
int ORC(unpack3)(const uint8_t *yz, int M, uint8_t *y, int32_t *n0);

/* CPU-baseline timers (seconds, CLOCK_MONOTONIC, one thread).  Work timed:
 *   time_build:      gather + pbwtCursorWriteForwards over all N sites;
 *   time_max_within: matchMaximalWithin with a counting callback over the yz just built.
 * Returns 0; *nmatch = number of callback invocations. */
int ORC(time_build_maxwithin)(const uint32_t *xbits, int wpc, int M, int N,
                              double *t_build, double *t_match, int64_t *nyz, int64_t *nmatch);

/* The same two timers with order-sensitive checksums of what was produced, for bench.py's parity gate (the GPU's
 * outputs for the same sites must hash to the same values before any timing counts).  For items i = 0.. with mix m_i:
 * S = sum m_i, T = sum (i+1) m_i (mod 2^64); a pack3 byte b mixes to b * 0x9E3779B97F4A7C15 + 1, a record to
 * oracle_mix_record().  ck = {S_yz, T_yz, S_rec, T_rec}.  time_build times the build alone (with_d selects
 * pbwtCursorWriteForwardsAD); ck = {S_yz, T_yz}. */
int ORC(time_build_maxwithin_ck)(const uint32_t *xbits, int wpc, int M, int N, double *t_build, double *t_match,
                                 int64_t *nyz, int64_t *nmatch, uint64_t *ck);
int ORC(time_build)(const uint32_t *xbits, int wpc, int M, int N, int with_d, double *t_build, int64_t *nyz, uint64_t *ck);

static inline uint64_t oracle_mix_record(int32_t ai, int32_t bi, int32_t start, int32_t end) {
  const uint64_t w0 = (uint64_t)(uint32_t)ai | ((uint64_t)(uint32_t)bi << 32);
  const uint64_t w1 = (uint64_t)(uint32_t)start | ((uint64_t)(uint32_t)end << 32);
  uint64_t m = (w0 * 0x9E3779B97F4A7C15ull) ^ (w1 * 0xC2B2AE3D27D4EB4Full);
  m ^= m >> 29;
  return m | 1u;
}
static inline void oracle_ck_bytes(const uint8_t *b, int64_t n, uint64_t *ck) {
  uint64_t S = 0, T = 0;
  for (int64_t i = 0; i < n; ++i) { const uint64_t m = (uint64_t)b[i] * 0x9E3779B97F4A7C15ull + 1u; S += m; T += (uint64_t)(i + 1) * m; }
  ck[0] = S; ck[1] = T;
}

#ifdef __cplusplus
}
#endif
#endif
