/* pbwtgpu.h -- C ABI of the B200 (sm_100a) PBWT construction and match engine.
 *
 * Plain C: opaque handles, plain pointers and sizes, int return codes (0 = ok, <0 = error,
 * never exit()).  This is the surface a maintainer of the reference binds to replace its hot
 * path (INTEGRATION.md shows the host shim that does so); each entry point names the reference
 * interface it stands in for (file:line into richarddurbin/pbwt).
 *
 * Panel format ("bit columns"): site-major; column k starts at word k*wpc of a uint32 array,
 * bit (h & 31) of word (h >> 5) = allele of haplotype h at site k, wpc = pbwtgpu_words_per_col(M)
 * = 4*ceil(M/128) so that every column is 16-byte aligned (a TMA bulk-copy unit).
 * This is what the readers' per-site x[] (pbwtIO.c:572-573, pbwtHtslib.c:134) becomes when a
 * batch of sites is handed over at once.
 *
 * Match records are exactly the arguments of the reference's report(ai,bi,start,end) callback
 * (pbwt.h:208,213), in callback order, zero-length records included.
 */
#ifndef PBWTGPU_H
#define PBWTGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pbwtgpu_ctx pbwtgpu_ctx;       /* one per GPU (device memory, stream, scratch) */
typedef struct pbwtgpu_panel pbwtgpu_panel;   /* bit columns resident in HBM */
typedef struct pbwtgpu_result pbwtgpu_result; /* device-resident outputs of one sweep */

typedef struct { int32_t ai, bi, start, end; } pbwtgpu_match;
typedef struct { int32_t query, allele, site; } pbwtgpu_nomatch; /* pbwtMatch.c:406-407 log line */

enum {
  PBWTGPU_OK = 0,
  PBWTGPU_ERR_CUDA = -1,      /* a CUDA call failed; see pbwtgpu_last_error */
  PBWTGPU_ERR_ARG = -2,       /* bad argument */
  PBWTGPU_ERR_NOMEM = -3,     /* host or device allocation failed */
  PBWTGPU_ERR_UNSUPPORTED = -4, /* shape outside what the kernels handle */
  PBWTGPU_ERR_NODEVICE = -5   /* no CUDA device / not sm_100: there is NO CPU fallback */
};

/* what a sweep should produce */
enum {
  PBWTGPU_WANT_YZ = 1u,        /* pack3 stream of the sorted columns  (pbwtCore.c:240-277) */
  PBWTGPU_WANT_D = 2u,         /* carry the divergence array           (pbwtCore.c:485-508) */
  PBWTGPU_WANT_MAXWITHIN = 4u, /* set-maximal matches, implies WANT_D  (pbwtMatch.c:115-142) */
  PBWTGPU_WANT_LONGWITHIN = 8u,/* all matches of >= L sites (-longWithin L, matchLongWithin2, pbwtMatch.c:85-113),
                                  implies WANT_D; L is the option "long_within"; exclusive with WANT_MAXWITHIN */
  PBWTGPU_WANT_STATS = 16u     /* with WANT_MAXWITHIN: the match records are CONSUMED on the device into the length histogram
                                  of -stats (pbwtMatch.c:130-131) and never materialised; pbwtgpu_result_fetch_hist */
};

int pbwtgpu_words_per_col(int M);

/* ---- context -------------------------------------------------------------------------------- */
int pbwtgpu_init(int device, pbwtgpu_ctx **ctx);
void pbwtgpu_destroy(pbwtgpu_ctx *ctx);
const char *pbwtgpu_strerror(int code);
const char *pbwtgpu_last_error(pbwtgpu_ctx *ctx);
/* Run on a caller-owned CUDA stream (cudaStream_t passed as void*); NULL restores the context's
 * own (non-blocking) stream.  Lets a harness bracket the kernels with events on the launching stream.
 * NB the legacy default stream's handle IS NULL: a caller that orders this library's work against another
 * library's (an NCCL collective writing the aggregates of pbwtgpu_shard_finish, say) must create an explicit
 * stream and use it for both. */
int pbwtgpu_set_stream(pbwtgpu_ctx *ctx, void *cuda_stream);
void *pbwtgpu_get_stream(pbwtgpu_ctx *ctx);
int pbwtgpu_synchronize(pbwtgpu_ctx *ctx);
/* Options (0 = automatic / off).  Tuning: "windows" (site windows swept concurrently), "match_capacity" (records the
 * first attempt allows for; by default 4,096 per haplotype or 1.25 x the densest output the context has produced),
 * "long_within" (L of PBWTGPU_WANT_LONGWITHIN), "blocking_sync" (host waits sleep on an event instead of spinning: hosts
 * with more waiting threads than cores), "edge_first_pct" / "edge_second_pct" / "edge_last_pct" (sites of a panel's first
 * two and last window in per cent of an inner window's; built in: 80 / 92 / 74).  Testing and profiling aids: "force_global"
 * (wide-panel kernels on narrow panels), "no_cluster" (cooperative-grid groups where clusters would fit),
 * "no_sort_pass1" (window nodes by sweeping instead of sorting), "no_sort_combine" (window combines by sweeping
 * pseudo-sites instead of the radix sort + range maxima), "no_fused_pack3" (sorted columns to global memory + the
 * stand-alone pack3 encoder instead of the encoder fused into the sweep), "pack3_window_capacity" (bytes per window of the
 * fused encoder's output: small values force its fall-back), "equal_windows" (no shortening of the panel's edge
 * windows), "window_times", "debug_no_candidates". */
int pbwtgpu_set_option(pbwtgpu_ctx *ctx, const char *name, long long value);
/* Per-kernel-class device times of the last sweep (ms, CUDA events on the launching stream) and
 * launch counts; arrays of PBWTGPU_NUM_KERNEL_CLASSES. */
enum { PBWTGPU_K_SWEEP = 0,    /* the output-producing sweep (pass 2 of the window-parallel scheme) */
       PBWTGPU_K_PACK3 = 1, PBWTGPU_K_UNPACK3 = 2, PBWTGPU_K_QUERY = 3, PBWTGPU_K_SYNTH = 4,
       PBWTGPU_K_AUX = 5,      /* tie groups, key bits, record gather */
       PBWTGPU_K_PREPASS = 6,  /* sweep kernel, pass 1: every window on its own */
       PBWTGPU_K_COMBINE = 7,  /* sweep kernel over pseudo-sites: window-prefix combine rounds */
       PBWTGPU_K_COPY = 8,     /* device-to-device copies of window / shard states (no kernels of ours: launch count 0) */
       PBWTGPU_NUM_KERNEL_CLASSES = 9 };
int pbwtgpu_profile(pbwtgpu_ctx *ctx, int enable);
int pbwtgpu_last_timings(pbwtgpu_ctx *ctx, float *ms, int *launches);

/* ---- panels --------------------------------------------------------------------------------- */
int pbwtgpu_panel_from_host(pbwtgpu_ctx *ctx, const uint32_t *xbits, int M, int N, pbwtgpu_panel **panel);
/* Counter-based synthetic mosaic panel generated on the device (bit-identical to
 * oracle/pbwt_oracle.c:orc_generate). flip_thresh = P(flip) * 2^32. */
int pbwtgpu_panel_synthetic(pbwtgpu_ctx *ctx, int M, int N, uint64_t seed, uint32_t L, uint32_t flip_thresh,
                            pbwtgpu_panel **panel);
/* Columns first_site .. first_site+N-1 of the same synthetic panel (a site shard for one GPU). */
int pbwtgpu_panel_synthetic_range(pbwtgpu_ctx *ctx, int M, int N, int first_site, uint64_t seed, uint32_t L,
                                  uint32_t flip_thresh, pbwtgpu_panel **panel);
int pbwtgpu_panel_to_host(pbwtgpu_ctx *ctx, const pbwtgpu_panel *panel, int first_site, int nsites, uint32_t *xbits);
void pbwtgpu_panel_free(pbwtgpu_ctx *ctx, pbwtgpu_panel *panel);

/* ---- device-resident sweeps ----------------------------------------------------------------- */
/* Build from bit columns: replaces the readers' per-site loop
 *   for j: u->y[j] = x[u->a[j]]; pbwtCursorWriteForwards(u)      (pbwtIO.c:572-574, pbwtCore.c:573-585)
 * over all N sites, and (WANT_MAXWITHIN) the matchMaximalWithin sweep over the same states
 * (pbwtMatch.c:115-142), in one pass.  a_start (host, may be NULL = identity) is p->aFstart. */
int pbwtgpu_build_device(pbwtgpu_ctx *ctx, const pbwtgpu_panel *panel, const int32_t *a_start, unsigned flags,
                         pbwtgpu_result **result);
/* Sweep a stored pack3 stream (host buffer as read by pbwtRead, pbwtIO.c:172-217): device unpack3
 * (pbwtCore.c:279-305) + pbwtCursorForwardsReadAD over all sites (pbwtCore.c:543-557)
 * [+ matchMaximalWithin].  WANT_YZ is ignored. */
int pbwtgpu_sweep_yz_device(pbwtgpu_ctx *ctx, const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *a_start,
                            unsigned flags, pbwtgpu_result **result);
int pbwtgpu_result_sizes(const pbwtgpu_result *result, int64_t *nyz, int64_t *nmatch);
/* D2H of whatever is non-NULL: yz[nyz], a_end[M] (= p->aFend, pbwtCore.c:587-591), d_end[M+1],
 * matches[nmatch]. */
int pbwtgpu_result_fetch(pbwtgpu_ctx *ctx, const pbwtgpu_result *result, uint8_t *yz, int32_t *a_end, int32_t *d_end,
                         pbwtgpu_match *matches);
/* The match records in GROUPED form: the reference's reporters emit runs of callbacks that share ai, start and end (the block
 * above a position, the block below it, pbwtMatch.c:133-134; the block of a query, :385,433-435) and only bi changes inside a
 * run.  The grouped stream is int32 words: {ai, start, end, n} followed by the n values of bi, group after group, in
 * callback order; expanding it -- report(ai, bi[j], start, end) for j < n -- reproduces the record sequence exactly.  About
 * a third of the bytes of the records (config 2: 95 MB instead of 295 MB), which is what crosses PCIe on the one-shot paths
 * the host shim uses.  pbwtgpu_result_group builds the stream on the device (once per result; *nwords = its length,
 * *ngroups = number of groups), pbwtgpu_result_fetch_groups copies it to dst[nwords]. */
typedef struct { int32_t ai, start, end, n; } pbwtgpu_group; /* followed by n x int32_t bi */
int pbwtgpu_result_group(pbwtgpu_ctx *ctx, pbwtgpu_result *result, int64_t *nwords, int64_t *ngroups);
int pbwtgpu_result_fetch_groups(pbwtgpu_ctx *ctx, const pbwtgpu_result *result, int32_t *dst);
/* WANT_STATS: hist[len], len = 0 .. N, = number of (position, site) reporters of matchMaximalWithin whose longer block
 * has that length -- exactly matchLengthHist of the reference (pbwtMatch.c:130-131,161-178). */
int pbwtgpu_result_fetch_hist(pbwtgpu_ctx *ctx, const pbwtgpu_result *result, int64_t *hist);
void pbwtgpu_result_free(pbwtgpu_ctx *ctx, pbwtgpu_result *result);

/* ---- multi-GPU: the SITES of one panel are partitioned across GPUs (one process per GPU) ---------------
 * Each GPU holds the columns [site_offset, site_offset + panel->N) of a panel of total_sites sites.
 *   1. pbwtgpu_shard_prepare: sweeps the local windows on their own and reduces them to the shard's aggregate
 *      node (its sort order and divergence relative to the panel's start order, with tie groups).
 *   2. the caller all-gathers the aggregates of all shards, in site order, into one DEVICE buffer
 *      (pbwtgpu_shard_aggregate_ints(M) int32 per shard; bench.py uses torch.distributed/NCCL over NVLink);
 *      this is the path's only exchange step.
 *   3. pbwtgpu_shard_finish: combines the aggregates of the earlier shards into the exact (a, d) entering this
 *      shard and sweeps the local windows from their exact start states.  Site labels in d and in match
 *      records are global.  a_end / d_end are only produced by the shard that ends at total_sites
 *      (others are filled with -1); pack3 streams and match records concatenate across shards in site order.
 * dev_dst / dev_aggs are device pointers on ctx's GPU. */
typedef struct pbwtgpu_shard pbwtgpu_shard;
size_t pbwtgpu_shard_aggregate_ints(int M);
int pbwtgpu_shard_prepare(pbwtgpu_ctx *ctx, const pbwtgpu_panel *panel, const int32_t *a_start, int site_offset,
                          int total_sites, unsigned flags, pbwtgpu_shard **shard);
int pbwtgpu_shard_export(pbwtgpu_ctx *ctx, pbwtgpu_shard *shard, int32_t *dev_dst);
int pbwtgpu_shard_finish(pbwtgpu_ctx *ctx, pbwtgpu_shard *shard, const int32_t *dev_aggs, int rank, unsigned flags,
                         pbwtgpu_result **result);
void pbwtgpu_shard_free(pbwtgpu_ctx *ctx, pbwtgpu_shard *shard);

/* ---- multi-GPU matchSequencesSweep (pbwtMatch.c:363-443): the SITES are partitioned across GPUs ---------------------
 * Every GPU gets the whole stored streams of panel and queries (they are small) and sweeps the columns
 * [first_site, first_site + nsites) with ALL queries; what crosses GPUs is
 *   1. one all-gather of the query stream's shard aggregates (pbwtgpu_shard_aggregate_ints(Q) int32 each): gives every
 *      shard the query PBWT's order entering its windows (the callback order, pbwtMatch.c:377-378);
 *   2. one all-gather of the panel's shard aggregates (pbwtgpu_shard_aggregate_ints(M) int32 each): the exact (a, d)
 *      entering every shard, as in pbwtgpu_shard_*;
 *   3. the queries' state (f, e) -- 2 Q int32 -- handed from each shard to the next, in site order: the one sequential
 *      step, a few hundred microseconds per shard (pbwt_b200/csrc/query_chain.cu).
 * Call order on every rank: open, export_queries, [all-gather 1], prepare, [all-gather 2], chain (rank r after it
 * received rank r-1's state_out; rank 0 passes NULL), [send state_out on], finish.  Records and "no match" events of the
 * shards concatenate in site order into exactly the reference's callback / log order; the matches still open at N are
 * reported by the last shard.  All dev_* pointers are device memory on ctx's GPU. */
typedef struct pbwtgpu_qshard pbwtgpu_qshard;
int pbwtgpu_qshard_open(pbwtgpu_ctx *ctx, const uint8_t *pyz, int64_t npyz, int M, int N, const int32_t *pa_start,
                        const uint8_t *qyz, int64_t nqyz, int Q, const int32_t *qa_start, int first_site, int nsites,
                        pbwtgpu_qshard **shard);
int pbwtgpu_qshard_export_queries(pbwtgpu_ctx *ctx, pbwtgpu_qshard *shard, int32_t *dev_dst);
int pbwtgpu_qshard_prepare(pbwtgpu_ctx *ctx, pbwtgpu_qshard *shard, const int32_t *dev_query_aggs, int rank,
                           int32_t *dev_panel_agg);
int pbwtgpu_qshard_chain(pbwtgpu_ctx *ctx, pbwtgpu_qshard *shard, const int32_t *dev_panel_aggs, int rank,
                         const int32_t *dev_state_in, int32_t *dev_state_out);
int pbwtgpu_qshard_finish(pbwtgpu_ctx *ctx, pbwtgpu_qshard *shard, pbwtgpu_match **matches, int64_t *nmatch,
                          pbwtgpu_nomatch **nomatch, int64_t *nnomatch);
void pbwtgpu_qshard_free(pbwtgpu_ctx *ctx, pbwtgpu_qshard *shard);

/* ---- host buffers ---------------------------------------------------------------------------------- */
/* Page-locked host memory from a process-wide, reusing pool: a caller that assembles its bit columns in such a
 * buffer (the host shim's reader batch, pbwtIO.c:558-586 batched) gets the H2D copy at the full PCIe rate instead
 * of the pageable one.  NULL when the allocation fails.  Give it back with pbwtgpu_free. */
void *pbwtgpu_host_alloc(size_t bytes);
/* Frees any buffer this library handed out (*yz, *zz, *matches, *nomatch, pbwtgpu_host_alloc): those are page-locked
 * blocks of the same pool (kept for reuse), so they MUST come back through here, not through free(). */
void pbwtgpu_free(void *p);
/* Returns the pool's cached (free) page-locked blocks to the system. */
void pbwtgpu_host_trim(void);

/* ---- one-shot host-buffer entry points (what the host shim calls) --------------------------- */
/* H2D of the columns, sweep, D2H.  *yz and *matches are allocated by the library (give them back with
 * pbwtgpu_free); a_end / d_end are caller-allocated (int[M], int[M+1]) or NULL. */
int pbwtgpu_build(pbwtgpu_ctx *ctx, const uint32_t *xbits, int M, int N, const int32_t *a_start, unsigned flags,
                  uint8_t **yz, int64_t *nyz, int32_t *a_end, int32_t *d_end, pbwtgpu_match **matches,
                  int64_t *nmatch);
/* pbwtgpu_build / pbwtgpu_max_within with the matches as a grouped stream (*groups, *nwords int32; give it back with
 * pbwtgpu_free): what the host shim calls -- it expands the groups while it walks them for report(). */
int pbwtgpu_build_grouped(pbwtgpu_ctx *ctx, const uint32_t *xbits, int M, int N, const int32_t *a_start, unsigned flags,
                          uint8_t **yz, int64_t *nyz, int32_t *a_end, int32_t *d_end, int32_t **groups, int64_t *nwords);
int pbwtgpu_max_within_grouped(pbwtgpu_ctx *ctx, const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *a_start,
                               int32_t *a_end, int32_t *d_end, int32_t **groups, int64_t *nwords);
/* matchMaximalWithin(p, report) over a stored stream (pbwtMatch.c:115-142). */
int pbwtgpu_max_within(pbwtgpu_ctx *ctx, const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *a_start,
                       int32_t *a_end, int32_t *d_end, pbwtgpu_match **matches, int64_t *nmatch);
/* pbwtLongMatches(p, 0) under -stats (pbwtMatch.c:148-178): the length histogram alone, hist[0 .. N]; only these
 * (N + 1) counters cross PCIe, not the match records. */
int pbwtgpu_max_within_stats(pbwtgpu_ctx *ctx, const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *a_start,
                             int64_t *hist);
/* pbwtLongMatches(p, L) with L > 0 = matchLongWithin2 (pbwtMatch.c:85-113, 148-183) over a stored stream: every
 * pair of haplotypes whose match ending at a site is at least L sites long and cannot be extended.  Records in the
 * reference's callback order, including its quirks (the last block of a site is never reported; at k == N the
 * alleles of site N-1 are consulted). */
int pbwtgpu_long_within(pbwtgpu_ctx *ctx, const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *a_start, int L,
                        int32_t *a_end, int32_t *d_end, pbwtgpu_match **matches, int64_t *nmatch);
/* Compatibility path for callers that walk a PbwtCursor themselves (pbwt.h:74-87): the cursor's state ENTERING site k
 * of a stored stream, i.e. after k x pbwtCursorForwardsReadAD (pbwtCore.c:543-557): a[M], d[M+1] (sentinels k+1) and,
 * for k < N, the sorted column y[M] with its zero count *c (u->y, u->c).  One device sweep per call: meant for jumping
 * to a site, not for stepping.  Any of y, a, d, c may be NULL. */
int pbwtgpu_cursor_at(pbwtgpu_ctx *ctx, const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *a_start, int k,
                      uint8_t *y, int32_t *a, int32_t *d, int32_t *c);
/* pbwtBuildReverse(p) (pbwtCore.c:151-191) over a stored stream: sweeps forwards to aFend (= aRstart), un-sorts every
 * column on the device, and builds the reverse stream zz (malloc()'d) and aRend from the sites in reverse order.
 * a_fend may be NULL. */
int pbwtgpu_build_reverse(pbwtgpu_ctx *ctx, const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *a_start,
                          int32_t *a_fend, uint8_t **zz, int64_t *nzz, int32_t *a_rend);
/* matchSequencesSweep(p, q, report) (pbwtMatch.c:363-443): panel and query streams as stored.  At most 8,192 queries per
 * call (PBWTGPU_ERR_UNSUPPORTED beyond); the host shim chunks larger query sets and merges the record streams exactly.
 * Wide panels are swept window-parallel: a clamped pass per window, the exact query state at every window boundary
 * (pbwt_b200/csrc/query_chain.cu), then all windows at once. */
int pbwtgpu_match_sweep(pbwtgpu_ctx *ctx, const uint8_t *pyz, int64_t npyz, int M, int N, const int32_t *pa_start,
                        const uint8_t *qyz, int64_t nqyz, int Q, const int32_t *qa_start, pbwtgpu_match **matches,
                        int64_t *nmatch, pbwtgpu_nomatch **nomatch, int64_t *nnomatch);
/* pack3 / unpack3 of whole streams on the device (pbwtCore.c:254-305): y is N columns of M bytes. */
int pbwtgpu_pack3(pbwtgpu_ctx *ctx, const uint8_t *y, int M, int N, uint8_t **yz, int64_t *nyz);
int pbwtgpu_unpack3(pbwtgpu_ctx *ctx, const uint8_t *yz, int64_t nyz, int M, int N, uint8_t *y, int32_t *zero_counts);

#ifdef __cplusplus
}
#endif
#endif /* PBWTGPU_H */
