// pbwtgpu.cu -- C ABI (include/pbwtgpu.h) over the sm_100a kernels.  No CPU fallback anywhere:
// every entry point either runs the CUDA kernels or returns an error code.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <climits>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "group_records.cuh"
#include "pack3.cuh"
#include "query_chain.cuh"
#include "sweep.cuh"
#include "synth.cuh"
#include "windows.cuh"

using namespace pbwtgpu;

struct pbwtgpu_ctx {
  int device = 0;
  cudaStream_t ownStream = nullptr;
  cudaStream_t stream = nullptr;
  std::string lastError;
  // options
  long long optWindows = 0;    // 0 = automatic
  long long optMatchCap = 0;   // 0 = automatic
  long long optBlockingSync = 0;  // host waits sleep on an event instead of spinning (hosts with more waiting threads than cores)
  cudaEvent_t syncEvent = nullptr;
  double recPerHapSite = 0.0;  // densest match output this context has produced (records per haplotype-site): a panel that
                               // once overflowed the first attempt's pool does not pay for a second pass 2 on every later call
  long long optWindowTimes = 0;  // debugging: print per-window CTA times of pass 2 to stderr
  long long optNoCandidates = 0;  // profiling aid
  long long optEdgePct[3] = {0, 0, 0};  // tuning aid: sites of windows 0 / 1 / last in % of an inner window's (0 = built-in)
  long long optForceGlobal = 0;  // testing: use the global-memory kernel even for narrow panels
  long long optLongT = 0;        // L of PBWTGPU_WANT_LONGWITHIN
  long long optEqualWindows = 0; // testing / profiling: equal windows even where the edge windows would be shortened
  long long optNoSortPass1 = 0;  // testing: pass 1 by sweeping even where the window sort applies
  long long optNoCluster = 0;    // testing: wide panels through the cooperative-grid path even where clusters fit
  long long optNoSortCombine = 0;  // testing: window combines by sweeping pseudo-sites even where the sort applies
  long long optNoFusedPack3 = 0;   // testing: sorted columns to global memory + the stand-alone encoder even where the sweep encodes
  long long optPack3Cap = 0;       // testing: bytes per window of the fused encoder's output (0 = automatic)
  // profiling
  bool profile = false;
  float ms[PBWTGPU_NUM_KERNEL_CLASSES] = {0};
  int launches[PBWTGPU_NUM_KERNEL_CLASSES] = {0};
  struct Pending { int cls; cudaEvent_t a, b; };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> eventPool;
  // freed device blocks kept for reuse (cudaMalloc is far slower than a sweep at small panels)
  struct Block { void *p; size_t bytes; };
  std::vector<Block> freeBlocks;
  std::unordered_map<void *, size_t> liveBlocks;  // true size of every block handed out (a cached block may be larger than asked for)
  size_t cacheCapBytes = (size_t)48 << 30;        // freed blocks kept for reuse: at most this much (60 % of the device's memory)
  int smCount = 0;
  // pinned host scratch for the few counters read back per sweep (one stream synchronisation instead of one per copy)
  unsigned long long *pinned = nullptr;
  static constexpr int kPinnedWords = 1024;
};

// Wait for everything enqueued on the context's stream.  By default the calling thread spins (lowest latency); with the
// option blocking_sync it sleeps on an event created with cudaEventBlockingSync, which matters when a host runs more
// waiting threads than it has cores (eight ranks with three pipeline threads each on a 16-core box).
static cudaError_t stream_sync(pbwtgpu_ctx *ctx) {
  if (!ctx->optBlockingSync) return cudaStreamSynchronize(ctx->stream);
  if (!ctx->syncEvent) {
    cudaError_t e = cudaEventCreateWithFlags(&ctx->syncEvent, cudaEventBlockingSync | cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  cudaError_t e = cudaEventRecord(ctx->syncEvent, ctx->stream);
  return e != cudaSuccess ? e : cudaEventSynchronize(ctx->syncEvent);
}

struct pbwtgpu_panel {
  uint32_t *bits = nullptr;
  int M = 0, N = 0, wpc = 0;
};

struct pbwtgpu_result {
  int M = 0, N = 0;
  uint8_t *yz = nullptr;
  int64_t nyz = 0;
  int32_t *aEnd = nullptr;
  int32_t *dEnd = nullptr;
  pbwtgpu_match *rec = nullptr;
  int64_t nrec = 0;
  unsigned int *hist = nullptr;  // WANT_STATS: [N + 1]
  int32_t *groups = nullptr;     // the records in grouped form (pbwtgpu_result_group), 4 * nGroups + nrec int32
  int64_t nGroups = -1;          // -1: not grouped yet
  std::vector<void *> owned;  // device blocks released by pbwtgpu_result_free
  std::vector<size_t> ownedBytes;
};

namespace {

int fail(pbwtgpu_ctx *ctx, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->lastError = buf;
  return code;
}

#define CUDA_TRY(ctx, expr)                                                                        \
  do {                                                                                             \
    cudaError_t e_ = (cudaError_t)(expr);                                                          \
    if (e_ != cudaSuccess)                                                                         \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? PBWTGPU_ERR_NOMEM : PBWTGPU_ERR_CUDA,     \
                  "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_));           \
  } while (0)

int dev_alloc(pbwtgpu_ctx *ctx, void **out, size_t bytes) {
  if (bytes == 0) bytes = 16;
  bytes = (bytes + 255) & ~(size_t)255;
  int best = -1;
  for (int i = 0; i < (int)ctx->freeBlocks.size(); ++i) {
    const size_t b = ctx->freeBlocks[i].bytes;
    if (b >= bytes && b <= 2 * bytes + (1u << 20) && (best < 0 || b < ctx->freeBlocks[best].bytes)) best = i;
  }
  if (best >= 0) {
    *out = ctx->freeBlocks[best].p;
    ctx->liveBlocks[*out] = ctx->freeBlocks[best].bytes;
    ctx->freeBlocks.erase(ctx->freeBlocks.begin() + best);
    return 0;
  }
  cudaError_t e = cudaMalloc(out, bytes);
  if (e != cudaSuccess) {  // drop the cache and retry once
    for (auto &b : ctx->freeBlocks) cudaFree(b.p);
    ctx->freeBlocks.clear();
    e = cudaMalloc(out, bytes);
  }
  if (e != cudaSuccess) return fail(ctx, PBWTGPU_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  ctx->liveBlocks[*out] = bytes;
  return 0;
}

// back to the context's cache under the block's TRUE size (the size asked for is not what is recorded: a reused block
// may be up to twice as large, and the cache's byte cap must count what is really held)
void dev_release(pbwtgpu_ctx *ctx, void *p, size_t /*bytesAskedFor*/) {
  if (!p) return;
  size_t bytes = 0;
  auto it = ctx->liveBlocks.find(p);
  if (it != ctx->liveBlocks.end()) { bytes = it->second; ctx->liveBlocks.erase(it); }
  size_t cached = 0;
  for (auto &b : ctx->freeBlocks) cached += b.bytes;
  // room is made by giving up the blocks that were freed longest ago (reuse takes blocks from anywhere, new ones are
  // appended): a cache filled with an earlier workload's blocks must not make the current one malloc every step
  // (config 4 after config 3 in one process: 70 ms of cudaMalloc / cudaFree per build)
  while (bytes && !ctx->freeBlocks.empty() && (ctx->freeBlocks.size() >= 64 || cached + bytes >= ctx->cacheCapBytes)) {
    cudaFree(ctx->freeBlocks.front().p);
    cached -= ctx->freeBlocks.front().bytes;
    ctx->freeBlocks.erase(ctx->freeBlocks.begin());
  }
  if (bytes && cached + bytes < ctx->cacheCapBytes) ctx->freeBlocks.push_back({p, bytes});
  else cudaFree(p);
}

// ---- pinned host memory handed across the C ABI (pbwtgpu_host_alloc, malloc-style outputs) ----
// One process-wide registry: pbwtgpu_free() takes no context.  Freed blocks are kept for reuse (cudaMallocHost of a
// few hundred MB costs far more than a sweep); pbwtgpu_host_trim() gives them back.
std::mutex gHostMu;
std::unordered_map<void *, size_t> gHostLive;
std::vector<std::pair<void *, size_t>> gHostFree;
constexpr size_t kHostCacheBytes = (size_t)8 << 30;

void *host_alloc_pinned(size_t bytes) {
  if (bytes == 0) bytes = 16;
  bytes = (bytes + 4095) & ~(size_t)4095;
  {
    std::lock_guard<std::mutex> lk(gHostMu);
    int best = -1;
    for (int i = 0; i < (int)gHostFree.size(); ++i) {
      const size_t b = gHostFree[i].second;
      if (b >= bytes && b <= 2 * bytes + (1u << 20) && (best < 0 || b < gHostFree[best].second)) best = i;
    }
    if (best >= 0) {
      void *p = gHostFree[best].first;
      gHostLive[p] = gHostFree[best].second;
      gHostFree.erase(gHostFree.begin() + best);
      return p;
    }
  }
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    std::lock_guard<std::mutex> lk(gHostMu);  // give the cache back and retry once
    for (auto &b : gHostFree) cudaFreeHost(b.first);
    gHostFree.clear();
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    gHostLive[p] = bytes;
    return p;
  }
  std::lock_guard<std::mutex> lk(gHostMu);
  gHostLive[p] = bytes;
  return p;
}

// true when p came from host_alloc_pinned (and has now been released)
bool host_release_pinned(void *p) {
  std::lock_guard<std::mutex> lk(gHostMu);
  auto it = gHostLive.find(p);
  if (it == gHostLive.end()) return false;
  const size_t bytes = it->second;
  gHostLive.erase(it);
  size_t cached = 0;
  for (auto &b : gHostFree) cached += b.second;
  if (gHostFree.size() < 16 && cached + bytes <= kHostCacheBytes) gHostFree.push_back({p, bytes});
  else cudaFreeHost(p);
  return true;
}

// releases a host buffer handed out by this library unless ownership was passed on (release())
struct HostGuard {
  void *p = nullptr;
  explicit HostGuard(void *q) : p(q) {}
  ~HostGuard() { if (p && !host_release_pinned(p)) free(p); }
  void *release() { void *q = p; p = nullptr; return q; }
};

// scoped temporary device allocations, returned to the context cache on scope exit
struct Temps {
  pbwtgpu_ctx *ctx;
  std::vector<std::pair<void *, size_t>> blocks;
  explicit Temps(pbwtgpu_ctx *c) : ctx(c) {}
  ~Temps() {
    for (auto &b : blocks) dev_release(ctx, b.first, b.second);
  }
  template <typename T>
  int get(T **out, size_t count) {
    void *p = nullptr;
    const int rc = dev_alloc(ctx, &p, count * sizeof(T));
    if (rc) return rc;
    blocks.push_back({p, count * sizeof(T)});
    *out = static_cast<T *>(p);
    return 0;
  }
};

template <typename T>
int result_alloc(pbwtgpu_ctx *ctx, pbwtgpu_result *r, T **out, size_t count) {
  void *p = nullptr;
  const int rc = dev_alloc(ctx, &p, count * sizeof(T));
  if (rc) return rc;
  r->owned.push_back(p);
  r->ownedBytes.push_back(count * sizeof(T));
  *out = static_cast<T *>(p);
  return 0;
}

struct ScopedTimer {  // brackets a kernel class with events on the launching stream
  pbwtgpu_ctx *ctx;
  int cls;
  cudaEvent_t a = nullptr, b = nullptr;
  ScopedTimer(pbwtgpu_ctx *c, int k) : ctx(c), cls(k) {
    ctx->launches[cls] += 0;
    if (!ctx->profile) return;
    a = take();
    b = take();
    cudaEventRecord(a, ctx->stream);
  }
  cudaEvent_t take() {
    if (!ctx->eventPool.empty()) {
      cudaEvent_t e = ctx->eventPool.back();
      ctx->eventPool.pop_back();
      return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
  ~ScopedTimer() {
    if (!ctx->profile) return;
    cudaEventRecord(b, ctx->stream);
    ctx->pending.push_back({cls, a, b});
  }
};

void reset_timings(pbwtgpu_ctx *ctx) {
  for (auto &p : ctx->pending) {
    ctx->eventPool.push_back(p.a);
    ctx->eventPool.push_back(p.b);
  }
  ctx->pending.clear();
  memset(ctx->ms, 0, sizeof ctx->ms);
  memset(ctx->launches, 0, sizeof ctx->launches);
}

// ---- the sweep driver shared by build (gather mode), stored-stream sweeps (sorted mode) and shards ----
//
// Windowed execution (gather mode, enough sites): the sites are cut into W windows, one per SM (or per group
// of SMs for wide panels).  prepare(): pass 1 sweeps every window on its own from the panel's start order,
// then a Hillis-Steele scan of window combines (windows.cu) turns those into the exact (a, d) at every window
// boundary.  finish(): pass 2 sweeps all windows concurrently from their exact start states and produces the
// outputs.  A shard of a longer panel (multi-GPU: sites are partitioned across GPUs) additionally exports the
// aggregate node of its whole site range and applies the combined aggregate of all earlier shards before pass 2.
struct WindowedSweep {
  pbwtgpu_ctx *ctx;
  Temps tmp;
  const uint32_t *src = nullptr;
  int wpc = 0, M = 0, N = 0;
  SweepMode mode = kGather;
  int siteOffset = 0, nTotal = 0;
  bool sharded = false;
  bool useGlobal = false;
  int W = 1, G = 1, chunk = 0, spw = 1, nodes = 0, bits = 1, slots = 1, cur = 0;
  int firstSites = 0, secondSites = 0;  // > 0: sites of windows 0 and 1; the last window takes what remains (window_range)
  uint32_t *xOutRev = nullptr;  // sorted mode: un-sorted columns in reverse site order (pbwtgpu_build_reverse)
  GlobalBuffers gb;  // wide panels: state, exchange and match-scan buffers of sweep_global.cu
  bool matchWanted = false;
  // slot 0 = state entering window 0, slot w+1 = node w (after the scan: the state at the end of window w)
  int32_t *ord[2] = {nullptr, nullptr}, *dv[2] = {nullptr, nullptr}, *g[2] = {nullptr, nullptr}, *gfd[2] = {nullptr, nullptr};
  uint32_t *keys = nullptr;
  bool haveDStart = false;
  // stored-stream sweeps combine windows in POSITION space: a window swept from the identity yields the exact
  // permutation of sorted positions over its sites (the sorted columns do not depend on the labels)
  bool positionSpace = false;
  int32_t *iota = nullptr, *ordScratch = nullptr;
  int forceSpw = 0;  // > 0: windows of exactly this many sites (a second sweep that must share another's boundaries)
  bool allowForceGlobal = true;  // honour the testing option force_global
  // ---- window-parallel query sweep (matchSequencesSweep, pbwtMatch.c:363-443); set queryQ / qcols / qwpc before init ----
  // pass 1 sweeps every window from the fresh query state with clamped tiers (SweepArgs::queryMode 1) while it computes
  // the window nodes; the chain (query_chain.cu) turns the clamped end states into the true state entering every
  // window; pass 2 (queryMode 2) runs all windows from their true panel AND query states and emits.
  int queryQ = 0;
  const uint32_t *qcols = nullptr;
  int qwpc = 0;
  const int32_t *qaWin = nullptr;          // [W][Q] query PBWT order entering every window (NULL = identity, W == 1 only)
  int32_t *qNodeOrd = nullptr;             // [nodes][M] the windows' own permutations (position space), for the chain
  int32_t *qf1 = nullptr, *qe1 = nullptr;  // [W][Q] clamped end states (pass 1)
  int32_t *qfT = nullptr, *qeT = nullptr;  // [W][Q] true states entering the windows (chain), pass 2 works in place

  explicit WindowedSweep(pbwtgpu_ctx *c) : ctx(c), tmp(c) {}
  size_t oA() const { return (size_t)M; }
  size_t oD() const { return (size_t)M + 1; }

  int launch(const SweepArgs &args, SweepMode md, bool withD, int withMatch) {
    if (!useGlobal) return launch_sweep_smem(args, md, withD, withMatch, ctx->stream);
    const int wl = std::max(W, 1);
    cudaError_t e = cudaMemsetAsync(gb.partials, 0, sizeof(uint32_t) * (size_t)wl * G * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(gb.gsync, 0, sweep_global_sync_bytes(wl), ctx->stream);
    if (e != cudaSuccess) return (int)e;
    return launch_sweep_global(args, md, withD, withMatch, gb, ctx->stream);
  }

  int init(const uint32_t *src_, int wpc_, int M_, int N_, SweepMode mode_, const int32_t *aStartHost, bool wantMatch,
           bool sharded_, int siteOffset_, int nTotal_) {
    src = src_; wpc = wpc_; M = M_; N = N_; mode = mode_; sharded = sharded_;
    siteOffset = siteOffset_; nTotal = nTotal_ > 0 ? nTotal_ : N_;
    if (M < 1 || N < 0) return fail(ctx, PBWTGPU_ERR_ARG, "bad panel shape M=%d N=%d", M, N);
    if (nTotal >= kPseudoLabelBase - 64) return fail(ctx, PBWTGPU_ERR_UNSUPPORTED, "N=%d too large", nTotal);
    // panels wider than one CTA's shared memory sweep with their state in global memory (sweep_global.cu)
    useGlobal = M > sweep_smem_max_m(true) || (ctx->optForceGlobal && allowForceGlobal) || queryQ > 0;
    matchWanted = wantMatch;

    W = 1;
    positionSpace = mode == kSorted;
    if (M >= 2) {
      const long long want = ctx->optWindows > 0 ? ctx->optWindows : ctx->smCount;
      const int minSites = ctx->optWindows > 0 ? 1 : 256;  // below this a window does not repay its two extra passes
      W = (int)std::max<long long>(1, std::min<long long>(want, N / minSites));
      if (forceSpw > 0) W = std::max(1, (N + forceSpw - 1) / forceSpw);
    }
    if (useGlobal) {
      sweep_global_plan(M, ctx->smCount, W, queryQ ? kSorted : kGather, true, wantMatch, queryQ > 0, queryQ, !ctx->optNoCluster, &G, &chunk,
                        &W, &gb.cluster);
    }
    spw = N > 0 ? (N + W - 1) / W : 1;
    if (forceSpw > 0) spw = forceSpw;
    W = N > 0 ? (N + spw - 1) / spw : 1;
    // A panel swept with matches: its first windows (everything still matches back to site 0) and its last one (every
    // block is reported at site N) carry 10-17 % more work per site than the others (option window_times), so windows
    // 0 / 1 / last get about 80 / 92 / 74 % of the sites of an inner window.  Of a sharded panel only the shard that
    // starts at site 0 shortens its first windows and only the one that ends at N its last.
    // (shared-memory kernel; wide panels keep equal windows)
    if (!useGlobal && wantMatch && W >= 8 && !ctx->optEqualWindows) {
      const bool atStart = siteOffset == 0, atEnd = siteOffset + N == nTotal;
      const double p0 = ctx->optEdgePct[0] > 0 ? ctx->optEdgePct[0] / 100.0 : 0.80, p1 = ctx->optEdgePct[1] > 0 ? ctx->optEdgePct[1] / 100.0 : 0.92,
                   pl = ctx->optEdgePct[2] > 0 ? ctx->optEdgePct[2] / 100.0 : 0.74;
      const double f0 = atStart ? p0 : 1.0, f1 = atStart ? p1 : 1.0, fl = atEnd ? pl : 1.0;
      if (atStart || atEnd) {
        // the last window takes what the others leave: of the few window lengths around N / (W - 3 + f0 + f1 + fl) take
        // the one whose most loaded window (sites over its share) is the lightest -- with the smallest length that fits,
        // a shard that only ends the panel gave its last window 93 % instead of 74 % (0.55 ms of 4.4 on that rank)
        double best = 1e30;
        for (int s = (int)(N / (W - 3 + f0 + f1 + fl)), tries = 0; tries < 4; ++s, ++tries) {
          const int first = std::min(s, (int)(f0 * s + 0.5)), second = std::min(s, (int)(f1 * s + 0.5));
          const long long last = (long long)N - first - second - (long long)(W - 3) * s;
          if (!(s >= 16 && first >= 1 && second >= 1 && last >= 1 && last <= s)) continue;
          const double cost = std::max(std::max((double)s, first / f0), std::max(second / f1, last / fl));
          if (cost < best) { best = cost; spw = s; firstSites = first; secondSites = second; }
        }
      }
    }
    nodes = sharded ? W : W - 1;  // a shard also needs the node of its last window (its aggregate)
    slots = nodes + 1;
    bits = 1;
    while ((1 << bits) < M) ++bits;
    if (useGlobal) {
      const size_t MS = (size_t)G * chunk;
      gb.G = G; gb.chunk = chunk;
      if (int rc = tmp.get(&gb.aBuf, (size_t)W * 2 * MS)) return rc;
      if (int rc = tmp.get(&gb.dBuf, (size_t)W * 2 * (MS + 4))) return rc;
      if (int rc = tmp.get(&gb.partials, (size_t)W * G * 4)) return rc;
      char *sy = nullptr;
      if (int rc = tmp.get(&sy, sweep_global_sync_bytes(W))) return rc;
      gb.gsync = sy;
      if (wantMatch || queryQ) {
        if ((long long)(N + 1) * G + 1 > INT_MAX)
          return fail(ctx, PBWTGPU_ERR_UNSUPPORTED, "N=%d too large for a match sweep over a wide panel (%d CTAs per window)", N, G);
        if (int rc = tmp.get(&gb.ySorted, sweep_global_column_words(G, chunk, W))) return rc;
        if (int rc = tmp.get(&gb.wmaxG, sweep_global_column_words(G, chunk, W))) return rc;
        if (int rc = tmp.get(&gb.segCnt, (size_t)(N + 1) * G + 1)) return rc;
        if (int rc = tmp.get(&gb.segSrc, (size_t)(N + 1) * G + 1)) return rc;
      }
      if (queryQ) {
        const size_t cw = sweep_global_column_words(G, chunk, W);
        if (int rc = tmp.get(&gb.zpre, cw)) return rc;
        if (int rc = tmp.get(&gb.any1, cw / 32)) return rc;
        if (int rc = tmp.get(&gb.any0, cw / 32)) return rc;
        if (int rc = tmp.get(&gb.wmax2, cw / 32)) return rc;
        if (int rc = tmp.get(&qf1, (size_t)W * queryQ)) return rc;
        if (int rc = tmp.get(&qe1, (size_t)W * queryQ)) return rc;
        if (int rc = tmp.get(&qfT, (size_t)W * queryQ)) return rc;
        if (int rc = tmp.get(&qeT, (size_t)W * queryQ)) return rc;
      }
    }
    const bool needNodes = nodes > 0;
    for (int b = 0; b < (needNodes ? 2 : 1); ++b) {
      if (int rc = tmp.get(&ord[b], (size_t)slots * M)) return rc;
      if (needNodes) {
        if (int rc = tmp.get(&dv[b], (size_t)slots * (M + 1))) return rc;
        if (int rc = tmp.get(&g[b], (size_t)slots * M)) return rc;
        if (int rc = tmp.get(&gfd[b], (size_t)slots * M)) return rc;
        CUDA_TRY(ctx, cudaMemsetAsync(dv[b], 0, sizeof(int32_t) * (M + 1), ctx->stream));  // slot 0: fresh cursor
      }
    }
    if (needNodes)
      if (int rc = tmp.get(&keys, (size_t)std::max(nodes, 1) * bits * wpc)) return rc;
    if (needNodes && positionSpace) {
      if (int rc = tmp.get(&iota, (size_t)M)) return rc;
      if (int rc = tmp.get(&ordScratch, (size_t)nodes * M)) return rc;
      CUDA_TRY(ctx, launch_iota(iota, M, ctx->stream));
      if (queryQ)
        if (int rc = tmp.get(&qNodeOrd, (size_t)nodes * M)) return rc;
    }
    {  // slot 0: the panel's start order
      std::vector<int32_t> ident;
      const int32_t *h = aStartHost;
      if (!h) {
        ident.resize(M);
        for (int i = 0; i < M; ++i) ident[i] = i;
        h = ident.data();
      }
      CUDA_TRY(ctx, cudaMemcpyAsync(ord[0], h, sizeof(int32_t) * M, cudaMemcpyHostToDevice, ctx->stream));
      CUDA_TRY(ctx, stream_sync(ctx));  // ident goes out of scope
      if (needNodes) CUDA_TRY(ctx, cudaMemcpyAsync(ord[1], ord[0], sizeof(int32_t) * M, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return 0;
  }

  // out[i] = I[i] (+) J[i] for cnt nodes (I earlier, J later); I is one node for all when sharedI
  int combine(int cnt, const int32_t *ordI, const int32_t *dvI, bool sharedI, const int32_t *gJ, const int32_t *gfdJ,
              int32_t *ordOut, int32_t *dvOut, int32_t *gOut, int32_t *gfdOut) {
    if (cnt <= 0) return 0;
    if (positionSpace) {
      // ids are sorted positions at the I/J boundary: I's end order is the identity there, so the pseudo-site
      // sweep only has to produce the divergence (its order output equals J's own permutation); the node's
      // permutation is then the composition ord_I[ord_J[.]] and its groups are re-derived in I's start ids
      if (int rc = combine_ids(cnt, iota, true, dvI, sharedI, gJ, gfdJ, ordScratch, dvOut, gOut, gfdOut)) return rc;
      ScopedTimer t(ctx, PBWTGPU_K_AUX);
      CUDA_TRY(ctx, launch_compose(ordOut, ordI, sharedI, ordScratch, M, cnt, ctx->stream));
      CUDA_TRY(ctx, launch_make_groups(ordOut, dvOut, nullptr, nullptr, gOut, gfdOut, M, cnt, ctx->stream));
      ctx->launches[PBWTGPU_K_AUX] += 2;
      return 0;
    }
    if (sort_combine()) {
      CombineSortArgs c{};
      c.M = M; c.cnt = cnt; c.sharedI = sharedI;
      c.ordI = ordI; c.dvI = dvI; c.gJ = gJ; c.gfdJ = gfdJ;
      c.ordOut = ordOut; c.dvOut = dvOut; c.gOut = gOut; c.gfdOut = gfdOut;
      ScopedTimer t(ctx, PBWTGPU_K_COMBINE);
      CUDA_TRY(ctx, launch_combine_sort(c, ctx->stream));
      ctx->launches[PBWTGPU_K_COMBINE] += 1;
      return 0;
    }
    return combine_ids(cnt, ordI, sharedI, dvI, sharedI, gJ, gfdJ, ordOut, dvOut, gOut, gfdOut);
  }
  // the combine as a radix sort + range maxima (combine_sort.cu): haplotype space, panels of the shared-memory kernels
  bool sort_combine() const { return !positionSpace && !useGlobal && M <= combine_sort_max_m() && !ctx->optNoSortCombine; }

  int combine_ids(int cnt, const int32_t *ordI, bool ordShared, const int32_t *dvI, bool dvShared, const int32_t *gJ,
                  const int32_t *gfdJ, int32_t *ordOut, int32_t *dvOut, int32_t *gOut, int32_t *gfdOut) {
    {
      ScopedTimer t(ctx, PBWTGPU_K_AUX);
      CUDA_TRY(ctx, launch_key_bits(gJ, M, cnt, bits, wpc, keys, ctx->stream));
      ctx->launches[PBWTGPU_K_AUX] += 1;
    }
    SweepArgs c{};
    c.src = keys; c.wpc = wpc; c.M = M; c.N = N; c.nWindows = cnt;
    c.perWindowSites = bits; c.srcWindowStride = (long long)bits * wpc; c.kLabelBase = kPseudoLabelBase;
    c.aStart = ordI; c.dStart = dvI; c.aStartShared = ordShared; c.dStartShared = dvShared;
    c.aEnd = ordOut; c.dEnd = dvOut;
    {
      ScopedTimer t(ctx, PBWTGPU_K_COMBINE);
      CUDA_TRY(ctx, launch(c, kGather, true, false));
      ctx->launches[PBWTGPU_K_COMBINE] += 1;
    }
    {
      ScopedTimer t(ctx, PBWTGPU_K_AUX);
      CUDA_TRY(ctx, launch_make_groups(ordOut, dvOut, gJ, gfdJ, gOut, gfdOut, M, cnt, ctx->stream));
      ctx->launches[PBWTGPU_K_AUX] += 1;
    }
    return 0;
  }

  void query_args(SweepArgs &a, int mode, int32_t *qf, int32_t *qe) {
    a.qcols = qcols; a.qwpc = qwpc; a.Q = queryQ;
    a.qaStart = qaWin; a.qaPerWindow = qaWin != nullptr;
    a.queryMode = mode;
    gb.qf = qf; gb.qe = qe;
  }
  int launch_query(const SweepArgs &args) {
    const int wl = std::max(W, 1);
    cudaError_t e = cudaMemsetAsync(gb.partials, 0, sizeof(uint32_t) * (size_t)wl * G * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(gb.gsync, 0, sweep_global_sync_bytes(wl), ctx->stream);
    if (e != cudaSuccess) return (int)e;
    return launch_query_sweep_global(args, gb, ctx->stream);
  }

  // true query state entering every window: fresh at the panel's first site (pbwtMatch.c:368-369), then the chain
  // (a shard: stateIn [2][Q] = f then e entering its first site, NULL for the first shard; stateOut [2][Q] leaving its last)
  int query_chain(const int32_t *stateIn = nullptr, int32_t *stateOut = nullptr) {
    QueryChainArgs c{};
    c.Q = queryQ; c.M = M; c.W = W; c.steps = nodes; c.spw = spw; c.firstLabel = siteOffset;
    c.dv = dv[cur]; c.nodeOrd = qNodeOrd; c.fClamped = qf1; c.eClamped = qe1;
    c.fTrue = qfT; c.eTrue = qeT;
    if (stateIn) { c.fIn = stateIn; c.eIn = stateIn + queryQ; }
    if (stateOut) { c.fOut = stateOut; c.eOut = stateOut + queryQ; }
    ScopedTimer t(ctx, PBWTGPU_K_QUERY);
    CUDA_TRY(ctx, launch_query_chain(c, ctx->stream));
    ctx->launches[PBWTGPU_K_QUERY] += 1;
    return 0;
  }

  // pass 1 + tie groups + inclusive scan: afterwards slot w+1 holds the exact state at the end of window w
  // (relative to the start order in slot 0)
  int prepare() {
    if (nodes <= 0) return queryQ ? query_chain() : 0;
    SweepArgs a1{};
    a1.src = src; a1.wpc = wpc; a1.M = M; a1.N = N; a1.sitesPerWindow = spw; a1.firstWindowSites = firstSites; a1.secondWindowSites = secondSites;
    a1.nWindows = nodes;
    a1.siteOffset = siteOffset; a1.nTotal = nTotal;
    a1.aStart = positionSpace ? iota : ord[0]; a1.aStartShared = 1;
    a1.aEnd = ord[0] + oA(); a1.dEnd = dv[0] + oD();
    // Pass 1 only has to deliver each window's node.  Building from panel columns on a panel that fits the
    // shared-memory kernels, the node is computed by a radix sort of the haplotypes' window keys (window_sort.cu)
    // instead of a site-by-site sweep; stored streams and wide panels sweep.
    const bool sortPass1 = !useGlobal && mode == kGather && M <= window_sort_max_m() && !ctx->optNoSortPass1;
    if (sortPass1) {
      uint32_t *T = nullptr;
      if (int rc = tmp.get(&T, window_sort_scratch_words(M, spw, nodes))) return rc;
      ScopedTimer t(ctx, PBWTGPU_K_PREPASS);
      CUDA_TRY(ctx, launch_window_sort(src, wpc, M, N, spw, firstSites, secondSites, siteOffset, a1.aStart, nodes, T, a1.aEnd, a1.dEnd, ctx->stream));
      ctx->launches[PBWTGPU_K_PREPASS] += 2;
    } else if (useGlobal && mode == kGather && !ctx->optNoSortPass1 && !ctx->optNoCluster && window_sort_wide_supported(M, spw)) {
      // wide panels (to ~120k haplotypes): the same sort, one 16-CTA cluster per window (window_sort_wide.cu)
      uint32_t *T = nullptr;
      if (int rc = tmp.get(&T, window_sort_wide_scratch_words(M, spw, nodes))) return rc;
      ScopedTimer t(ctx, PBWTGPU_K_PREPASS);
      CUDA_TRY(ctx, launch_window_sort_wide(src, wpc, M, N, spw, siteOffset, a1.aStart, nodes, T, a1.aEnd, a1.dEnd, ctx->stream));
      ctx->launches[PBWTGPU_K_PREPASS] += 3;
    } else if (useGlobal && mode == kGather && !ctx->optNoSortPass1 && window_sort_global_supported(M, nodes, ctx->smCount)) {
      // wider still: the sort with its elements in global memory, a group of co-resident CTAs per window
      uint32_t *T = nullptr;
      if (int rc = tmp.get(&T, window_sort_global_scratch_words(M, spw, nodes, ctx->smCount))) return rc;
      ScopedTimer t(ctx, PBWTGPU_K_PREPASS);
      CUDA_TRY(ctx, launch_window_sort_global(src, wpc, M, N, spw, siteOffset, a1.aStart, nodes, ctx->smCount, T, a1.aEnd, a1.dEnd,
                                              ctx->stream));
      ctx->launches[PBWTGPU_K_PREPASS] += 3;
    } else if (queryQ) {
      // the query sweep's clamped pass rides on pass 1: in position space the window's own sweep holds the true order
      // (ties keep their entry order) and its fresh divergence is the true one clamped at the window's first site
      query_args(a1, 1, qf1, qe1);
      {
        ScopedTimer t(ctx, PBWTGPU_K_PREPASS);
        CUDA_TRY(ctx, launch_query(a1));
        ctx->launches[PBWTGPU_K_PREPASS] += 1;
      }
      ScopedTimer tc(ctx, PBWTGPU_K_COPY);
      CUDA_TRY(ctx, cudaMemcpyAsync(qNodeOrd, a1.aEnd, sizeof(int32_t) * oA() * nodes, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
      ScopedTimer t(ctx, PBWTGPU_K_PREPASS);
      CUDA_TRY(ctx, launch(a1, mode, true, false));
      ctx->launches[PBWTGPU_K_PREPASS] += 1;
    }
    {
      ScopedTimer t(ctx, PBWTGPU_K_AUX);
      CUDA_TRY(ctx, launch_make_groups(ord[0] + oA(), dv[0] + oD(), nullptr, nullptr, g[0] + oA(), gfd[0] + oA(), M, nodes, ctx->stream));
      ctx->launches[PBWTGPU_K_AUX] += 1;
    }
    cur = 0;
    for (int stride = 1; stride < nodes; stride *= 2) {
      const int nxt = cur ^ 1, cnt = nodes - stride;
      const size_t s1 = (size_t)(1 + stride);  // first slot that changes this round
      if (sort_combine()) {  // one launch per round: the combines, and the nodes below the stride (already final) copied through
        CombineSortArgs c{};
        c.M = M; c.cnt = cnt; c.sharedI = 0;
        c.ordI = ord[cur] + oA(); c.dvI = dv[cur] + oD(); c.gJ = g[cur] + s1 * oA(); c.gfdJ = gfd[cur] + s1 * oA();
        c.ordOut = ord[nxt] + s1 * oA(); c.dvOut = dv[nxt] + s1 * oD(); c.gOut = g[nxt] + s1 * oA(); c.gfdOut = gfd[nxt] + s1 * oA();
        c.ncopy = stride;
        c.ordCopySrc = ord[cur] + oA(); c.dvCopySrc = dv[cur] + oD(); c.gCopySrc = g[cur] + oA(); c.gfdCopySrc = gfd[cur] + oA();
        c.ordCopyDst = ord[nxt] + oA(); c.dvCopyDst = dv[nxt] + oD(); c.gCopyDst = g[nxt] + oA(); c.gfdCopyDst = gfd[nxt] + oA();
        ScopedTimer t(ctx, PBWTGPU_K_COMBINE);
        CUDA_TRY(ctx, launch_combine_sort(c, ctx->stream));
        ctx->launches[PBWTGPU_K_COMBINE] += 1;
        cur = nxt;
        continue;
      }
      if (int rc = combine(cnt, ord[cur] + oA(), dv[cur] + oD(), false, g[cur] + s1 * oA(), gfd[cur] + s1 * oA(),
                           ord[nxt] + s1 * oA(), dv[nxt] + s1 * oD(), g[nxt] + s1 * oA(), gfd[nxt] + s1 * oA()))
        return rc;
      // nodes below the stride are already final
      ScopedTimer tc(ctx, PBWTGPU_K_COPY);
      CUDA_TRY(ctx, cudaMemcpyAsync(ord[nxt] + oA(), ord[cur] + oA(), sizeof(int32_t) * oA() * stride, cudaMemcpyDeviceToDevice, ctx->stream));
      CUDA_TRY(ctx, cudaMemcpyAsync(dv[nxt] + oD(), dv[cur] + oD(), sizeof(int32_t) * oD() * stride, cudaMemcpyDeviceToDevice, ctx->stream));
      CUDA_TRY(ctx, cudaMemcpyAsync(g[nxt] + oA(), g[cur] + oA(), sizeof(int32_t) * oA() * stride, cudaMemcpyDeviceToDevice, ctx->stream));
      CUDA_TRY(ctx, cudaMemcpyAsync(gfd[nxt] + oA(), gfd[cur] + oA(), sizeof(int32_t) * oA() * stride, cudaMemcpyDeviceToDevice, ctx->stream));
      cur = nxt;
    }
    // positions relative to the sweep's start -> haplotypes: compose with the start order (a shard's nodes stay in
    // position space until apply_prefix: the order entering the shard is not known before)
    if (positionSpace && !sharded) {
      const int nxt = cur ^ 1;
      ScopedTimer t(ctx, PBWTGPU_K_AUX);
      CUDA_TRY(ctx, launch_compose(ord[nxt] + oA(), ord[cur], true, ord[cur] + oA(), M, nodes, ctx->stream));
      CUDA_TRY(ctx, cudaMemcpyAsync(ord[cur] + oA(), ord[nxt] + oA(), sizeof(int32_t) * oA() * nodes, cudaMemcpyDeviceToDevice, ctx->stream));
      ctx->launches[PBWTGPU_K_AUX] += 1;
    }
    haveDStart = true;
    if (queryQ && !sharded) return query_chain();  // (a shard's chain also needs the state handed on by the shard before it)
    return 0;
  }

  // aggregate node of the whole shard: ord[M] | dv[M+1] | g[M] | gfd[M]
  static size_t agg_ints(int M) { return 4 * (size_t)M + 1; }
  int export_aggregate(int32_t *dst) {
    if (!sharded || nodes < 1) return fail(ctx, PBWTGPU_ERR_ARG, "export_aggregate: not a shard");
    const size_t s = (size_t)nodes;  // last node's slot
    ScopedTimer tc(ctx, PBWTGPU_K_COPY);
    CUDA_TRY(ctx, cudaMemcpyAsync(dst, ord[cur] + s * oA(), sizeof(int32_t) * M, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(dst + M, dv[cur] + s * oD(), sizeof(int32_t) * (M + 1), cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(dst + 2 * M + 1, g[cur] + s * oA(), sizeof(int32_t) * M, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(dst + 3 * M + 1, gfd[cur] + s * oA(), sizeof(int32_t) * M, cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
  }

  // aggs: [R][agg_ints] aggregates of all shards in site order.  Combines shards 0..rank-1 into the state
  // entering this shard and applies it to every local node and to slot 0.
  int apply_prefix(const int32_t *aggs, int rank) {
    if (rank <= 0) {
      if (positionSpace && sharded && nodes > 0) {  // the first shard: its nodes are relative to the panel's start order
        const int nxt = cur ^ 1;
        ScopedTimer t(ctx, PBWTGPU_K_AUX);
        CUDA_TRY(ctx, launch_compose(ord[nxt] + oA(), ord[cur], true, ord[cur] + oA(), M, nodes, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(ord[cur] + oA(), ord[nxt] + oA(), sizeof(int32_t) * oA() * nodes, cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->launches[PBWTGPU_K_AUX] += 1;
      }
      return 0;
    }
    const size_t AI = agg_ints(M);
    int32_t *P[2];
    if (int rc = tmp.get(&P[0], AI)) return rc;
    if (int rc = tmp.get(&P[1], AI)) return rc;
    int pc = 0;
    if (rank >= 3 && rank - 1 <= std::max(W, 1)) {
      // P = Z_0 (+) ... (+) Z_{rank-1} by an inclusive Hillis-Steele scan over the earlier shards' aggregates
      // (ceil(log2 rank) combine launches instead of rank-1 dependent ones); the last node is the state entering this shard
      const size_t R = (size_t)rank;
      int32_t *sO[2], *sD[2], *sG[2], *sF[2];
      uint32_t *keysP = nullptr, *keysSaved = keys;
      for (int b = 0; b < 2; ++b) {
        if (int rc = tmp.get(&sO[b], R * M)) return rc;
        if (int rc = tmp.get(&sD[b], R * (M + 1))) return rc;
        if (int rc = tmp.get(&sG[b], R * M)) return rc;
        if (int rc = tmp.get(&sF[b], R * M)) return rc;
      }
      if (int rc = tmp.get(&keysP, R * bits * wpc)) return rc;
      const size_t pitch = sizeof(int32_t) * AI;
      CUDA_TRY(ctx, cudaMemcpy2DAsync(sO[0], sizeof(int32_t) * M, aggs, pitch, sizeof(int32_t) * M, R, cudaMemcpyDeviceToDevice, ctx->stream));
      CUDA_TRY(ctx, cudaMemcpy2DAsync(sD[0], sizeof(int32_t) * (M + 1), aggs + M, pitch, sizeof(int32_t) * (M + 1), R, cudaMemcpyDeviceToDevice, ctx->stream));
      CUDA_TRY(ctx, cudaMemcpy2DAsync(sG[0], sizeof(int32_t) * M, aggs + 2 * M + 1, pitch, sizeof(int32_t) * M, R, cudaMemcpyDeviceToDevice, ctx->stream));
      CUDA_TRY(ctx, cudaMemcpy2DAsync(sF[0], sizeof(int32_t) * M, aggs + 3 * M + 1, pitch, sizeof(int32_t) * M, R, cudaMemcpyDeviceToDevice, ctx->stream));
      keys = keysP;
      int sc = 0;
      for (size_t stride = 1; stride < R; stride *= 2) {
        const int sn = sc ^ 1;
        const int cnt = (int)(R - stride);
        const int rc = combine(cnt, sO[sc], sD[sc], false, sG[sc] + stride * M, sF[sc] + stride * M, sO[sn] + stride * M,
                               sD[sn] + stride * (M + 1), sG[sn] + stride * M, sF[sn] + stride * M);
        if (rc) { keys = keysSaved; return rc; }
        CUDA_TRY(ctx, cudaMemcpyAsync(sO[sn], sO[sc], sizeof(int32_t) * M * stride, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(sD[sn], sD[sc], sizeof(int32_t) * (M + 1) * stride, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(sG[sn], sG[sc], sizeof(int32_t) * M * stride, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(sF[sn], sF[sc], sizeof(int32_t) * M * stride, cudaMemcpyDeviceToDevice, ctx->stream));
        sc = sn;
      }
      keys = keysSaved;
      CUDA_TRY(ctx, cudaMemcpyAsync(P[0], sO[sc] + (R - 1) * M, sizeof(int32_t) * M, cudaMemcpyDeviceToDevice, ctx->stream));
      CUDA_TRY(ctx, cudaMemcpyAsync(P[0] + M, sD[sc] + (R - 1) * (M + 1), sizeof(int32_t) * (M + 1), cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
      CUDA_TRY(ctx, cudaMemcpyAsync(P[0], aggs, sizeof(int32_t) * AI, cudaMemcpyDeviceToDevice, ctx->stream));
      for (int r = 1; r < rank; ++r) {  // P = P (+) Z_r
        const int32_t *Z = aggs + (size_t)r * AI;
        if (int rc = combine(1, P[pc], P[pc] + M, true, Z + 2 * M + 1, Z + 3 * M + 1, P[pc ^ 1], P[pc ^ 1] + M,
                             P[pc ^ 1] + 2 * M + 1, P[pc ^ 1] + 3 * M + 1))
          return rc;
        pc ^= 1;
      }
    }
    if (positionSpace) {  // the prefix is a permutation of the panel's start positions: name the haplotypes
      ScopedTimer t(ctx, PBWTGPU_K_AUX);
      CUDA_TRY(ctx, launch_compose(P[pc ^ 1], ord[cur], true, P[pc], M, 1, ctx->stream));
      CUDA_TRY(ctx, cudaMemcpyAsync(P[pc ^ 1] + M, P[pc] + M, sizeof(int32_t) * (M + 1), cudaMemcpyDeviceToDevice, ctx->stream));
      ctx->launches[PBWTGPU_K_AUX] += 1;
      pc ^= 1;
    }
    const int nxt = cur ^ 1;
    if (int rc = combine(nodes, P[pc], P[pc] + M, true, g[cur] + oA(), gfd[cur] + oA(), ord[nxt] + oA(), dv[nxt] + oD(),
                         g[nxt] + oA(), gfd[nxt] + oA()))
      return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(ord[nxt], P[pc], sizeof(int32_t) * M, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(dv[nxt], P[pc] + M, sizeof(int32_t) * (M + 1), cudaMemcpyDeviceToDevice, ctx->stream));
    cur = nxt;
    haveDStart = true;
    return 0;
  }

  // pass 2 (or the only pass): all windows from their exact start states; outputs into res
  int finish(unsigned flags, pbwtgpu_result *res) {
    // -longWithin on an empty panel reports nothing (the cursor's alleles are all zero, pbwtCore.c:406)
    const int matchMode = (flags & PBWTGPU_WANT_LONGWITHIN) ? (N > 0 ? 2 : 0) : (flags & PBWTGPU_WANT_MAXWITHIN) ? 1 : 0;
    const bool wantMatch = matchMode != 0;
    const bool wantD = wantMatch || (flags & (PBWTGPU_WANT_MAXWITHIN | PBWTGPU_WANT_LONGWITHIN)) || (flags & PBWTGPU_WANT_D);
    if (matchMode == 2 && ctx->optLongT < 1) return fail(ctx, PBWTGPU_ERR_ARG, "long_within: set option long_within to L >= 1 first");
    if (wantMatch && useGlobal && !gb.segCnt) return fail(ctx, PBWTGPU_ERR_ARG, "sweep was prepared without match flags");
    const bool wantYz = (flags & PBWTGPU_WANT_YZ) && mode == kGather;
    // -stats: the matches are consumed on the device (length histogram); no record pool, no gather
    const bool statsOnly = (flags & PBWTGPU_WANT_STATS) && matchMode == 1;
    if ((flags & PBWTGPU_WANT_STATS) && matchMode != 1) return fail(ctx, PBWTGPU_ERR_ARG, "WANT_STATS needs WANT_MAXWITHIN");
    res->M = M;
    res->N = N;
    if (statsOnly) {
      if (int rc = result_alloc(ctx, res, &res->hist, (size_t)nTotal + 1)) return rc;
      CUDA_TRY(ctx, cudaMemsetAsync(res->hist, 0, sizeof(unsigned int) * ((size_t)nTotal + 1), ctx->stream));
    }
    if (int rc = result_alloc(ctx, res, &res->aEnd, (size_t)M)) return rc;
    if (wantD)
      if (int rc = result_alloc(ctx, res, &res->dEnd, (size_t)M + 1)) return rc;
    // pack3: run detection fused into the sweep where the shared-memory kernel builds from panel columns (every window
    // appends its columns' run heads to its own region; pack3.cu formats the bytes from those lists), else the sorted
    // columns go to global memory for the encoder over columns (wide panels, -longWithin, and the fallback when a window
    // outgrows its region or a column has more runs than the list pass takes)
    bool fusedP3 = wantYz && N > 0 && !useGlobal && mode == kGather && matchMode != 2 && !xOutRev && !ctx->optNoFusedPack3;
    uint32_t *yOut = nullptr;
    uint32_t *p3Out = nullptr;
    int32_t *p3Count = nullptr;  // [W] run heads per window, [W] = overflow flag
    int32_t *p3SiteOff = nullptr, *p3SiteCnt = nullptr;
    long long p3Cap = 0;  // run heads per window
    if (fusedP3) {
      const long long sitesMax = std::max(spw, std::max(firstSites, secondSites));
      // (the first sites of a panel, before the order has formed, have thousands of runs each; later ones a few dozen)
      p3Cap = ctx->optPack3Cap > 0 ? ctx->optPack3Cap : std::min<long long>(sitesMax * (long long)M, std::max<long long>(256 << 10, sitesMax * 512));
      if (int rc = tmp.get(&p3Out, (size_t)W * (size_t)p3Cap)) return rc;
      if (int rc = tmp.get(&p3Count, (size_t)W + 1)) return rc;
      if (int rc = tmp.get(&p3SiteOff, (size_t)N)) return rc;
      if (int rc = tmp.get(&p3SiteCnt, (size_t)N)) return rc;
      if (W + 8 > pbwtgpu_ctx::kPinnedWords) return fail(ctx, PBWTGPU_ERR_UNSUPPORTED, "too many windows");
    } else if (wantYz && N > 0) {
      if (int rc = tmp.get(&yOut, (size_t)N * wpc)) return rc;
    }

    unsigned long long *recCount = nullptr;
    int *chunkCounter = nullptr, *chunkTab = nullptr;
    pbwtgpu_match *pool = nullptr;
    long long recCap = 0;
    int poolChunks = 0;
    if (wantMatch) {
      if (int rc = tmp.get(&recCount, (size_t)W)) return rc;
      if (int rc = tmp.get(&chunkCounter, 1)) return rc;
      recCap = ctx->optMatchCap > 0 ? ctx->optMatchCap : std::min<long long>(std::max<long long>(1 << 20, 4096LL * M), 1LL << 28);
      if (ctx->optMatchCap <= 0)  // (what this context has seen so far, with a quarter of head room)
        recCap = std::min<long long>(std::max<long long>(recCap, (long long)(1.25 * ctx->recPerHapSite * (double)M * (double)std::max(N, 1))), 1LL << 31);
    }
    if (siteOffset + N != nTotal) {  // not the shard that ends the panel: no final state here
      CUDA_TRY(ctx, cudaMemsetAsync(res->aEnd, 0xFF, sizeof(int32_t) * M, ctx->stream));
      if (res->dEnd) CUDA_TRY(ctx, cudaMemsetAsync(res->dEnd, 0xFF, sizeof(int32_t) * (M + 1), ctx->stream));
    }
    SweepArgs a{};
    a.src = src; a.wpc = wpc; a.M = M; a.N = N;
    a.sitesPerWindow = spw; a.firstWindowSites = firstSites; a.secondWindowSites = secondSites; a.nWindows = W;
    a.siteOffset = siteOffset; a.nTotal = nTotal;
    a.aStart = ord[cur];
    a.dStart = haveDStart ? dv[cur] : nullptr;
    a.aEnd = res->aEnd; a.dEnd = res->dEnd; a.endOnlyLast = 1;
    a.yOut = yOut;
    a.xOut = xOutRev;
    a.recCount = recCount; a.chunkCounter = chunkCounter;
    a.reportAtN = 1;
    a.longT = (int)ctx->optLongT;
    a.debugNoCandidates = (int)ctx->optNoCandidates;
    a.lenHist = res->hist;
    std::vector<unsigned long long> counts(W, 0);
    std::vector<long long> offs;  // per-window record offsets: a host temporary that must outlive the last synchronisation
    long long *winNanos = nullptr;
    if (ctx->optWindowTimes) {
      if (int rc = tmp.get(&winNanos, (size_t)std::max(W, 4))) return rc;
      a.winNanos = winNanos;
    }
    // record streams: one per window (shared-memory kernel) or one per CTA (wide-panel kernel)
    const int producers = useGlobal ? W * G : W;
    const long long nSeg = useGlobal ? (long long)(N + 1) * G + 1 : 0;
    long long *segDst = nullptr;
    void *scanTemp = nullptr;
    size_t scanTempBytes = 0;
    long long segTotal = 0;
    if (wantMatch && useGlobal) {
      if (int rc = tmp.get(&segDst, (size_t)nSeg)) return rc;
      scanTempBytes = segment_scan_temp_bytes(nSeg);
      char *st = nullptr;
      if (int rc = tmp.get(&st, scanTempBytes)) return rc;
      scanTemp = st;
    }
    // pack3 byte counts are enqueued right behind the sweep, so that one synchronisation returns every counter
    int32_t *colBytes = nullptr;
    int64_t *colBase = nullptr;
    if (wantYz) {
      if (int rc = tmp.get(&colBytes, (size_t)N + 1)) return rc;
      if (int rc = tmp.get(&colBase, (size_t)N + 1)) return rc;
    }
    // pinned readback layout: [0] chunks used, [1] pack3 bytes, [2] segment total, [3 ...] per-window record counts
    unsigned long long *pin = ctx->pinned;
    if (W + 3 > pbwtgpu_ctx::kPinnedWords) return fail(ctx, PBWTGPU_ERR_UNSUPPORTED, "too many windows");
    bool poolGrown = false;
    for (int attempt = 0; attempt < 3; ++attempt) {
      a.yOut = yOut;
      a.p3Out = fusedP3 ? p3Out : nullptr; a.p3Cap = p3Cap; a.p3Count = p3Count; a.p3Overflow = fusedP3 ? p3Count + W : nullptr;
      a.p3SiteOff = p3SiteOff; a.p3SiteCnt = p3SiteCnt;
      if (fusedP3) CUDA_TRY(ctx, cudaMemsetAsync(p3Count, 0, sizeof(int32_t) * ((size_t)W + 1), ctx->stream));
      if (wantMatch && !statsOnly) {
        poolChunks = (int)((recCap + kRecChunk - 1) >> kRecChunkLog) + producers;
        if (int rc = tmp.get(&pool, (size_t)poolChunks << kRecChunkLog)) return rc;
        if (int rc = tmp.get(&chunkTab, (size_t)producers * poolChunks)) return rc;
        CUDA_TRY(ctx, cudaMemsetAsync(chunkCounter, 0, sizeof(int), ctx->stream));
        if (useGlobal) CUDA_TRY(ctx, cudaMemsetAsync(gb.segCnt, 0, sizeof(int32_t) * (size_t)nSeg, ctx->stream));
        a.recPool = pool; a.poolChunks = poolChunks; a.chunkTab = chunkTab;
      } else if (statsOnly) {
        CUDA_TRY(ctx, cudaMemsetAsync(chunkCounter, 0, sizeof(int), ctx->stream));
        if (useGlobal) CUDA_TRY(ctx, cudaMemsetAsync(gb.segCnt, 0, sizeof(int32_t) * (size_t)nSeg, ctx->stream));
      }
      {
        ScopedTimer t(ctx, PBWTGPU_K_SWEEP);
        CUDA_TRY(ctx, launch(a, mode, wantD, matchMode));
        ctx->launches[PBWTGPU_K_SWEEP] += 1;
      }
      pin[0] = pin[1] = pin[2] = 0;
      int32_t *pinP3 = reinterpret_cast<int32_t *>(&pin[3 + W]);  // [0] the overflow flag of the run lists
      if (fusedP3) {  // sort every site's run heads, count the bytes; the flag comes back with the other counters
        ScopedTimer t(ctx, PBWTGPU_K_PACK3);
        char *scanScratch = nullptr;
        const size_t scanBytes = p3_scan_scratch_bytes(N);
        if (int rc = tmp.get(&scanScratch, scanBytes)) return rc;
        CUDA_TRY(ctx, p3_count_runs(p3Out, p3Cap, p3SiteOff, p3SiteCnt, M, N, spw, firstSites, secondSites, colBytes, colBase, scanScratch,
                                    scanBytes, p3Count + W, ctx->stream));
        ctx->launches[PBWTGPU_K_PACK3] += 2;
        CUDA_TRY(ctx, cudaMemcpyAsync(pinP3, p3Count + W, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
      } else if (wantYz) {  // (a rerun rewrites the same sorted columns; counting them again is cheap)
        ScopedTimer t(ctx, PBWTGPU_K_PACK3);
        char *scanScratch = nullptr;
        const size_t scanBytes = p3_scan_scratch_bytes(N);
        if (int rc = tmp.get(&scanScratch, scanBytes)) return rc;
        CUDA_TRY(ctx, p3_count_columns(yOut, wpc, M, N, colBytes, colBase, scanScratch, scanBytes, ctx->stream));
        ctx->launches[PBWTGPU_K_PACK3] += (N > 0) + 1;
      }
      if (wantYz) CUDA_TRY(ctx, cudaMemcpyAsync(&pin[1], colBase + N, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
      if (wantMatch) {
        if (useGlobal) {
          ScopedTimer t(ctx, PBWTGPU_K_AUX);
          CUDA_TRY(ctx, launch_segment_scan(gb.segCnt, nSeg, segDst, scanTemp, scanTempBytes, ctx->stream));
          ctx->launches[PBWTGPU_K_AUX] += 1;
          CUDA_TRY(ctx, cudaMemcpyAsync(&pin[2], segDst + nSeg - 1, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        } else {
          CUDA_TRY(ctx, cudaMemcpyAsync(&pin[3], recCount, sizeof(unsigned long long) * W, cudaMemcpyDeviceToHost, ctx->stream));
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(&pin[0], chunkCounter, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      }
      if (wantMatch || wantYz) CUDA_TRY(ctx, stream_sync(ctx));
      bool again = false;
      if (fusedP3 && pinP3[0]) {  // a window outgrew its region (or a column the list pass): sorted columns to global memory instead
        fusedP3 = false;
        if (int rc = tmp.get(&yOut, (size_t)N * wpc)) return rc;
        again = true;
      }
      if (wantMatch && !statsOnly) {
        const int used = (int)(pin[0] & 0xffffffffull);
        segTotal = (long long)pin[2];
        if (!useGlobal)
          for (int w = 0; w < W; ++w) counts[w] = pin[3 + w];
        if (used > poolChunks) {
          if (poolGrown) return fail(ctx, PBWTGPU_ERR_CUDA, "match record pool overflow persisted");
          long long total = segTotal;  // the pool was too small: every count is known now, size it exactly and rerun
          for (auto cnt : counts) total += (long long)cnt;
          recCap = total;
          poolGrown = true;
          again = true;
        }
      }
      if (wantMatch && !statsOnly && N > 0) {
        long long total = segTotal;
        for (auto cnt : counts) total += (long long)cnt;
        ctx->recPerHapSite = std::max(ctx->recPerHapSite, (double)total / ((double)M * (double)N));
      }
      if (!again) break;
      if (attempt == 2) return fail(ctx, PBWTGPU_ERR_CUDA, "sweep output buffers overflowed three times");
    }
    const int64_t yzTotal = (int64_t)pin[1];
    if (winNanos) {
      std::vector<long long> ns(std::max(W, 4));
      CUDA_TRY(ctx, cudaMemcpyAsync(ns.data(), winNanos, sizeof(long long) * ns.size(), cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(ctx, stream_sync(ctx));
      if (useGlobal)
        fprintf(stderr, "[pbwtgpu] wide kernel, CTA 0, cycles per site (%d sites, G=%d chunk=%d W=%d cluster=%d): phase A %.0f, exchange %.0f, scatter+scans %.0f, end barrier %.0f\n",
                spw, G, chunk, W, gb.cluster, (double)ns[0] / spw, (double)ns[1] / spw, (double)ns[2] / spw, (double)ns[3] / spw);
      else {
        fprintf(stderr, "[pbwtgpu] pass-2 window times (us) and records:");
        for (int w = 0; w < W; ++w) fprintf(stderr, " %d:%.0f/%llu", w, ns[w] * 1e-3, counts[w]);
        fprintf(stderr, "\n");
      }
    }
    if (statsOnly) {
      // (the records were counted, not written: nothing to gather)
    } else if (wantMatch && useGlobal) {
      res->nrec = segTotal;
      if (int rc = result_alloc(ctx, res, &res->rec, (size_t)std::max<long long>(segTotal, 1))) return rc;
      ScopedTimer t(ctx, PBWTGPU_K_AUX);
      CUDA_TRY(ctx, launch_gather_segments(pool, chunkTab, poolChunks, gb.segSrc, segDst, nSeg - 1, segTotal, G, spw, N, W,
                                           res->rec, ctx->stream));
      ctx->launches[PBWTGPU_K_AUX] += 1;
    } else if (wantMatch) {
      offs.assign(W, 0);
      long long total = 0;
      int maxChunks = 0;
      for (int w = 0; w < W; ++w) {
        offs[w] = total;
        total += (long long)counts[w];
        maxChunks = std::max<int>(maxChunks, (int)((counts[w] + kRecChunk - 1) >> kRecChunkLog));
      }
      res->nrec = total;
      if (int rc = result_alloc(ctx, res, &res->rec, (size_t)std::max<long long>(total, 1))) return rc;
      long long *offsDev = nullptr;
      if (int rc = tmp.get(&offsDev, (size_t)W)) return rc;
      CUDA_TRY(ctx, cudaMemcpyAsync(offsDev, offs.data(), sizeof(long long) * W, cudaMemcpyHostToDevice, ctx->stream));
      {
        ScopedTimer t(ctx, PBWTGPU_K_AUX);
        CUDA_TRY(ctx, launch_gather_records(pool, chunkTab, poolChunks, offsDev, recCount, W, maxChunks, res->rec, ctx->stream));
        ctx->launches[PBWTGPU_K_AUX] += maxChunks > 0;
      }
    }

    if (wantYz) {
      res->nyz = yzTotal;
      if (int rc = result_alloc(ctx, res, &res->yz, (size_t)yzTotal)) return rc;
      ScopedTimer t(ctx, PBWTGPU_K_PACK3);
      if (fusedP3) {
        CUDA_TRY(ctx, p3_emit_runs(p3Out, p3Cap, p3SiteOff, p3SiteCnt, M, N, spw, firstSites, secondSites, colBase, res->yz, ctx->stream));
        ctx->launches[PBWTGPU_K_PACK3] += 1;
      } else {
        CUDA_TRY(ctx, p3_emit_columns(yOut, wpc, M, N, colBase, res->yz, ctx->stream));
        ctx->launches[PBWTGPU_K_PACK3] += (N > 0);
      }
    }
    CUDA_TRY(ctx, stream_sync(ctx));  // temporaries go back to the cache with this object
    return 0;
  }
};

// pass 2 of the window-parallel query sweep: every window from its true panel state (slots of the scan) and its true
// query state (the chain), records through the chunk pool, (site, CTA) segments into callback order, "no match" events
// sorted by (site, query position).  Outputs are device arrays owned by ws.tmp.
int finish_query_sweep(WindowedSweep &ws, pbwtgpu_match **recDev, unsigned long long *nrec, pbwtgpu_nomatch **nmDev,
                       unsigned long long *nnm) {
  pbwtgpu_ctx *ctx = ws.ctx;
  Temps &tmp = ws.tmp;
  const int M = ws.M, N = ws.N, Q = ws.queryQ, G = ws.G, W = ws.W;
  const long long nSeg = (long long)(N + 1) * G + 1;
  long long *segDst = nullptr;
  char *scanTemp = nullptr;
  if (int rc = tmp.get(&segDst, (size_t)nSeg)) return rc;
  const size_t scanTempBytes = segment_scan_temp_bytes(nSeg);
  if (int rc = tmp.get(&scanTemp, scanTempBytes)) return rc;
  unsigned long long *counts = nullptr;  // [0] unused, [1] "no match" events
  int *chunkCounter = nullptr, *chunkTab = nullptr;
  if (int rc = tmp.get(&counts, 2)) return rc;
  if (int rc = tmp.get(&chunkCounter, 1)) return rc;
  long long recCap = ctx->optMatchCap > 0 ? ctx->optMatchCap : std::max<long long>(1 << 20, 4096LL * Q);
  long long nmCap = 1 << 16;
  pbwtgpu_match *pool = nullptr;
  pbwtgpu_nomatch *nm = nullptr;
  unsigned long long got[2] = {0, 0};
  int poolChunks = 0;
  const int producers = W * G;
  // the chain's output is consumed in place by pass 2: keep a copy for a rerun after a pool overflow
  int32_t *qfKeep = nullptr, *qeKeep = nullptr;
  if (int rc = tmp.get(&qfKeep, (size_t)W * Q)) return rc;
  if (int rc = tmp.get(&qeKeep, (size_t)W * Q)) return rc;
  CUDA_TRY(ctx, cudaMemcpyAsync(qfKeep, ws.qfT, sizeof(int32_t) * (size_t)W * Q, cudaMemcpyDeviceToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(qeKeep, ws.qeT, sizeof(int32_t) * (size_t)W * Q, cudaMemcpyDeviceToDevice, ctx->stream));
  for (int attempt = 0; attempt < 2; ++attempt) {
    poolChunks = (int)((recCap + kRecChunk - 1) >> kRecChunkLog) + producers;
    if (int rc = tmp.get(&pool, (size_t)poolChunks << kRecChunkLog)) return rc;
    if (int rc = tmp.get(&chunkTab, (size_t)producers * poolChunks)) return rc;
    if (int rc = tmp.get(&nm, (size_t)nmCap)) return rc;
    if (int rc = tmp.get(&ws.gb.nmKeys, (size_t)nmCap)) return rc;
    CUDA_TRY(ctx, cudaMemsetAsync(chunkCounter, 0, sizeof(int), ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(counts, 0, 2 * sizeof(unsigned long long), ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(ws.gb.segCnt, 0, sizeof(int32_t) * (size_t)nSeg, ctx->stream));
    if (attempt) {
      CUDA_TRY(ctx, cudaMemcpyAsync(ws.qfT, qfKeep, sizeof(int32_t) * (size_t)W * Q, cudaMemcpyDeviceToDevice, ctx->stream));
      CUDA_TRY(ctx, cudaMemcpyAsync(ws.qeT, qeKeep, sizeof(int32_t) * (size_t)W * Q, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    SweepArgs a{};
    a.src = ws.src; a.wpc = ws.wpc; a.M = M; a.N = N;
    a.sitesPerWindow = ws.spw; a.nWindows = W;
    a.siteOffset = ws.siteOffset; a.nTotal = ws.nTotal;
    a.aStart = ws.ord[ws.cur];
    a.dStart = ws.haveDStart ? ws.dv[ws.cur] : nullptr;
    a.recPool = pool; a.poolChunks = poolChunks; a.chunkCounter = chunkCounter; a.chunkTab = chunkTab;
    a.nomatch = nm; a.nomatchCap = nmCap; a.qCounts = counts;
    ws.query_args(a, 2, ws.qfT, ws.qeT);
    {
      ScopedTimer t(ctx, PBWTGPU_K_QUERY);
      CUDA_TRY(ctx, ws.launch_query(a));
      ctx->launches[PBWTGPU_K_QUERY] += 1;
    }
    int used = 0;
    long long segTotal = 0;
    {
      ScopedTimer t(ctx, PBWTGPU_K_AUX);
      CUDA_TRY(ctx, launch_segment_scan(ws.gb.segCnt, nSeg, segDst, scanTemp, scanTempBytes, ctx->stream));
      ctx->launches[PBWTGPU_K_AUX] += 1;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(&segTotal, segDst + nSeg - 1, sizeof segTotal, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(got, counts, sizeof got, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(&used, chunkCounter, sizeof used, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, stream_sync(ctx));
    got[0] = (unsigned long long)segTotal;
    if (used <= poolChunks && (long long)got[1] <= nmCap) break;
    if (attempt == 1) return fail(ctx, PBWTGPU_ERR_CUDA, "query sweep buffer overflow persisted");
    recCap = std::max<long long>(recCap, (long long)got[0]);
    nmCap = std::max<long long>(nmCap, (long long)got[1]);
  }
  pbwtgpu_match *rec = nullptr;
  pbwtgpu_nomatch *nmSorted = nm;
  if (int rc = tmp.get(&rec, (size_t)std::max<unsigned long long>(got[0], 1))) return rc;
  {
    ScopedTimer t(ctx, PBWTGPU_K_AUX);
    CUDA_TRY(ctx, launch_gather_segments(pool, chunkTab, poolChunks, ws.gb.segSrc, segDst, nSeg - 1, (long long)got[0], G, ws.spw, N, W,
                                         rec, ctx->stream));
    ctx->launches[PBWTGPU_K_AUX] += 1;
  }
  if (got[1]) {  // events were appended in arrival order: put them in (site, query-PBWT position) order
    if (int rc = tmp.get(&nmSorted, (size_t)got[1])) return rc;
    CUDA_TRY(ctx, sort_nomatch(ws.gb.nmKeys, nm, (long long)got[1], nmSorted, ctx->stream));
  }
  *recDev = rec; *nrec = got[0]; *nmDev = nmSorted; *nnm = got[1];
  return 0;
}

int run_sweep(pbwtgpu_ctx *ctx, const uint32_t *src, int wpc, int M, int N, SweepMode mode, const int32_t *aStartHost,
              unsigned flags, pbwtgpu_result *res) {
  WindowedSweep ws(ctx);
  if (int rc = ws.init(src, wpc, M, N, mode, aStartHost, (flags & (PBWTGPU_WANT_MAXWITHIN | PBWTGPU_WANT_LONGWITHIN)) != 0, false, 0, 0)) return rc;
  if (int rc = ws.prepare()) return rc;
  return ws.finish(flags, res);
}

int decode_stream(pbwtgpu_ctx *ctx, Temps &tmp, const uint8_t *yzHost, int64_t nyz, int M, int N, uint32_t **ycolsOut,
                  int32_t *zeroCountsDev) {
  const int wpc = pbwtgpu_words_per_col(M);
  uint8_t *yzDev = nullptr;
  uint32_t *ycols = nullptr;
  int64_t *tiles = nullptr;
  if (int rc = tmp.get(&yzDev, (size_t)nyz + 16)) return rc;
  if (int rc = tmp.get(&ycols, (size_t)std::max(N, 1) * wpc)) return rc;
  const int64_t nTiles = p3_decode_num_tiles(nyz);
  if (int rc = tmp.get(&tiles, (size_t)nTiles + 1)) return rc;
  if (nyz) CUDA_TRY(ctx, cudaMemcpyAsync(yzDev, yzHost, (size_t)nyz, cudaMemcpyHostToDevice, ctx->stream));
  {
    ScopedTimer t(ctx, PBWTGPU_K_UNPACK3);
    CUDA_TRY(ctx, p3_decode_columns(yzDev, nyz, M, N, wpc, tiles, ycols, zeroCountsDev, ctx->stream));
    ctx->launches[PBWTGPU_K_UNPACK3] += nTiles ? 3 : 0;
  }
  int64_t symbols = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(&symbols, tiles + nTiles, sizeof symbols, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, stream_sync(ctx));
  if (symbols != (int64_t)M * N)  // the reference's isCheck assertion, pbwtCore.c:301-302
    return fail(ctx, PBWTGPU_ERR_ARG, "pack3 stream decodes to %lld symbols, expected M*N = %lld", (long long)symbols,
                (long long)M * N);
  *ycolsOut = ycols;
  return 0;
}

}  // namespace

// ================================================================================================
extern "C" {

int pbwtgpu_words_per_col(int M) { return 4 * ((M + 127) / 128); }

const char *pbwtgpu_strerror(int code) {
  switch (code) {
    case PBWTGPU_OK: return "ok";
    case PBWTGPU_ERR_CUDA: return "CUDA error";
    case PBWTGPU_ERR_ARG: return "invalid argument";
    case PBWTGPU_ERR_NOMEM: return "out of memory";
    case PBWTGPU_ERR_UNSUPPORTED: return "unsupported shape";
    case PBWTGPU_ERR_NODEVICE: return "no sm_100 CUDA device (there is no CPU fallback)";
    default: return "unknown error";
  }
}

int pbwtgpu_init(int device, pbwtgpu_ctx **out) {
  if (!out) return PBWTGPU_ERR_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) return PBWTGPU_ERR_NODEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PBWTGPU_ERR_NODEVICE;
  if (prop.major != 10) return PBWTGPU_ERR_NODEVICE;  // the kernels are built for sm_100a only
  if (cudaSetDevice(device) != cudaSuccess) return PBWTGPU_ERR_CUDA;
  pbwtgpu_ctx *ctx = new pbwtgpu_ctx;
  ctx->device = device;
  ctx->smCount = prop.multiProcessorCount;
  ctx->cacheCapBytes = (size_t)(0.6 * (double)prop.totalGlobalMem);  // (config 4's temporaries alone are ~50 GB)
  if (cudaStreamCreateWithFlags(&ctx->ownStream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return PBWTGPU_ERR_CUDA;
  }
  ctx->stream = ctx->ownStream;
  if (cudaMallocHost((void **)&ctx->pinned, sizeof(unsigned long long) * pbwtgpu_ctx::kPinnedWords) != cudaSuccess) {
    cudaStreamDestroy(ctx->ownStream);
    delete ctx;
    return PBWTGPU_ERR_NOMEM;
  }
  *out = ctx;
  return 0;
}

void pbwtgpu_destroy(pbwtgpu_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  stream_sync(ctx);
  reset_timings(ctx);
  for (auto e : ctx->eventPool) cudaEventDestroy(e);
  if (ctx->syncEvent) cudaEventDestroy(ctx->syncEvent);
  for (auto &b : ctx->freeBlocks) cudaFree(b.p);
  cudaFreeHost(ctx->pinned);
  cudaStreamDestroy(ctx->ownStream);
  delete ctx;
}

const char *pbwtgpu_last_error(pbwtgpu_ctx *ctx) { return ctx ? ctx->lastError.c_str() : "no context"; }

int pbwtgpu_set_stream(pbwtgpu_ctx *ctx, void *s) {
  if (!ctx) return PBWTGPU_ERR_ARG;
  // cached device blocks are reused without regard to the stream they were last used on: drain the old one first
  CUDA_TRY(ctx, stream_sync(ctx));
  ctx->stream = s ? (cudaStream_t)s : ctx->ownStream;
  return 0;
}
void *pbwtgpu_get_stream(pbwtgpu_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int pbwtgpu_synchronize(pbwtgpu_ctx *ctx) {
  if (!ctx) return PBWTGPU_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CUDA_TRY(ctx, stream_sync(ctx));
  return 0;
}

int pbwtgpu_set_option(pbwtgpu_ctx *ctx, const char *name, long long value) {
  if (!ctx || !name) return PBWTGPU_ERR_ARG;
  if (!strcmp(name, "windows")) ctx->optWindows = value;
  else if (!strcmp(name, "match_capacity")) ctx->optMatchCap = value;
  else if (!strcmp(name, "blocking_sync")) ctx->optBlockingSync = value;
  else if (!strcmp(name, "window_times")) ctx->optWindowTimes = value;
  else if (!strcmp(name, "force_global")) ctx->optForceGlobal = value;
  else if (!strcmp(name, "no_cluster")) ctx->optNoCluster = value;
  else if (!strcmp(name, "no_sort_combine")) ctx->optNoSortCombine = value;
  else if (!strcmp(name, "no_fused_pack3")) ctx->optNoFusedPack3 = value;
  else if (!strcmp(name, "pack3_window_capacity")) ctx->optPack3Cap = value;
  else if (!strcmp(name, "long_within")) ctx->optLongT = value;
  else if (!strcmp(name, "no_sort_pass1")) ctx->optNoSortPass1 = value;
  else if (!strcmp(name, "equal_windows")) ctx->optEqualWindows = value;
  else if (!strcmp(name, "debug_no_candidates")) ctx->optNoCandidates = value;
  else if (!strcmp(name, "edge_first_pct")) ctx->optEdgePct[0] = value;
  else if (!strcmp(name, "edge_second_pct")) ctx->optEdgePct[1] = value;
  else if (!strcmp(name, "edge_last_pct")) ctx->optEdgePct[2] = value;
  else return fail(ctx, PBWTGPU_ERR_ARG, "unknown option %s", name);
  return 0;
}

int pbwtgpu_profile(pbwtgpu_ctx *ctx, int enable) {
  if (!ctx) return PBWTGPU_ERR_ARG;
  ctx->profile = enable != 0;
  reset_timings(ctx);
  return 0;
}

int pbwtgpu_last_timings(pbwtgpu_ctx *ctx, float *ms, int *launches) {
  if (!ctx) return PBWTGPU_ERR_ARG;
  CUDA_TRY(ctx, stream_sync(ctx));
  for (auto &p : ctx->pending) {
    float t = 0;
    if (cudaEventElapsedTime(&t, p.a, p.b) == cudaSuccess) ctx->ms[p.cls] += t;
    ctx->eventPool.push_back(p.a);
    ctx->eventPool.push_back(p.b);
  }
  ctx->pending.clear();
  for (int i = 0; i < PBWTGPU_NUM_KERNEL_CLASSES; ++i) {
    if (ms) ms[i] = ctx->ms[i];
    if (launches) launches[i] = ctx->launches[i];
  }
  memset(ctx->ms, 0, sizeof ctx->ms);
  memset(ctx->launches, 0, sizeof ctx->launches);
  return 0;
}

// ---- panels ------------------------------------------------------------------------------------
int pbwtgpu_panel_from_host(pbwtgpu_ctx *ctx, const uint32_t *xbits, int M, int N, pbwtgpu_panel **out) {
  if (!ctx || !out || M < 1 || N < 0 || (N > 0 && !xbits)) return fail(ctx, PBWTGPU_ERR_ARG, "panel_from_host: bad args");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  pbwtgpu_panel *p = new pbwtgpu_panel;
  p->M = M;
  p->N = N;
  p->wpc = pbwtgpu_words_per_col(M);
  void *d = nullptr;
  if (int rc = dev_alloc(ctx, &d, (size_t)std::max(N, 1) * p->wpc * 4)) { delete p; return rc; }
  p->bits = (uint32_t *)d;
  if (N) CUDA_TRY(ctx, cudaMemcpyAsync(p->bits, xbits, (size_t)N * p->wpc * 4, cudaMemcpyHostToDevice, ctx->stream));
  *out = p;
  return 0;
}

int pbwtgpu_panel_synthetic(pbwtgpu_ctx *ctx, int M, int N, uint64_t seed, uint32_t L, uint32_t flip_thresh,
                            pbwtgpu_panel **out) {
  return pbwtgpu_panel_synthetic_range(ctx, M, N, 0, seed, L, flip_thresh, out);
}

int pbwtgpu_panel_synthetic_range(pbwtgpu_ctx *ctx, int M, int N, int first_site, uint64_t seed, uint32_t L,
                                  uint32_t flip_thresh, pbwtgpu_panel **out) {
  if (!ctx || !out || M < 1 || N < 0 || L < 1 || first_site < 0) return fail(ctx, PBWTGPU_ERR_ARG, "panel_synthetic: bad args");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  pbwtgpu_panel *p = new pbwtgpu_panel;
  p->M = M;
  p->N = N;
  p->wpc = pbwtgpu_words_per_col(M);
  void *d = nullptr;
  if (int rc = dev_alloc(ctx, &d, (size_t)std::max(N, 1) * p->wpc * 4)) { delete p; return rc; }
  p->bits = (uint32_t *)d;
  {
    ScopedTimer t(ctx, PBWTGPU_K_SYNTH);
    CUDA_TRY(ctx, launch_synth(p->bits, p->wpc, M, N, first_site, seed, L, flip_thresh, ctx->stream));
    ctx->launches[PBWTGPU_K_SYNTH] += N > 0;
  }
  *out = p;
  return 0;
}

int pbwtgpu_panel_to_host(pbwtgpu_ctx *ctx, const pbwtgpu_panel *p, int first, int n, uint32_t *xbits) {
  if (!ctx || !p || first < 0 || n < 0 || first + n > p->N) return fail(ctx, PBWTGPU_ERR_ARG, "panel_to_host: bad range");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (n)
    CUDA_TRY(ctx, cudaMemcpyAsync(xbits, p->bits + (size_t)first * p->wpc, (size_t)n * p->wpc * 4, cudaMemcpyDeviceToHost,
                                  ctx->stream));
  CUDA_TRY(ctx, stream_sync(ctx));
  return 0;
}

void pbwtgpu_panel_free(pbwtgpu_ctx *ctx, pbwtgpu_panel *p) {
  if (!p) return;
  if (ctx) {
    cudaSetDevice(ctx->device);
    stream_sync(ctx);
    dev_release(ctx, p->bits, (size_t)std::max(p->N, 1) * p->wpc * 4);
  }
  delete p;
}

// ---- sweeps ------------------------------------------------------------------------------------
int pbwtgpu_build_device(pbwtgpu_ctx *ctx, const pbwtgpu_panel *panel, const int32_t *a_start, unsigned flags,
                         pbwtgpu_result **out) {
  if (!ctx || !panel || !out) return fail(ctx, PBWTGPU_ERR_ARG, "build_device: bad args");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  pbwtgpu_result *res = new pbwtgpu_result;
  const int rc = run_sweep(ctx, panel->bits, panel->wpc, panel->M, panel->N, kGather, a_start, flags, res);
  if (rc) { pbwtgpu_result_free(ctx, res); return rc; }
  *out = res;
  return 0;
}

int pbwtgpu_sweep_yz_device(pbwtgpu_ctx *ctx, const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *a_start,
                            unsigned flags, pbwtgpu_result **out) {
  if (!ctx || !out || nyz < 0 || (nyz > 0 && !yz) || M < 1 || N < 0) return fail(ctx, PBWTGPU_ERR_ARG, "sweep_yz: bad args");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  Temps tmp(ctx);
  uint32_t *ycols = nullptr;
  if (int rc = decode_stream(ctx, tmp, yz, nyz, M, N, &ycols, nullptr)) return rc;
  pbwtgpu_result *res = new pbwtgpu_result;
  const int rc = run_sweep(ctx, ycols, pbwtgpu_words_per_col(M), M, N, kSorted, a_start, flags | PBWTGPU_WANT_D, res);
  if (rc) { pbwtgpu_result_free(ctx, res); return rc; }
  if (cudaError_t e = stream_sync(ctx)) {  // temporaries are recycled on return
    pbwtgpu_result_free(ctx, res);
    return fail(ctx, PBWTGPU_ERR_CUDA, "sweep_yz: %s", cudaGetErrorString(e));
  }
  *out = res;
  return 0;
}

int pbwtgpu_result_sizes(const pbwtgpu_result *r, int64_t *nyz, int64_t *nmatch) {
  if (!r) return PBWTGPU_ERR_ARG;
  if (nyz) *nyz = r->nyz;
  if (nmatch) *nmatch = r->nrec;
  return 0;
}

int pbwtgpu_result_fetch(pbwtgpu_ctx *ctx, const pbwtgpu_result *r, uint8_t *yz, int32_t *a_end, int32_t *d_end,
                         pbwtgpu_match *matches) {
  if (!ctx || !r) return fail(ctx, PBWTGPU_ERR_ARG, "result_fetch: bad args");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));  // (callers may come from any host thread)
  if (yz && r->yz && r->nyz)
    CUDA_TRY(ctx, cudaMemcpyAsync(yz, r->yz, (size_t)r->nyz, cudaMemcpyDeviceToHost, ctx->stream));
  if (a_end && r->aEnd)
    CUDA_TRY(ctx, cudaMemcpyAsync(a_end, r->aEnd, sizeof(int32_t) * r->M, cudaMemcpyDeviceToHost, ctx->stream));
  if (d_end && r->dEnd)
    CUDA_TRY(ctx, cudaMemcpyAsync(d_end, r->dEnd, sizeof(int32_t) * (r->M + 1), cudaMemcpyDeviceToHost, ctx->stream));
  if (matches && r->rec && r->nrec)
    CUDA_TRY(ctx, cudaMemcpyAsync(matches, r->rec, sizeof(pbwtgpu_match) * (size_t)r->nrec, cudaMemcpyDeviceToHost,
                                  ctx->stream));
  CUDA_TRY(ctx, stream_sync(ctx));
  return 0;
}

int pbwtgpu_result_group(pbwtgpu_ctx *ctx, pbwtgpu_result *r, int64_t *nwords, int64_t *ngroups) {
  if (!ctx || !r) return fail(ctx, PBWTGPU_ERR_ARG, "result_group: bad args");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (r->nGroups < 0) {
    long long g = 0;
    if (r->nrec > 0) {
      if (r->nrec > 0x7fffffffLL) return fail(ctx, PBWTGPU_ERR_UNSUPPORTED, "result_group: %lld records", (long long)r->nrec);
      Temps tmp(ctx);
      char *state = nullptr, *temp = nullptr;
      const size_t tempBytes = group_records_temp_bytes(r->nrec);
      if (int rc = tmp.get(&state, group_records_state_bytes(r->nrec))) return rc;
      if (int rc = tmp.get(&temp, tempBytes)) return rc;
      ScopedTimer t(ctx, PBWTGPU_K_AUX);
      CUDA_TRY(ctx, launch_group_scan(r->rec, r->nrec, state, temp, tempBytes, ctx->stream));
      CUDA_TRY(ctx, cudaMemcpyAsync(&ctx->pinned[0], group_scan_total(state, r->nrec), sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(ctx, stream_sync(ctx));
      g = (long long)ctx->pinned[0];
      if (int rc = result_alloc(ctx, r, &r->groups, (size_t)(4 * g + r->nrec))) return rc;
      CUDA_TRY(ctx, launch_write_groups(r->rec, state, r->nrec, r->groups, ctx->stream));
      ctx->launches[PBWTGPU_K_AUX] += 2;
      CUDA_TRY(ctx, stream_sync(ctx));  // the temporaries go back to the cache
    }
    r->nGroups = g;
  }
  if (nwords) *nwords = 4 * r->nGroups + r->nrec;
  if (ngroups) *ngroups = r->nGroups;
  return 0;
}

int pbwtgpu_result_fetch_groups(pbwtgpu_ctx *ctx, const pbwtgpu_result *r, int32_t *dst) {
  if (!ctx || !r || !dst || r->nGroups < 0) return fail(ctx, PBWTGPU_ERR_ARG, "result_fetch_groups: call pbwtgpu_result_group first");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t n = (size_t)(4 * r->nGroups + r->nrec);
  if (n) CUDA_TRY(ctx, cudaMemcpyAsync(dst, r->groups, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, stream_sync(ctx));
  return 0;
}

int pbwtgpu_result_fetch_hist(pbwtgpu_ctx *ctx, const pbwtgpu_result *r, int64_t *hist) {
  if (!ctx || !r || !hist || !r->hist) return fail(ctx, PBWTGPU_ERR_ARG, "result_fetch_hist: the sweep was not run with WANT_STATS");
  std::vector<unsigned int> h((size_t)r->N + 1);
  CUDA_TRY(ctx, cudaMemcpyAsync(h.data(), r->hist, sizeof(unsigned int) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, stream_sync(ctx));
  for (size_t i = 0; i < h.size(); ++i) hist[i] = (int64_t)h[i];
  return 0;
}

void pbwtgpu_result_free(pbwtgpu_ctx *ctx, pbwtgpu_result *r) {
  if (!r) return;
  if (ctx) {
    cudaSetDevice(ctx->device);
    stream_sync(ctx);
    for (size_t i = 0; i < r->owned.size(); ++i) dev_release(ctx, r->owned[i], r->ownedBytes[i]);
  }
  delete r;
}

// ---- multi-GPU site shards ---------------------------------------------------------------------------
struct pbwtgpu_shard {
  WindowedSweep ws;
  explicit pbwtgpu_shard(pbwtgpu_ctx *c) : ws(c) {}
};

size_t pbwtgpu_shard_aggregate_ints(int M) { return WindowedSweep::agg_ints(M); }

int pbwtgpu_shard_prepare(pbwtgpu_ctx *ctx, const pbwtgpu_panel *panel, const int32_t *a_start, int site_offset,
                          int total_sites, unsigned flags, pbwtgpu_shard **out) {
  if (!ctx || !panel || !out || site_offset < 0 || total_sites < site_offset + panel->N || panel->N < 1)
    return fail(ctx, PBWTGPU_ERR_ARG, "shard_prepare: bad args");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  pbwtgpu_shard *sh = new pbwtgpu_shard(ctx);
  int rc = sh->ws.init(panel->bits, panel->wpc, panel->M, panel->N, kGather, a_start, (flags & (PBWTGPU_WANT_MAXWITHIN | PBWTGPU_WANT_LONGWITHIN)) != 0, true,
                       site_offset, total_sites);
  if (!rc) rc = sh->ws.prepare();
  if (rc) { delete sh; return rc; }
  *out = sh;
  return 0;
}

int pbwtgpu_shard_export(pbwtgpu_ctx *ctx, pbwtgpu_shard *sh, int32_t *dev_dst) {
  if (!ctx || !sh || !dev_dst) return fail(ctx, PBWTGPU_ERR_ARG, "shard_export: bad args");
  if (int rc = sh->ws.export_aggregate(dev_dst)) return rc;
  CUDA_TRY(ctx, stream_sync(ctx));
  return 0;
}

int pbwtgpu_shard_finish(pbwtgpu_ctx *ctx, pbwtgpu_shard *sh, const int32_t *dev_aggs, int rank, unsigned flags,
                         pbwtgpu_result **out) {
  if (!ctx || !sh || !out || rank < 0 || (rank > 0 && !dev_aggs)) return fail(ctx, PBWTGPU_ERR_ARG, "shard_finish: bad args");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (int rc = sh->ws.apply_prefix(dev_aggs, rank)) return rc;
  pbwtgpu_result *res = new pbwtgpu_result;
  const int rc = sh->ws.finish(flags, res);
  if (rc) { pbwtgpu_result_free(ctx, res); return rc; }
  *out = res;
  return 0;
}

void pbwtgpu_shard_free(pbwtgpu_ctx *ctx, pbwtgpu_shard *sh) {
  if (!sh) return;
  if (ctx) stream_sync(ctx);
  delete sh;
}

// ---- one-shot host-buffer entry points -----------------------------------------------------------
// Outputs handed to the caller come from the pinned registry (full PCIe rate, reused across calls); the caller gives
// them back with pbwtgpu_free.
static int fetch_to_malloc(pbwtgpu_ctx *ctx, pbwtgpu_result *res, uint8_t **yz, int64_t *nyz, int32_t *a_end,
                           int32_t *d_end, pbwtgpu_match **matches, int64_t *nmatch) {
  HostGuard hy(yz ? host_alloc_pinned((size_t)std::max<int64_t>(res->nyz, 1)) : nullptr);
  HostGuard hm(matches ? host_alloc_pinned(sizeof(pbwtgpu_match) * (size_t)std::max<int64_t>(res->nrec, 1)) : nullptr);
  if ((yz && !hy.p) || (matches && !hm.p)) return fail(ctx, PBWTGPU_ERR_NOMEM, "pinned host allocation failed");
  if (int rc = pbwtgpu_result_fetch(ctx, res, (uint8_t *)hy.p, a_end, d_end, (pbwtgpu_match *)hm.p)) return rc;
  if (yz) { *yz = (uint8_t *)hy.release(); if (nyz) *nyz = res->nyz; }
  if (matches) { *matches = (pbwtgpu_match *)hm.release(); if (nmatch) *nmatch = res->nrec; }
  return 0;
}

int pbwtgpu_build(pbwtgpu_ctx *ctx, const uint32_t *xbits, int M, int N, const int32_t *a_start, unsigned flags,
                  uint8_t **yz, int64_t *nyz, int32_t *a_end, int32_t *d_end, pbwtgpu_match **matches,
                  int64_t *nmatch) {
  pbwtgpu_panel *panel = nullptr;
  if (int rc = pbwtgpu_panel_from_host(ctx, xbits, M, N, &panel)) return rc;
  pbwtgpu_result *res = nullptr;
  int rc = pbwtgpu_build_device(ctx, panel, a_start, flags, &res);
  if (!rc) rc = fetch_to_malloc(ctx, res, (flags & PBWTGPU_WANT_YZ) ? yz : nullptr, nyz, a_end, d_end,
                                (flags & (PBWTGPU_WANT_MAXWITHIN | PBWTGPU_WANT_LONGWITHIN)) ? matches : nullptr, nmatch);
  pbwtgpu_result_free(ctx, res);
  pbwtgpu_panel_free(ctx, panel);
  return rc;
}

int pbwtgpu_max_within(pbwtgpu_ctx *ctx, const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *a_start,
                       int32_t *a_end, int32_t *d_end, pbwtgpu_match **matches, int64_t *nmatch) {
  pbwtgpu_result *res = nullptr;
  int rc = pbwtgpu_sweep_yz_device(ctx, yz, nyz, M, N, a_start, matches ? PBWTGPU_WANT_MAXWITHIN : PBWTGPU_WANT_D, &res);
  if (!rc) rc = fetch_to_malloc(ctx, res, nullptr, nullptr, a_end, d_end, matches, nmatch);
  pbwtgpu_result_free(ctx, res);
  return rc;
}

// yz / a_end / d_end as fetch_to_malloc, the matches as a grouped stream
static int fetch_grouped(pbwtgpu_ctx *ctx, pbwtgpu_result *res, uint8_t **yz, int64_t *nyz, int32_t *a_end, int32_t *d_end, int32_t **groups,
                         int64_t *nwords) {
  int64_t nw = 0;
  if (int rc = pbwtgpu_result_group(ctx, res, &nw, nullptr)) return rc;
  HostGuard hg(host_alloc_pinned(sizeof(int32_t) * (size_t)std::max<int64_t>(nw, 1)));
  if (!hg.p) return fail(ctx, PBWTGPU_ERR_NOMEM, "pinned host allocation failed (grouped records)");
  if (int rc = pbwtgpu_result_fetch_groups(ctx, res, (int32_t *)hg.p)) return rc;
  if (int rc = fetch_to_malloc(ctx, res, yz, nyz, a_end, d_end, nullptr, nullptr)) return rc;
  *groups = (int32_t *)hg.release();
  *nwords = nw;
  return 0;
}

int pbwtgpu_build_grouped(pbwtgpu_ctx *ctx, const uint32_t *xbits, int M, int N, const int32_t *a_start, unsigned flags, uint8_t **yz,
                          int64_t *nyz, int32_t *a_end, int32_t *d_end, int32_t **groups, int64_t *nwords) {
  if (!groups || !nwords || !(flags & (PBWTGPU_WANT_MAXWITHIN | PBWTGPU_WANT_LONGWITHIN)) || (flags & PBWTGPU_WANT_STATS))
    return fail(ctx, PBWTGPU_ERR_ARG, "build_grouped: needs a match flag and somewhere to put the groups");
  pbwtgpu_panel *panel = nullptr;
  if (int rc = pbwtgpu_panel_from_host(ctx, xbits, M, N, &panel)) return rc;
  pbwtgpu_result *res = nullptr;
  int rc = pbwtgpu_build_device(ctx, panel, a_start, flags, &res);
  if (!rc) rc = fetch_grouped(ctx, res, (flags & PBWTGPU_WANT_YZ) ? yz : nullptr, nyz, a_end, d_end, groups, nwords);
  pbwtgpu_result_free(ctx, res);
  pbwtgpu_panel_free(ctx, panel);
  return rc;
}

int pbwtgpu_max_within_grouped(pbwtgpu_ctx *ctx, const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *a_start, int32_t *a_end,
                               int32_t *d_end, int32_t **groups, int64_t *nwords) {
  if (!groups || !nwords) return fail(ctx, PBWTGPU_ERR_ARG, "max_within_grouped: bad args");
  pbwtgpu_result *res = nullptr;
  int rc = pbwtgpu_sweep_yz_device(ctx, yz, nyz, M, N, a_start, PBWTGPU_WANT_MAXWITHIN, &res);
  if (!rc) rc = fetch_grouped(ctx, res, nullptr, nullptr, a_end, d_end, groups, nwords);
  pbwtgpu_result_free(ctx, res);
  return rc;
}

int pbwtgpu_max_within_stats(pbwtgpu_ctx *ctx, const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *a_start, int64_t *hist) {
  if (!hist) return fail(ctx, PBWTGPU_ERR_ARG, "max_within_stats: bad args");
  pbwtgpu_result *res = nullptr;
  int rc = pbwtgpu_sweep_yz_device(ctx, yz, nyz, M, N, a_start, PBWTGPU_WANT_MAXWITHIN | PBWTGPU_WANT_STATS, &res);
  if (!rc) rc = pbwtgpu_result_fetch_hist(ctx, res, hist);
  pbwtgpu_result_free(ctx, res);
  return rc;
}

int pbwtgpu_cursor_at(pbwtgpu_ctx *ctx, const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *a_start, int k,
                      uint8_t *y, int32_t *a, int32_t *d, int32_t *c) {
  if (!ctx || nyz < 0 || (nyz > 0 && !yz) || M < 1 || N < 0 || k < 0 || k > N) return fail(ctx, PBWTGPU_ERR_ARG, "cursor_at: bad args");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  Temps tmp(ctx);
  const int wpc = pbwtgpu_words_per_col(M);
  uint32_t *ycols = nullptr;
  int32_t *zc = nullptr;
  if (int rc = tmp.get(&zc, (size_t)N + 1)) return rc;
  if (int rc = decode_stream(ctx, tmp, yz, nyz, M, N, &ycols, zc)) return rc;
  // state entering site k = the result of sweeping the first k sorted columns (pbwtCursorForwardsReadAD x k)
  pbwtgpu_result *res = new pbwtgpu_result;
  int rc = run_sweep(ctx, ycols, wpc, M, k, kSorted, a_start, PBWTGPU_WANT_D, res);
  if (!rc) rc = pbwtgpu_result_fetch(ctx, res, nullptr, a, d, nullptr);
  pbwtgpu_result_free(ctx, res);
  if (rc) return rc;
  if (k < N) {  // the column the cursor holds there, and its zero count (u->y, u->c)
    if (y) {
      uint8_t *yDev = nullptr;
      if (int rc2 = tmp.get(&yDev, (size_t)M)) return rc2;
      CUDA_TRY(ctx, bits_to_bytes(ycols + (size_t)k * wpc, M, 1, wpc, yDev, ctx->stream));
      CUDA_TRY(ctx, cudaMemcpyAsync(y, yDev, (size_t)M, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (c) CUDA_TRY(ctx, cudaMemcpyAsync(c, zc + k, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, stream_sync(ctx));
  }
  return 0;
}

int pbwtgpu_build_reverse(pbwtgpu_ctx *ctx, const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *a_start,
                          int32_t *a_fend, uint8_t **zz, int64_t *nzz, int32_t *a_rend) {
  if (!ctx || nyz < 0 || (nyz > 0 && !yz) || M < 1 || N < 0 || !zz || !nzz || !a_rend)
    return fail(ctx, PBWTGPU_ERR_ARG, "build_reverse: bad args");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  Temps tmp(ctx);
  const int wpc = pbwtgpu_words_per_col(M);
  uint32_t *ycols = nullptr, *xRev = nullptr;
  if (int rc = decode_stream(ctx, tmp, yz, nyz, M, N, &ycols, nullptr)) return rc;
  if (int rc = tmp.get(&xRev, (size_t)std::max(N, 1) * wpc)) return rc;
  CUDA_TRY(ctx, cudaMemsetAsync(xRev, 0, sizeof(uint32_t) * (size_t)std::max(N, 1) * wpc, ctx->stream));
  std::vector<int32_t> aFend((size_t)M);
  {  // forwards to the end (pbwtCore.c:159-164), un-sorting every column on the way (:172)
    pbwtgpu_result *fwd = new pbwtgpu_result;
    WindowedSweep ws(ctx);
    int rc = ws.init(ycols, wpc, M, N, kSorted, a_start, false, false, 0, 0);
    ws.xOutRev = xRev;
    if (!rc) rc = ws.prepare();
    if (!rc) rc = ws.finish(0, fwd);
    if (!rc) rc = pbwtgpu_result_fetch(ctx, fwd, nullptr, aFend.data(), nullptr, nullptr);
    pbwtgpu_result_free(ctx, fwd);
    if (rc) return rc;
  }
  // the reverse cursor starts from aFend and sweeps the sites backwards (pbwtCore.c:166-176)
  pbwtgpu_result *rev = new pbwtgpu_result;
  int rc = run_sweep(ctx, xRev, wpc, M, N, kGather, aFend.data(), PBWTGPU_WANT_YZ, rev);
  if (!rc) rc = fetch_to_malloc(ctx, rev, zz, nzz, a_rend, nullptr, nullptr, nullptr);
  pbwtgpu_result_free(ctx, rev);
  if (rc) return rc;
  if (a_fend) memcpy(a_fend, aFend.data(), sizeof(int32_t) * (size_t)M);
  return 0;
}

int pbwtgpu_long_within(pbwtgpu_ctx *ctx, const uint8_t *yz, int64_t nyz, int M, int N, const int32_t *a_start, int L,
                        int32_t *a_end, int32_t *d_end, pbwtgpu_match **matches, int64_t *nmatch) {
  if (!ctx || L < 1 || !matches) return fail(ctx, PBWTGPU_ERR_ARG, "long_within: bad args (L must be >= 1)");
  const long long saved = ctx->optLongT;
  ctx->optLongT = L;
  pbwtgpu_result *res = nullptr;
  int rc = pbwtgpu_sweep_yz_device(ctx, yz, nyz, M, N, a_start, PBWTGPU_WANT_LONGWITHIN, &res);
  ctx->optLongT = saved;
  if (!rc) rc = fetch_to_malloc(ctx, res, nullptr, nullptr, a_end, d_end, matches, nmatch);
  pbwtgpu_result_free(ctx, res);
  return rc;
}

// records / events of a query sweep: device -> library-owned host buffers
static int query_results_to_host(pbwtgpu_ctx *ctx, const pbwtgpu_match *recDev, unsigned long long nrec, const pbwtgpu_nomatch *nmDev,
                                 unsigned long long nnm, pbwtgpu_match **matches, int64_t *nmatch, pbwtgpu_nomatch **nomatch,
                                 int64_t *nnomatch) {
  HostGuard gm(host_alloc_pinned(sizeof(pbwtgpu_match) * (size_t)std::max<unsigned long long>(nrec, 1)));
  HostGuard gn(host_alloc_pinned(sizeof(pbwtgpu_nomatch) * (size_t)std::max<unsigned long long>(nnm, 1)));
  if (!gm.p || !gn.p) return fail(ctx, PBWTGPU_ERR_NOMEM, "pinned host allocation failed (records)");
  if (nrec) CUDA_TRY(ctx, cudaMemcpyAsync(gm.p, recDev, sizeof(pbwtgpu_match) * nrec, cudaMemcpyDeviceToHost, ctx->stream));
  if (nnm) CUDA_TRY(ctx, cudaMemcpyAsync(gn.p, nmDev, sizeof(pbwtgpu_nomatch) * nnm, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, stream_sync(ctx));
  *matches = (pbwtgpu_match *)gm.release();
  *nmatch = (int64_t)nrec;
  if (nomatch) *nomatch = (pbwtgpu_nomatch *)gn.release();
  if (nnomatch) *nnomatch = (int64_t)nnm;
  return 0;
}

int pbwtgpu_match_sweep(pbwtgpu_ctx *ctx, const uint8_t *pyz, int64_t npyz, int M, int N, const int32_t *pa_start,
                        const uint8_t *qyz, int64_t nqyz, int Q, const int32_t *qa_start, pbwtgpu_match **matches,
                        int64_t *nmatch, pbwtgpu_nomatch **nomatch, int64_t *nnomatch) {
  if (!ctx || !matches || !nmatch || M < 1 || Q < 1 || N < 0) return fail(ctx, PBWTGPU_ERR_ARG, "match_sweep: bad args");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  // narrow panels: state in shared memory, one CTA, site after site.  Wide panels (and query sets beyond what that kernel
  // holds): state in global memory, groups of cooperating CTAs, and the SITES cut into windows that run concurrently
  // (query_chain.cu supplies the exact query state entering every window).
  const bool wide = M > query_sweep_max_m() || ctx->optForceGlobal || Q > query_sweep_max_q(M);
  const int maxQ = wide ? query_sweep_global_max_q() : query_sweep_max_q(M);
  if (Q > maxQ) return fail(ctx, PBWTGPU_ERR_UNSUPPORTED, "match_sweep: Q=%d exceeds %d queries per call", Q, maxQ);
  Temps tmp(ctx);
  uint32_t *pcols = nullptr, *qcols = nullptr;
  if (int rc = decode_stream(ctx, tmp, pyz, npyz, M, N, &pcols, nullptr)) return rc;
  if (int rc = decode_stream(ctx, tmp, qyz, nqyz, Q, N, &qcols, nullptr)) return rc;

  if (wide) {
    WindowedSweep wp(ctx), wq(ctx);
    wp.queryQ = Q; wp.qcols = qcols; wp.qwpc = pbwtgpu_words_per_col(Q);
    if (int rc = wp.init(pcols, pbwtgpu_words_per_col(M), M, N, kSorted, pa_start, false, false, 0, 0)) return rc;
    // the query PBWT's own order entering every window (pbwtMatch.c:377-378: queries are visited in that order): an
    // a-only window-parallel sweep of the Q-wide query stream over the same window boundaries
    int32_t *qaWin = nullptr;
    if (wp.W > 1 && Q >= 2) {
      wq.forceSpw = wp.spw;
      wq.allowForceGlobal = false;
      if (int rc = wq.init(qcols, pbwtgpu_words_per_col(Q), Q, N, kSorted, qa_start, false, false, 0, 0)) return rc;
      if (wq.W != wp.W) return fail(ctx, PBWTGPU_ERR_CUDA, "match_sweep: query windows %d != panel windows %d", wq.W, wp.W);
      if (int rc = wq.prepare()) return rc;
      qaWin = wq.ord[wq.cur];
    } else if (wp.W > 1) {  // a single query is always at position 0
      if (int rc = tmp.get(&qaWin, (size_t)wp.W)) return rc;
      CUDA_TRY(ctx, cudaMemsetAsync(qaWin, 0, sizeof(int32_t) * wp.W, ctx->stream));
    } else if (qa_start) {
      if (int rc = tmp.get(&qaWin, (size_t)Q)) return rc;
      CUDA_TRY(ctx, cudaMemcpyAsync(qaWin, qa_start, sizeof(int32_t) * Q, cudaMemcpyHostToDevice, ctx->stream));
    }
    wp.qaWin = qaWin;
    if (int rc = wp.prepare()) return rc;
    pbwtgpu_match *recDev = nullptr;
    pbwtgpu_nomatch *nmDev = nullptr;
    unsigned long long nrec = 0, nnm = 0;
    if (int rc = finish_query_sweep(wp, &recDev, &nrec, &nmDev, &nnm)) return rc;
    return query_results_to_host(ctx, recDev, nrec, nmDev, nnm, matches, nmatch, nomatch, nnomatch);
  }

  int32_t *pa = nullptr, *qa = nullptr;
  if (pa_start) {
    if (int rc = tmp.get(&pa, (size_t)M)) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(pa, pa_start, sizeof(int32_t) * M, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (qa_start) {
    if (int rc = tmp.get(&qa, (size_t)Q)) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(qa, qa_start, sizeof(int32_t) * Q, cudaMemcpyHostToDevice, ctx->stream));
  }
  unsigned long long *counts = nullptr;  // [0] records, [1] nomatch events
  int *chunkCounter = nullptr, *chunkTab = nullptr;
  if (int rc = tmp.get(&counts, 2)) return rc;
  if (int rc = tmp.get(&chunkCounter, 1)) return rc;
  long long recCap = ctx->optMatchCap > 0 ? ctx->optMatchCap : std::max<long long>(1 << 20, 4096LL * Q);
  long long nmCap = 1 << 16;
  pbwtgpu_match *pool = nullptr;
  pbwtgpu_nomatch *nm = nullptr;
  unsigned long long got[2] = {0, 0};
  int poolChunks = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    poolChunks = (int)((recCap + kRecChunk - 1) >> kRecChunkLog) + 1;
    if (int rc = tmp.get(&pool, (size_t)poolChunks << kRecChunkLog)) return rc;
    if (int rc = tmp.get(&chunkTab, (size_t)poolChunks)) return rc;
    if (int rc = tmp.get(&nm, (size_t)nmCap)) return rc;
    CUDA_TRY(ctx, cudaMemsetAsync(chunkCounter, 0, sizeof(int), ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(counts, 0, 2 * sizeof(unsigned long long), ctx->stream));
    SweepArgs a{};
    a.src = pcols; a.wpc = pbwtgpu_words_per_col(M); a.M = M; a.N = N;
    a.sitesPerWindow = N > 0 ? N : 1; a.nWindows = 1;
    a.aStart = pa;
    a.qcols = qcols; a.qwpc = pbwtgpu_words_per_col(Q); a.Q = Q; a.qaStart = qa;
    a.recPool = pool; a.poolChunks = poolChunks; a.chunkCounter = chunkCounter; a.chunkTab = chunkTab;
    a.nomatch = nm; a.nomatchCap = nmCap; a.qCounts = counts;
    {
      ScopedTimer t(ctx, PBWTGPU_K_QUERY);
      CUDA_TRY(ctx, launch_query_sweep_smem(a, ctx->stream));
      ctx->launches[PBWTGPU_K_QUERY] += 1;
    }
    int used = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(got, counts, sizeof got, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(&used, chunkCounter, sizeof used, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, stream_sync(ctx));
    if (used <= poolChunks && (long long)got[1] <= nmCap) break;
    if (attempt == 1) return fail(ctx, PBWTGPU_ERR_CUDA, "query sweep buffer overflow persisted");
    recCap = std::max<long long>(recCap, (long long)got[0]);
    nmCap = std::max<long long>(nmCap, (long long)got[1]);
  }
  pbwtgpu_match *recDev = nullptr;
  if (int rc = tmp.get(&recDev, (size_t)std::max<unsigned long long>(got[0], 1))) return rc;
  {
    long long *offsDev = nullptr;
    if (int rc = tmp.get(&offsDev, 1)) return rc;
    CUDA_TRY(ctx, cudaMemsetAsync(offsDev, 0, sizeof(long long), ctx->stream));
    ScopedTimer t(ctx, PBWTGPU_K_AUX);
    const int maxChunks = (int)((got[0] + kRecChunk - 1) >> kRecChunkLog);
    CUDA_TRY(ctx, launch_gather_records(pool, chunkTab, poolChunks, offsDev, counts, 1, maxChunks, recDev, ctx->stream));
    ctx->launches[PBWTGPU_K_AUX] += maxChunks > 0;
  }
  return query_results_to_host(ctx, recDev, got[0], nm, got[1], matches, nmatch, nomatch, nnomatch);
}

// ---- matchSequencesSweep with the SITES partitioned across GPUs (include/pbwtgpu.h) -------------------------------
struct pbwtgpu_qshard {
  Temps tmp;
  WindowedSweep wp, wq;  // panel (with the queries riding on its passes) and the Q-wide query stream's own order
  int M = 0, Q = 0;
  int32_t *qaZero = nullptr;
  explicit pbwtgpu_qshard(pbwtgpu_ctx *c) : tmp(c), wp(c), wq(c) {}
};

int pbwtgpu_qshard_open(pbwtgpu_ctx *ctx, const uint8_t *pyz, int64_t npyz, int M, int N, const int32_t *pa_start,
                        const uint8_t *qyz, int64_t nqyz, int Q, const int32_t *qa_start, int first_site, int nsites,
                        pbwtgpu_qshard **out) {
  if (!ctx || !out || M < 1 || Q < 1 || N < 1 || first_site < 0 || nsites < 1 || first_site + nsites > N)
    return fail(ctx, PBWTGPU_ERR_ARG, "qshard_open: bad args");
  if (Q > query_sweep_global_max_q()) return fail(ctx, PBWTGPU_ERR_UNSUPPORTED, "qshard_open: Q=%d exceeds %d queries per call", Q, query_sweep_global_max_q());
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  pbwtgpu_qshard *qs = new pbwtgpu_qshard(ctx);
  qs->M = M; qs->Q = Q;
  auto bail = [&](int rc) { stream_sync(ctx); delete qs; return rc; };
  // every GPU decodes the whole streams (they are small) and sweeps its own range of columns
  uint32_t *pcols = nullptr, *qcols = nullptr;
  if (int rc = decode_stream(ctx, qs->tmp, pyz, npyz, M, N, &pcols, nullptr)) return bail(rc);
  if (int rc = decode_stream(ctx, qs->tmp, qyz, nqyz, Q, N, &qcols, nullptr)) return bail(rc);
  const int wpc = pbwtgpu_words_per_col(M), qwpc = pbwtgpu_words_per_col(Q);
  qs->wp.queryQ = Q; qs->wp.qcols = qcols + (size_t)first_site * qwpc; qs->wp.qwpc = qwpc;
  if (int rc = qs->wp.init(pcols + (size_t)first_site * wpc, wpc, M, nsites, kSorted, pa_start, false, true, first_site, N)) return bail(rc);
  if (Q >= 2) {
    qs->wq.forceSpw = qs->wp.spw;
    qs->wq.allowForceGlobal = false;
    if (int rc = qs->wq.init(qcols + (size_t)first_site * qwpc, qwpc, Q, nsites, kSorted, qa_start, false, true, first_site, N)) return bail(rc);
    if (qs->wq.W != qs->wp.W) return bail(fail(ctx, PBWTGPU_ERR_CUDA, "qshard_open: query windows %d != panel windows %d", qs->wq.W, qs->wp.W));
    if (int rc = qs->wq.prepare()) return bail(rc);
  } else {
    if (int rc = qs->tmp.get(&qs->qaZero, (size_t)qs->wp.W + 1)) return bail(rc);
    if (cudaMemsetAsync(qs->qaZero, 0, sizeof(int32_t) * (qs->wp.W + 1), ctx->stream) != cudaSuccess) return bail(fail(ctx, PBWTGPU_ERR_CUDA, "qshard_open: memset"));
  }
  *out = qs;
  return 0;
}

int pbwtgpu_qshard_export_queries(pbwtgpu_ctx *ctx, pbwtgpu_qshard *qs, int32_t *dev_dst) {
  if (!ctx || !qs || !dev_dst) return fail(ctx, PBWTGPU_ERR_ARG, "qshard_export_queries: bad args");
  if (qs->Q >= 2) {
    if (int rc = qs->wq.export_aggregate(dev_dst)) return rc;
  } else {
    CUDA_TRY(ctx, cudaMemsetAsync(dev_dst, 0, sizeof(int32_t) * WindowedSweep::agg_ints(qs->Q), ctx->stream));
  }
  CUDA_TRY(ctx, stream_sync(ctx));
  return 0;
}

int pbwtgpu_qshard_prepare(pbwtgpu_ctx *ctx, pbwtgpu_qshard *qs, const int32_t *dev_query_aggs, int rank, int32_t *dev_panel_agg) {
  if (!ctx || !qs || rank < 0 || (rank > 0 && !dev_query_aggs) || !dev_panel_agg) return fail(ctx, PBWTGPU_ERR_ARG, "qshard_prepare: bad args");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (qs->Q >= 2) {
    if (int rc = qs->wq.apply_prefix(dev_query_aggs, rank)) return rc;
    qs->wp.qaWin = qs->wq.ord[qs->wq.cur];  // the query order entering every local window
  } else {
    qs->wp.qaWin = qs->qaZero;
  }
  if (int rc = qs->wp.prepare()) return rc;
  if (int rc = qs->wp.export_aggregate(dev_panel_agg)) return rc;
  CUDA_TRY(ctx, stream_sync(ctx));
  return 0;
}

int pbwtgpu_qshard_chain(pbwtgpu_ctx *ctx, pbwtgpu_qshard *qs, const int32_t *dev_panel_aggs, int rank, const int32_t *dev_state_in,
                         int32_t *dev_state_out) {
  if (!ctx || !qs || rank < 0 || (rank > 0 && (!dev_panel_aggs || !dev_state_in)) || !dev_state_out)
    return fail(ctx, PBWTGPU_ERR_ARG, "qshard_chain: bad args");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (int rc = qs->wp.apply_prefix(dev_panel_aggs, rank)) return rc;
  if (int rc = qs->wp.query_chain(rank > 0 ? dev_state_in : nullptr, dev_state_out)) return rc;
  CUDA_TRY(ctx, stream_sync(ctx));
  return 0;
}

int pbwtgpu_qshard_finish(pbwtgpu_ctx *ctx, pbwtgpu_qshard *qs, pbwtgpu_match **matches, int64_t *nmatch, pbwtgpu_nomatch **nomatch,
                          int64_t *nnomatch) {
  if (!ctx || !qs || !matches || !nmatch) return fail(ctx, PBWTGPU_ERR_ARG, "qshard_finish: bad args");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  pbwtgpu_match *recDev = nullptr;
  pbwtgpu_nomatch *nmDev = nullptr;
  unsigned long long nrec = 0, nnm = 0;
  if (int rc = finish_query_sweep(qs->wp, &recDev, &nrec, &nmDev, &nnm)) return rc;
  return query_results_to_host(ctx, recDev, nrec, nmDev, nnm, matches, nmatch, nomatch, nnomatch);
}

void pbwtgpu_qshard_free(pbwtgpu_ctx *ctx, pbwtgpu_qshard *qs) {
  if (!qs) return;
  if (ctx) stream_sync(ctx);
  delete qs;
}

int pbwtgpu_pack3(pbwtgpu_ctx *ctx, const uint8_t *y, int M, int N, uint8_t **yz, int64_t *nyz) {
  if (!ctx || !yz || !nyz || M < 1 || N < 0 || (N > 0 && !y)) return fail(ctx, PBWTGPU_ERR_ARG, "pack3: bad args");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  Temps tmp(ctx);
  const int wpc = pbwtgpu_words_per_col(M);
  uint8_t *yDev = nullptr, *yzDev = nullptr;
  uint32_t *ycols = nullptr;
  int32_t *colBytes = nullptr;
  int64_t *colBase = nullptr;
  if (int rc = tmp.get(&yDev, (size_t)M * std::max(N, 1))) return rc;
  if (int rc = tmp.get(&ycols, (size_t)wpc * std::max(N, 1))) return rc;
  if (int rc = tmp.get(&colBytes, (size_t)N + 1)) return rc;
  if (int rc = tmp.get(&colBase, (size_t)N + 1)) return rc;
  if (N) CUDA_TRY(ctx, cudaMemcpyAsync(yDev, y, (size_t)M * N, cudaMemcpyHostToDevice, ctx->stream));
  int64_t total = 0;
  {
    ScopedTimer t(ctx, PBWTGPU_K_PACK3);
    CUDA_TRY(ctx, bytes_to_bits(yDev, M, N, wpc, ycols, ctx->stream));
    CUDA_TRY(ctx, p3_count_columns(ycols, wpc, M, N, colBytes, colBase, nullptr, 0, ctx->stream));
    ctx->launches[PBWTGPU_K_PACK3] += 2 * (N > 0) + 1;
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(&total, colBase + N, sizeof total, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, stream_sync(ctx));
  if (int rc = tmp.get(&yzDev, (size_t)total + 16)) return rc;
  {
    ScopedTimer t(ctx, PBWTGPU_K_PACK3);
    CUDA_TRY(ctx, p3_emit_columns(ycols, wpc, M, N, colBase, yzDev, ctx->stream));
    ctx->launches[PBWTGPU_K_PACK3] += (N > 0);
  }
  HostGuard h(host_alloc_pinned((size_t)std::max<int64_t>(total, 1)));
  if (!h.p) return fail(ctx, PBWTGPU_ERR_NOMEM, "pinned host allocation failed (yz)");
  if (total) CUDA_TRY(ctx, cudaMemcpyAsync(h.p, yzDev, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, stream_sync(ctx));
  *yz = (uint8_t *)h.release();
  *nyz = total;
  return 0;
}

int pbwtgpu_unpack3(pbwtgpu_ctx *ctx, const uint8_t *yz, int64_t nyz, int M, int N, uint8_t *y, int32_t *zero_counts) {
  if (!ctx || !y || M < 1 || N < 0 || nyz < 0) return fail(ctx, PBWTGPU_ERR_ARG, "unpack3: bad args");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  Temps tmp(ctx);
  uint32_t *ycols = nullptr;
  int32_t *zc = nullptr;
  uint8_t *yDev = nullptr;
  if (int rc = tmp.get(&zc, (size_t)N + 1)) return rc;
  if (int rc = decode_stream(ctx, tmp, yz, nyz, M, N, &ycols, zc)) return rc;
  if (int rc = tmp.get(&yDev, (size_t)M * std::max(N, 1))) return rc;
  CUDA_TRY(ctx, bits_to_bytes(ycols, M, N, pbwtgpu_words_per_col(M), yDev, ctx->stream));
  if (N) CUDA_TRY(ctx, cudaMemcpyAsync(y, yDev, (size_t)M * N, cudaMemcpyDeviceToHost, ctx->stream));
  if (zero_counts && N)
    CUDA_TRY(ctx, cudaMemcpyAsync(zero_counts, zc, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, stream_sync(ctx));
  return 0;
}

void *pbwtgpu_host_alloc(size_t bytes) { return host_alloc_pinned(bytes); }

void pbwtgpu_free(void *p) {
  if (p && !host_release_pinned(p)) free(p);
}

void pbwtgpu_host_trim(void) {
  std::lock_guard<std::mutex> lk(gHostMu);
  for (auto &b : gHostFree) cudaFreeHost(b.first);
  gHostFree.clear();
}

}  // extern "C"
