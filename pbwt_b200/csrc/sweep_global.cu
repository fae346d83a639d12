// sweep_global.cu -- per-site forward sweep for panels too wide for one CTA's shared memory
// (M > 13,312 with d / 25,600 without; BASELINE configs 3-5: M = 1e5 .. 1e6).
//
// Same algorithm and outputs as sweep_smem.cu (pbwtCursorForwardsA/AD, the gather build, sorted-column
// sweeps: pbwtCore.c:458-557, pbwtIO.c:572-574), different residency: a[] and d[] live in global memory,
// double buffered, and stay L2-resident (16 B per haplotype: 16 MB at M = 1e6, L2 is 126 MB).
// A window of sites is swept by a GROUP of G co-resident CTAs (cooperative launch); CTA g of a group owns
// the contiguous sorted positions [g*chunk, (g+1)*chunk).  Per site:
//   phase A  each CTA: coalesced ld.cg of a[]/d[] (L1 must be bypassed: other CTAs rewrite them every
//            site), gather of the allele bit from the site's column (read-only, L1-cached), ballots,
//            warp-local zero counts and the run-max scan for d (as in sweep_smem.cu); d' staged in smem.
//   exchange each CTA publishes (zeros, run-max summary) with an epoch flag; every CTA reads all G partials
//            (flag spin, no barrier) and derives its incoming zero count / carry and the column's c.
//   phase C  stable-partition scatter of a (and d') into the other buffer.
//   barrier  one group-wide barrier per site (monotonic arrival counter in global memory).
// Windows are independent groups, so the window-parallel scheme of windows.cu applies unchanged.
//
// Match scans on wide panels (matchMaximalWithin, pbwtMatch.c:115-142; matchSequencesSweep, pbwtMatch.c:363-443):
// every CTA publishes its part of the sorted column y_k, the per-word maxima of d_k and (queries) the zero
// prefix, the group meets at a second barrier, and then each CTA runs the warp-cooperative scans of
// match_scan.cuh on state k in global memory (L2) -- maxWithin over the candidates among its own positions,
// the query sweep over its share of the queries (in query-PBWT order, one warp per query).  Records go to a
// per-CTA chunked stream; (site, CTA) segment tables put them into callback order afterwards (windows.cu).
#include <algorithm>
#include <type_traits>

#include "match_scan.cuh"
#include "sweep.cuh"

namespace pbwtgpu {
namespace {

constexpr int kThreads = 1024;
constexpr int kMaxChunk = 32768;  // sorted positions per CTA (d' staging: 128 KiB of shared memory)
constexpr int kBatch = 8;         // words per warp whose global loads are issued back to back

struct GroupSync {  // per group, in global memory
  unsigned int arrive;         // monotonic barrier counter
  unsigned int pad[31];
};

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int &total) {
  int x = v;
#pragma unroll
  for (int dl = 1; dl < 32; dl <<= 1) {
    const int t = __shfl_up_sync(kFull, x, dl);
    if (lane >= dl) x += t;
  }
  total = __shfl_sync(kFull, x, 31);
  return x - v;
}

__device__ __forceinline__ unsigned ld_volatile(const unsigned *p) {
  unsigned v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// ---- thread-block clusters: hardware barrier + distributed shared memory for groups of up to 16 CTAs ----
__device__ __forceinline__ void cluster_sync_all() {  // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_u32(const void *localSmem, unsigned ctaRank, uint32_t v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(localSmem)), "r"(ctaRank));
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(v) : "memory");
}

// run-max summary: bit 31 = "contains a run head", low 31 bits = max d since the last head (or overall)
__device__ __forceinline__ uint32_t rm_combine(uint32_t a, uint32_t b) {  // a earlier, b later
  return (b >> 31) ? b : ((a & 0x80000000u) | max(a & 0x7fffffffu, b));
}

constexpr int kQCap = 8192;  // queries per query sweep (their PBWT order is kept in every CTA's shared memory)

// state k in global memory (L2), for the scans of match_scan.cuh; other CTAs rewrite these arrays every site,
// so every load bypasses L1
struct GlobalState {
  const int32_t *aC, *dC;
  const uint32_t *yk;
  const int32_t *wm;
  int M, sentinel;  // d[0] = d[M] = k + 1
  // second level, one entry per 32 words (1024 positions); query sweep only
  const uint32_t *any1 = nullptr, *any0 = nullptr;  // bit j: word 32*e + j holds a one / a zero among its valid positions
  const int32_t *wm2 = nullptr;                     // max of wmax over words [32e, 32e + 32)
  __device__ __forceinline__ uint32_t anyw(int e, uint32_t x) const { return __ldcg((x ? any1 : any0) + e); }
  __device__ __forceinline__ int wmax2(int e) const { return __ldcg(wm2 + e); }
  // alleles equal to x among the valid positions of word w
  __device__ __forceinline__ uint32_t hits(int w, uint32_t x) const {
    const uint32_t v = __ldcg(yk + w);
    const int base = w << 5;
    const uint32_t vm = (M - base >= 32) ? kFull : (M > base ? ((1u << (M - base)) - 1u) : 0u);
    return (x ? v : ~v) & vm;
  }
  __device__ __forceinline__ int d(int i) const { return (i <= 0 || i >= M) ? sentinel : __ldcg(dC + i); }
  __device__ __forceinline__ uint32_t yw(int w) const { return __ldcg(yk + w); }
  __device__ __forceinline__ int wmax(int w) const { return __ldcg(wm + w); }
  __device__ __forceinline__ int a(int i) const { return __ldcg(aC + i); }
};

// First event at or after `from` walking down: a position p with d[p] > t (the block ends: returns false,
// r = p, or M), or before that a position with allele x (returns true, r = p).  pbwtMatch.c:381-383,400-402.
__device__ __forceinline__ bool first_allele_down(const GlobalState &s, int from, int t, uint32_t x, int lane, int &r) {
  const int M = s.M;
  int w = from >> 5;
  unsigned notBefore = kFull << (from & 31);
  for (;;) {
    if ((w << 5) >= M) { r = M; return false; }
    const int idx = (w << 5) + lane;
    const unsigned mB = __ballot_sync(kFull, idx >= M || s.d(idx) > t) & notBefore;
    const uint32_t yw = s.yw(w);
    const unsigned vm = (M - (w << 5) >= 32) ? kFull : ((1u << (M - (w << 5))) - 1u);
    const unsigned hits = (x ? yw : ~yw) & vm & notBefore;
    const unsigned ev = mB | hits;
    if (ev) {
      const int b = __ffs(ev) - 1;
      r = min((w << 5) + b, M);
      return !((mB >> b) & 1u);
    }
    // skip clean words, 32 at a time: a word is an event if it may hold a block end or the allele
    int ws = w + 1;
    for (;;) {
      const int wi = ws + lane;
      bool evw = (wi << 5) >= M;
      if (!evw) {
        const uint32_t y2 = s.yw(wi);
        evw = s.wmax(wi) > t || (x ? y2 : ~y2) != 0u;
      }
      const unsigned mw = __ballot_sync(kFull, evw);
      if (mw) { ws += __ffs(mw) - 1; break; }
      ws += 32;
    }
    w = ws;
    notBefore = kFull;
  }
}

// lowest position in [lo, hi) carrying allele x, or -1 (the "topmost hit" of an upward tier, pbwtMatch.c:394-396)
__device__ __forceinline__ int first_allele_in(const GlobalState &s, int lo, int hi, uint32_t x, int lane) {
  if (lo >= hi) return -1;
  const int w0 = lo >> 5, w1 = (hi - 1) >> 5;
  for (int wb = w0; wb <= w1; wb += 32) {
    const int w = wb + lane;
    uint32_t bits = 0u;
    if (w <= w1) {
      const uint32_t yw = s.yw(w);
      bits = x ? yw : ~yw;
      if (w == w0) bits &= kFull << (lo & 31);
      if (w == w1 && (hi & 31)) bits &= (1u << (hi & 31)) - 1u;
    }
    const unsigned m = __ballot_sync(kFull, bits != 0u);
    if (m) {
      const int l = __ffs(m) - 1;
      const uint32_t bw = __shfl_sync(kFull, bits, l);
      return ((wb + l) << 5) + __ffs(bw) - 1;
    }
  }
  return -1;
}

// nearest position <= from carrying allele x, or -1; two-level: word, its group of 32 words, then 32 groups per step
__device__ __forceinline__ int prev_allele(const GlobalState &s, int from, uint32_t x, int lane) {
  if (from < 0) return -1;
  const int w = from >> 5;
  {
    const uint32_t b = s.hits(w, x) & (kFull >> (31 - (from & 31)));
    if (b) return (w << 5) + 31 - __clz(b);
  }
  const int e = w >> 5;
  {  // earlier words of the same group
    const uint32_t m = s.anyw(e, x) & ((1u << (w & 31)) - 1u);
    if (m) {
      const int w2 = (e << 5) + 31 - __clz(m);
      return (w2 << 5) + 31 - __clz(s.hits(w2, x));
    }
  }
  for (int eb = e - 1; eb >= 0; eb -= 32) {
    const int ei = eb - lane;
    const uint32_t m = ei >= 0 ? s.anyw(ei, x) : 0u;
    const unsigned bal = __ballot_sync(kFull, m != 0u);
    if (bal) {
      const int l = __ffs(bal) - 1;
      const uint32_t mm = __shfl_sync(kFull, m, l);
      const int w2 = ((eb - l) << 5) + 31 - __clz(mm);
      return (w2 << 5) + 31 - __clz(s.hits(w2, x));
    }
  }
  return -1;
}

// nearest position >= from carrying allele x, or M
__device__ __forceinline__ int next_allele(const GlobalState &s, int from, uint32_t x, int lane) {
  const int M = s.M;
  if (from >= M) return M;
  const int w = from >> 5;
  {
    const uint32_t b = s.hits(w, x) & (kFull << (from & 31));
    if (b) return (w << 5) + __ffs(b) - 1;
  }
  const int e = w >> 5, nGroups = (((M + 31) >> 5) + 31) >> 5;
  {  // later words of the same group
    const uint32_t m = (w & 31) == 31 ? 0u : s.anyw(e, x) & (kFull << ((w & 31) + 1));
    if (m) {
      const int w2 = (e << 5) + __ffs(m) - 1;
      return (w2 << 5) + __ffs(s.hits(w2, x)) - 1;
    }
  }
  for (int eb = e + 1; eb < nGroups; eb += 32) {
    const int ei = eb + lane;
    const uint32_t m = ei < nGroups ? s.anyw(ei, x) : 0u;
    const unsigned bal = __ballot_sync(kFull, m != 0u);
    if (bal) {
      const int l = __ffs(bal) - 1;
      const uint32_t mm = __shfl_sync(kFull, m, l);
      const int w2 = ((eb + l) << 5) + __ffs(mm) - 1;
      return (w2 << 5) + __ffs(s.hits(w2, x)) - 1;
    }
  }
  return M;
}

// max of d over positions [lo, hi] (1 <= lo <= hi < M): partial words exactly, whole words through their maxima,
// whole groups of 32 words through the second level
__device__ __forceinline__ int range_max_d(const GlobalState &s, int lo, int hi, int lane) {
  const int w0 = lo >> 5, w1 = hi >> 5;
  int m = 0;
  {
    const int idx = (w0 << 5) + lane;
    if (idx >= lo && idx <= hi) m = s.d(idx);
  }
  if (w1 > w0) {
    const int idx = (w1 << 5) + lane;
    if (idx <= hi) m = max(m, s.d(idx));
    int wa = w0 + 1;
    const int wb = w1 - 1;  // whole words [wa, wb]
    const int ea = (wa + 31) >> 5, eb = ((wb + 1) >> 5) - 1;  // whole groups [ea, eb]
    if (ea <= eb) {
      for (int w = wa + lane; w < (ea << 5); w += 32) m = max(m, s.wmax(w));
      for (int e = ea + lane; e <= eb; e += 32) m = max(m, s.wmax2(e));
      wa = (eb + 1) << 5;
    }
    for (int w = wa + lane; w <= wb; w += 32) m = max(m, s.wmax(w));
  }
  return __reduce_max_sync(kFull, m);
}

// Extra arguments of the wide-panel kernel (all device pointers).
struct GlobalArgs {
  int G, chunk;
  int32_t *aBuf, *dBuf;
  uint32_t *partials;
  GroupSync *gsync;
  // published per site by every CTA, double buffered by site parity: [2][G*chunk/32]
  uint32_t *ySorted;
  int32_t *wmaxG;
  int32_t *zpre;      // query sweep: zeros before each word of the sorted column
  uint32_t *any1, *any0;  // query sweep, second level (one entry per 32 words): [2][G*chunk/1024] each
  int32_t *wmax2;
  // records: one chunked stream per CTA (chunk table row = blockIdx.x), ordered by the segment tables
  int32_t *segCnt;    // [(N+1)*G] records of (column, CTA); row N = the final state
  long long *segSrc;  // [(N+1)*G] offset of the segment in its CTA's stream
  // query sweep state, by query id
  int32_t *qf, *qe;
  unsigned long long *nmKeys;  // sort key (site * Q + query-PBWT position) of each "no match" event
};

// CLUSTER: the group is one thread-block cluster (G <= 16): the per-site exchange goes through distributed
// shared memory and the group barriers are hardware cluster barriers (release/acquire at cluster scope, which
// also orders the a[]/d[] stores to global memory) instead of flags and an arrival counter in global memory.
// MATCH: 0 none, 1 matchMaximalWithin, 2 matchLongWithin2 (-longWithin L)
template <int MODE, bool WITH_D, int MATCH, bool QUERY, bool CLUSTER>
__global__ void __launch_bounds__(kThreads, 1) sweep_global_kernel(const SweepArgs p, const GlobalArgs ga) {
  extern __shared__ __align__(16) uint32_t smem[];
  constexpr bool SCAN = MATCH != 0 || QUERY;
  const int G = ga.G, chunk = ga.chunk;
  const int chunkWords = chunk / 32;
  const int QW = QUERY ? ((p.Q + 31) / 32 + 31) / 32 * 32 : 0;  // words of a query column, padded
  uint32_t *ycolS = smem;                                               // [chunkWords]
  int32_t *dnS = reinterpret_cast<int32_t *>(smem + chunkWords);        // [chunk] (WITH_D)
  int32_t *wmaxS = dnS + (WITH_D ? chunk : 0);                          // [chunkWords] (SCAN)
  uint32_t *cmaskS = reinterpret_cast<uint32_t *>(wmaxS + (SCAN ? chunkWords : 0));  // [chunkWords] (MATCH)
  int32_t *zlocS = reinterpret_cast<int32_t *>(cmaskS + (MATCH ? chunkWords : 0));   // [chunkWords] (QUERY)
  int4 *resS = reinterpret_cast<int4 *>(zlocS + (QUERY ? chunkWords : 0));           // [kResMax] (SCAN)
  int32_t *ctabS = reinterpret_cast<int32_t *>(resS + (SCAN ? kResMax : 0));         // [kTabMirror] (SCAN)
  int32_t *nmS = ctabS + (SCAN ? kTabMirror : 0);                                    // [kResMax] (QUERY)
  uint32_t *qwS = reinterpret_cast<uint32_t *>(nmS + (QUERY ? kResMax : 0));         // [QW] query column
  int32_t *qzS = reinterpret_cast<int32_t *>(qwS + QW);                              // [QW] zeros before each word
  int32_t *qaS = qzS + QW;                                                           // [2][QW*32] query PBWT order
  // the same arrays as indices into smem[] for the hot loops (accesses through the derived pointers above are
  // compiled as generic loads/stores with a per-access shared-window conversion in cluster kernels)
  const int oY = 0, oDn = chunkWords, oWmax = oDn + (WITH_D ? chunk : 0), oZloc = oWmax + (SCAN ? chunkWords : 0) + (MATCH ? chunkWords : 0);
  // a-only builds: the prefix-array values a thread loads in phase A wait in shared memory for its phase C (the same
  // thread, the same slots) instead of being loaded from L2 a second time behind the group's exchange (config 3's pass 2
  // 66.0 -> 63.9 ms).  With d the extra 4 bytes per position cost more than they save: smaller chunks, more CTAs per
  // window, fewer windows (measured: 100k x 50k +8 %, 1M x 20k 2x slower).
  constexpr bool STAGE_A = MODE == kGather && !SCAN && !WITH_D;
  const int oAst = oWmax;  // [chunk] (STAGE_A; no scan arrays follow in these variants)
  __shared__ int wexS[32];      // zeros in this CTA's warps below each warp
  __shared__ int wsumZ[32];
  __shared__ uint32_t wsumR[32];
  __shared__ int wsumC[32];
  __shared__ int xZ, xC;        // zeros before this CTA, zeros in the whole column
  __shared__ uint32_t xR;       // run-max carry entering this CTA
  __shared__ uint32_t cpart[16][2];  // CLUSTER: (zeros, run-max summary) of every CTA of the group, written by its owner

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int win = blockIdx.x / G, g = blockIdx.x % G;
  const int M = p.M, wpc = p.wpc;
  const int nTotal = p.nTotal ? p.nTotal : p.N;
  const size_t MS = (size_t)G * chunk;  // padded state stride per buffer
  int32_t *aW = ga.aBuf + (size_t)win * 2 * MS;
  int32_t *dW = WITH_D ? ga.dBuf + (size_t)win * 2 * (MS + 4) : nullptr;
  uint32_t *part = ga.partials + (size_t)win * G * 4;  // per CTA: zeros, rm, flag, unused
  GroupSync *gs = ga.gsync + win;
  const int WG = G * chunkWords;  // words of the whole (padded) sorted column
  uint32_t *ySortedW = SCAN ? ga.ySorted + (size_t)win * 2 * WG : nullptr;
  int32_t *wmaxW = SCAN ? ga.wmaxG + (size_t)win * 2 * WG : nullptr;
  int32_t *zpreW = QUERY ? ga.zpre + (size_t)win * 2 * WG : nullptr;
  const int WG2 = WG / 32;
  uint32_t *any1W = QUERY ? ga.any1 + (size_t)win * 2 * WG2 : nullptr, *any0W = QUERY ? ga.any0 + (size_t)win * 2 * WG2 : nullptr;
  int32_t *wmax2W = QUERY ? ga.wmax2 + (size_t)win * 2 * WG2 : nullptr;

  int k0, k1, colShift;
  const uint32_t *src;
  if (p.perWindowSites) {
    k0 = p.kLabelBase; k1 = k0 + p.perWindowSites; colShift = -k0;
    src = p.src + (size_t)win * p.srcWindowStride;
  } else {
    k0 = p.siteOffset + win * p.sitesPerWindow; k1 = min(p.siteOffset + p.N, k0 + p.sitesPerWindow); colShift = -p.siteOffset;
    src = p.src;
  }
  const int p0 = g * chunk;                 // first sorted position of this CTA
  const int w0g = p0 / 32;                  // its first word of the sorted column
  const int EW = chunkWords / 32;           // words per warp
  const uint32_t ltMask = lanemask_lt(lane), laneBit = 1u << lane;
  unsigned int epoch = 0;                   // barriers passed so far

  auto group_barrier = [&]() {
    if (CLUSTER) {
      cluster_sync_all();
      return;
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(&gs->arrive, 1u);
      const unsigned target = (epoch + 1) * (unsigned)G;
      while (ld_volatile(&gs->arrive) < target) {}
      __threadfence();
    }
    ++epoch;
    __syncthreads();
  };

  RecSink sink;
  sink.pool = SCAN ? reinterpret_cast<int4 *>(p.recPool) : nullptr;
  sink.tab = SCAN && p.chunkTab ? p.chunkTab + (size_t)blockIdx.x * p.poolChunks : nullptr;
  sink.mirror = ctabS;
  sink.nChunks = 0;
  sink.cap = p.poolChunks;
  sink.hist = p.lenHist;
  long long recBase = 0;  // records produced so far by this CTA (uniform)
  auto ensure_chunks = [&](long long upTo) {
    const int need = (int)((upTo + kRecChunk - 1) >> kRecChunkLog);
    if (need > sink.nChunks) {
      if (tid == 0)
        for (int c2 = sink.nChunks; c2 < need; ++c2) {
          int id = atomicAdd(p.chunkCounter, 1);
          if (id >= p.poolChunks) id = -1;  // pool exhausted: records are dropped, the host reruns larger
          if (c2 < p.poolChunks) p.chunkTab[(size_t)blockIdx.x * p.poolChunks + c2] = id;  // (the table row ends there)
          ctabS[c2 & (kTabMirror - 1)] = id;
        }
      sink.nChunks = need;
      __syncthreads();
    }
  };
  // this CTA's share of the queries: positions [qLo, qHi) of the query PBWT order
  const int qPer = QUERY ? (p.Q + G - 1) / G : 0;
  const int qLo = min(p.Q, g * qPer), qHi = min(p.Q, qLo + qPer);
  int32_t *const qfW = QUERY ? ga.qf + (size_t)win * p.Q : nullptr;  // this window's query state, by query id
  int32_t *const qeW = QUERY ? ga.qe + (size_t)win * p.Q : nullptr;
  // queryMode 1: divergence tiers that lie before the window count as its first site (the window knows nothing earlier)
  const int qClamp = (QUERY && p.queryMode == 1) ? k0 : 0;
  if (QUERY) {
    const int32_t *qa0 = p.qaStart ? p.qaStart + (p.qaPerWindow ? (size_t)win * p.Q : 0) : nullptr;
    for (int i = tid; i < QW * 32; i += kThreads) qaS[i] = i < p.Q ? (qa0 ? qa0[i] : i) : 0;
    if (p.queryMode != 2)
      for (int i = qLo + tid; i < qHi; i += kThreads) {  // pbwtMatch.c:368-369: f and d start at zero (the window's first label)
        const int jj = qa0 ? qa0[i] : i;
        qfW[jj] = 0;
        qeW[jj] = k0;
      }
  }

  // records of state k (label kLabel, segment row) for the query share of this CTA: report lists, in rounds
  auto query_emit = [&](const GlobalState &st, int nR, int kLabel) {
    long long roundTotal = 0;
    for (int c0 = 0; c0 < nR; c0 += 32) {
      const int4 en0 = c0 + lane < nR ? resS[c0 + lane] : make_int4(0, 0, 0, 0);
      const int cntL = en0.y - en0.x;
      const int tot = __reduce_add_sync(kFull, cntL);
      if (sink.pool && tot) {
        const long long base = recBase + roundTotal;
        ensure_chunks(base + tot);
        const int mcnt = __shfl_sync(kFull, cntL, warp);
        if (mcnt > 0 && mcnt < kBigBlock) {
          const int4 en = resS[c0 + warp];
          const long long o = base + __reduce_add_sync(kFull, lane < warp ? cntL : 0);
          for (int t = lane; t < mcnt; t += 32) sink.put(o + t, make_int4(en.w, st.a(en.x + t), en.z, kLabel));
        }
        unsigned big = __ballot_sync(kFull, cntL >= kBigBlock);
        while (big) {
          const int j = __ffs(big) - 1;
          big &= big - 1;
          const int4 en = resS[c0 + j];
          const long long o = base + __reduce_add_sync(kFull, lane < j ? cntL : 0);
          for (int t = tid; t < en.y - en.x; t += kThreads) sink.put(o + t, make_int4(en.w, st.a(en.x + t), en.z, kLabel));
        }
      }
      roundTotal += tot;
    }
    recBase += roundTotal;
  };

  // ---- window start state ----
  for (int i = p0 + tid; i < p0 + chunk; i += kThreads) {
    if (i < M) {
      aW[i] = p.aStart ? p.aStart[(p.aStartShared ? 0 : (size_t)win * M) + i] : i;
      if (WITH_D) dW[i] = (i > 0 && p.dStart) ? p.dStart[(p.dStartShared ? 0 : (size_t)win * (M + 1)) + i] : 0;
    }
  }
  group_barrier();

  // debugging aid (option window_times): cycles CTA 0 spends in phase A / exchange / scatter+scans / end barrier
  // (accumulators in shared memory: held in registers they cost the hot loop ten registers and spills)
  __shared__ long long tAcc[5];
  const bool timing = p.winNanos != nullptr && blockIdx.x == 0 && tid == 0;
  auto tick = [&](int slot) {
    if (timing) {
      const long long now = clock64();
      tAcc[slot] += now - tAcc[4];
      tAcc[4] = now;
    }
  };
  if (timing) {
    tAcc[0] = tAcc[1] = tAcc[2] = tAcc[3] = 0;
    tAcc[4] = clock64();
  }

  int cur = 0;
  for (int k = k0; k < k1; ++k) {
    const uint32_t *col = src + (size_t)(k + colShift) * wpc;
    if (!WITH_D) asm volatile("" : "+l"(col));  // (keeps the pointer in registers: rematerialised, it costs a multiply-add per
                                                // gathered word; with d the registers are needed elsewhere)
    const int32_t *aC = aW + (size_t)cur * MS;
    int32_t *aN = aW + (size_t)(cur ^ 1) * MS;
    const int32_t *dC = WITH_D ? dW + (size_t)cur * (MS + 4) : nullptr;
    int32_t *dN = WITH_D ? dW + (size_t)(cur ^ 1) * (MS + 4) : nullptr;
    if (QUERY && tid < QW) qwS[tid] = tid < p.qwpc ? p.qcols[(size_t)(k + colShift) * p.qwpc + tid] : 0u;

    // ---------------- phase A ----------------
    const int wordBase = warp * EW;          // CTA-local word index of this warp's first word
    const int posBase = p0 + wordBase * 32;  // its first sorted position
    uint32_t pb;                             // allele just before this warp's first position
    if (posBase == 0 || posBase > M) pb = 2u;  // resolved at the first word: position 0 is a run head
    else if (MODE == kGather) {
      const int hp = __ldcg(aC + posBase - 1);
      pb = (__ldg(col + (hp >> 5)) >> (hp & 31)) & 1u;
    } else pb = __ldg(col + ((posBase - 1) >> 5)) >> 31;
    int zw = 0, carry = 0, lmax = 0;
    bool anyHead = false;
    int patchIdx = -1;  // CTA-local index (word*32+lane) of the warp's first run head, held by that lane only
    // kBatch words in flight per warp: the loads are L2 round trips.  FULL batches (entirely inside the warp's words
    // and below M: all but the last one) carry no validity predicates.
    // NB words per batch (kBatch, and 4 / 2 / 1 for what is left of the warp's EW words: a batch never reaches beyond them,
    // so no issue slot is spent on words that do not exist); FULL batches lie entirely below M and carry no predicates
    auto phaseA_batch = [&](auto fullTag, auto nbTag, const int e0) {
      constexpr bool FULL = decltype(fullTag)::value;
      constexpr int NB = decltype(nbTag)::value;
      int hB[NB], dB[NB];
      uint32_t cB[NB];
      // one 64-bit address per batch, the words of the batch at immediate offsets (an index per load costs four more
      // instructions of address arithmetic each)
      const int iBase = posBase + e0 * 32 + lane;
      const int32_t *aP = aC + iBase;
      const int32_t *dP = WITH_D ? dC + iBase : nullptr;
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        const bool valid = FULL || iBase + u * 32 < M;
        hB[u] = ((MODE == kGather || p.xOut) && valid) ? __ldcg(aP + u * 32) : 0;
        dB[u] = (WITH_D && valid) ? __ldcg(dP + u * 32) : 0;
        if (STAGE_A) smem[oAst + (wordBase + e0 + u) * 32 + lane] = (uint32_t)hB[u];
      }
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        const int base = posBase + (e0 + u) * 32;
        const bool valid = FULL || base + lane < M;
        if (MODE == kGather) cB[u] = valid ? __ldg(col + ((unsigned)hB[u] >> 5)) : kFull;
        else cB[u] = (FULL || base < M) ? __ldg(col + ((unsigned)base >> 5)) : 0u;
      }
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        const int e = e0 + u;
        const int i = posBase + e * 32 + lane;
        uint32_t Bw;
        if (MODE == kGather) {
          Bw = __ballot_sync(kFull, (cB[u] >> (hB[u] & 31)) & 1u);
        } else if (FULL) {
          Bw = cB[u];
        } else {
          const int base = posBase + e * 32;
          const uint32_t vm = (M - base >= 32) ? kFull : (M > base ? ((1u << (M - base)) - 1u) : 0u);
          Bw = cB[u] | ~vm;
        }
        if (MODE == kSorted && p.xOut && ((Bw >> lane) & 1u) && i < M)  // un-sort: x_k[a_k[i]] = y_k[i], reverse site order
          atomicOr(&p.xOut[(size_t)(p.N - 1 - (k + colShift)) * wpc + (hB[u] >> 5)], 1u << (hB[u] & 31));
        if (pb == 2u) pb = (~Bw) & 1u;
        const uint32_t H = Bw ^ ((Bw << 1) | pb);
        pb = Bw >> 31;
        if (lane == 0) smem[oY + wordBase + e] = Bw;
        if (QUERY && lane == 0) smem[oZloc + wordBase + e] = (uint32_t)zw;  // zeros of this warp before the word
        if (WITH_D) {
          int dv = dB[u];
          if (i == 0) dv = k + 1;  // sentinel d[0] (pbwtCore.c:414,507)
          if (SCAN) {  // word maximum of d: lets the match scans skip long blocks 1024 positions per step
            const int wm = __reduce_max_sync(kFull, dv);
            if (lane == 0) smem[oWmax + wordBase + e] = (uint32_t)wm;
          }
          int dnv = dv;
          if (H == 0u) {
            lmax = max(lmax, dv);
          } else {
            carry = max(carry, __reduce_max_sync(kFull, lmax));
            lmax = 0;
            int v = dv;
#pragma unroll
            for (int dl = 1; dl < 32; dl <<= 1) {
              const int t = __shfl_up_sync(kFull, v, dl);
              if (lane >= dl && ((H >> (lane - dl + 1)) & ((1u << dl) - 1u)) == 0u) v = max(v, t);
            }
            const int prev = __shfl_up_sync(kFull, v, 1);
            const bool headBelow = (H & ltMask) != 0u;
            const int prevMax = lane == 0 ? carry : (headBelow ? prev : max(carry, prev));
            if (H & laneBit) dnv = max(dv, prevMax);
            if (!anyHead) {
              anyHead = true;
              if (lane == __ffs(H) - 1) patchIdx = (wordBase + e) * 32 + lane;
            }
            carry = __shfl_sync(kFull, v, 31);
          }
          smem[oDn + (wordBase + e) * 32 + lane] = (uint32_t)dnv;
        }
        zw += __popc(~Bw);
      }
    };
    {
      auto run = [&](auto nbTag, const int e0) {
        if (posBase + (e0 + decltype(nbTag)::value) * 32 <= M) phaseA_batch(std::true_type{}, nbTag, e0);
        else phaseA_batch(std::false_type{}, nbTag, e0);
      };
      int e0 = 0;
      for (; e0 + kBatch <= EW; e0 += kBatch) run(std::integral_constant<int, kBatch>{}, e0);
      if (EW - e0 >= 4) { run(std::integral_constant<int, 4>{}, e0); e0 += 4; }
      if (EW - e0 >= 2) { run(std::integral_constant<int, 2>{}, e0); e0 += 2; }
      if (EW - e0 >= 1) run(std::integral_constant<int, 1>{}, e0);
    }
    if (WITH_D) carry = max(carry, __reduce_max_sync(kFull, lmax));
    if (lane == 0) {
      wsumZ[warp] = zw;
      if (WITH_D) wsumR[warp] = (anyHead ? 0x80000000u : 0u) | (uint32_t)carry;
    }
    __syncthreads();
    tick(0);

    // ---------------- exchange: CTA totals out, all partials of the group in ----------------
    int zExLane, zCta;
    zExLane = warp_excl_scan(wsumZ[lane], lane, zCta);
    uint32_t rInc = WITH_D ? wsumR[lane] : 0u;  // inclusive run-max scan over the CTA's warps
    if (WITH_D) {
#pragma unroll
      for (int dl = 1; dl < 32; dl <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, rInc, dl);
        if (lane >= dl) rInc = rm_combine(t, rInc);
      }
    }
    if (CLUSTER) {
      if (warp == 0) {
        const uint32_t rCta = __shfl_sync(kFull, rInc, 31);
        if (lane < G) {  // lane l delivers this CTA's totals to CTA l of the cluster
          st_cluster_u32(&cpart[g][0], (unsigned)lane, (uint32_t)zCta);
          st_cluster_u32(&cpart[g][1], (unsigned)lane, rCta);
        }
        wexS[lane] = zExLane;
      }
      cluster_sync_all();
      if (warp == 0) {
        const uint32_t z = lane < G ? cpart[lane][0] : 0u;
        const uint32_t r = lane < G ? cpart[lane][1] : 0u;
        const int zAll = __reduce_add_sync(kFull, (int)z);
        const int zBefore = __reduce_add_sync(kFull, lane < g ? (int)z : 0);
        uint32_t x = lane < g ? r : 0u;  // identity = (no head, 0)
        if (WITH_D) {
#pragma unroll
          for (int dl = 1; dl < 32; dl <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, x, dl);
            if (lane >= dl) x = rm_combine(t, x);
          }
          x = __shfl_sync(kFull, x, 31);
        }
        if (lane == 0) { xZ = zBefore; xC = zAll; xR = x; }
      }
    } else
    if (warp == 0) {
      const unsigned tag = (unsigned)(k - k0) + 1u;
      const uint32_t rCta = __shfl_sync(kFull, rInc, 31);
      if (lane == 0) {
        part[g * 4 + 0] = (uint32_t)zCta;
        part[g * 4 + 1] = rCta;
        __threadfence();
        atomicExch(&part[g * 4 + 2], tag);
      }
      // gather the group's partials (G <= 148 -> up to 5 per lane)
      int zBefore = 0, zAll = 0;
      uint32_t rBefore = 0u;  // combined summary of CTAs below g, in order
      for (int base = 0; base < G; base += 32) {
        const int gg = base + lane;
        uint32_t z = 0, r = 0;
        if (gg < G) {
          while (ld_volatile(&part[gg * 4 + 2]) != tag) {}
          __threadfence();
          z = ld_volatile(&part[gg * 4 + 0]);
          r = ld_volatile(&part[gg * 4 + 1]);
        }
        zAll += __reduce_add_sync(kFull, (int)z);
        zBefore += __reduce_add_sync(kFull, gg < g ? (int)z : 0);
        if (WITH_D) {  // ordered combine of the CTAs below g within this batch of 32
          uint32_t x = gg < g ? r : 0u;  // identity = (no head, 0)
#pragma unroll
          for (int dl = 1; dl < 32; dl <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, x, dl);
            if (lane >= dl) x = rm_combine(t, x);
          }
          rBefore = rm_combine(rBefore, __shfl_sync(kFull, x, 31));
        }
      }
      if (lane == 0) { xZ = zBefore; xC = zAll; xR = rBefore; }
      wexS[lane] = zExLane;
    }
    __syncthreads();
    tick(1);
    const int c = xC;
    const int zIn = xZ + __shfl_sync(kFull, zExLane, warp);
    int cIn = 0;
    if (WITH_D) {  // carry entering this warp: CTAs below, then this CTA's warps below
      const uint32_t below = __shfl_sync(kFull, rInc, warp > 0 ? warp - 1 : 0);
      const uint32_t comb = warp > 0 ? rm_combine(xR, below) : xR;
      cIn = (int)(comb & 0x7fffffffu);
    }

    const int par = (k - k0) & 1;
    if (SCAN) {  // publish this CTA's part of the sorted column, of the word maxima of d (and of the zero prefix)
      for (int w = tid; w < chunkWords; w += kThreads) {  // chunkWords is a multiple of 32: whole warps, whole groups
        ySortedW[(size_t)par * WG + w0g + w] = ycolS[w];
        wmaxW[(size_t)par * WG + w0g + w] = wmaxS[w];
        if (QUERY) {
          zpreW[(size_t)par * WG + w0g + w] = xZ + wexS[w / EW] + zlocS[w];
          const int base = (w0g + w) << 5;
          const uint32_t vm = (M - base >= 32) ? kFull : (M > base ? ((1u << (M - base)) - 1u) : 0u);
          const unsigned b1 = __ballot_sync(kFull, (ycolS[w] & vm) != 0u), b0 = __ballot_sync(kFull, (~ycolS[w] & vm) != 0u);
          const int m2 = __reduce_max_sync(kFull, wmaxS[w]);
          if (lane == 0) {
            const size_t e2 = (size_t)par * WG2 + ((w0g + w) >> 5);
            any1W[e2] = b1; any0W[e2] = b0; wmax2W[e2] = m2;
          }
        }
      }
      group_barrier();
    }

    // ---------------- phase C: scatter ----------------
    {
      int zacc = zIn;
      auto phaseC_batch = [&](auto fullTag, auto nbTag, const int e0) {
        constexpr bool FULL = decltype(fullTag)::value;
        constexpr int NB = decltype(nbTag)::value;
        int hB[NB];
        const int iBase = p0 + (wordBase + e0) * 32 + lane;
        const int32_t *aP = aC + iBase;
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          if (STAGE_A) hB[u] = (int)smem[oAst + (wordBase + e0 + u) * 32 + lane];
          else hB[u] = (FULL || iBase + u * 32 < M) ? __ldcg(aP + u * 32) : 0;
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          const int e = e0 + u;
          const int li = (wordBase + e) * 32 + lane;
          const int i = p0 + li;
          const uint32_t Bw = smem[oY + wordBase + e];
          const int z = zacc + __popc(~Bw & ltMask);
          const unsigned pos = (unsigned)((Bw & laneBit) ? c + i - z : z);
          if (FULL || i < M) aN[pos] = hB[u];
          if (WITH_D) {
            int dv = (int)smem[oDn + li];
            if (li == patchIdx) dv = max(dv, cIn);
            if (FULL || i < M) dN[pos] = dv;
          }
          zacc += __popc(~Bw);
        }
      };
      auto run = [&](auto nbTag, const int e0) {
        if (p0 + (wordBase + e0 + decltype(nbTag)::value) * 32 <= M) phaseC_batch(std::true_type{}, nbTag, e0);
        else phaseC_batch(std::false_type{}, nbTag, e0);
      };
      int e0 = 0;
      for (; e0 + kBatch <= EW; e0 += kBatch) run(std::integral_constant<int, kBatch>{}, e0);
      if (EW - e0 >= 4) { run(std::integral_constant<int, 4>{}, e0); e0 += 4; }
      if (EW - e0 >= 2) { run(std::integral_constant<int, 2>{}, e0); e0 += 2; }
      if (EW - e0 >= 1) run(std::integral_constant<int, 1>{}, e0);
    }
    if (p.yOut) {  // sorted column k, pad bits cleared
      for (int w = tid; w < chunkWords; w += kThreads) {
        const int gw = w0g + w;
        if (gw < wpc) {
          const int base = gw * 32;
          const uint32_t vm = (M - base >= 32) ? kFull : (M > base ? ((1u << (M - base)) - 1u) : 0u);
          p.yOut[(size_t)(k + colShift) * wpc + gw] = ycolS[w] & vm;
        }
      }
    }
    if (p.cOut && g == 0 && tid == 0) p.cOut[k + colShift] = c;

    if (MATCH) {
      // ---------------- matchMaximalWithin on state k: candidates among this CTA's positions ----------------
      const GlobalState st{aC, dC, ySortedW + (size_t)par * WG, wmaxW + (size_t)par * WG, M, k + 1};
      uint32_t myMask = 0;
      int tw = 0;
      for (int e = 0; e < EW; ++e) {
        const int lw = wordBase + e, gw = w0g + lw, base = gw * 32;
        if (base >= M) break;
        const uint32_t cw = ycolS[lw];
        const uint32_t pbit = gw == 0 ? (~cw & 1u) : (lw > 0 ? ycolS[lw - 1] >> 31 : st.yw(gw - 1) >> 31);
        const uint32_t nbit = gw + 1 >= WG ? 1u : (lw + 1 < chunkWords ? ycolS[lw + 1] & 1u : st.yw(gw + 1) & 1u);
        const uint32_t vm = (M - base >= 32) ? kFull : ((1u << (M - base)) - 1u);
        const uint32_t m = MATCH == 2 ? long_candidate_mask(st, cw, pbit, base, vm, k - p.longT, lane)
                                      : candidate_mask(st, cw, pbit, nbit, base, vm, lane);
        if (lane == e) myMask = m;
        tw += __popc(m);
      }
      if (lane < EW) cmaskS[wordBase + lane] = myMask;
      if (lane == 0) wsumC[warp] = p.debugNoCandidates ? 0 : tw;
      __syncthreads();
      const long long before = recBase;
      if (MATCH == 2) long_rounds(st, k, k - p.longT, M, EW, wsumC, cmaskS, resS, sink, recBase, ensure_chunks, tid, lane, warp, p0);
      else match_rounds(st, k, true, M, EW, wsumC, cmaskS, resS, sink, recBase, ensure_chunks, tid, lane, warp, p0);
      if (tid == 0) {
        const size_t seg = (size_t)(k + colShift) * G + g;
        ga.segCnt[seg] = (int32_t)(recBase - before);
        ga.segSrc[seg] = before;
      }
    }
    if (QUERY) {
      // ---------------- matchSequencesSweep on state k (pbwtMatch.c:376-426): one warp per query ----------------
      GlobalState st{aC, dC, ySortedW + (size_t)par * WG, wmaxW + (size_t)par * WG, M, k + 1};
      st.any1 = any1W + (size_t)par * WG2; st.any0 = any0W + (size_t)par * WG2; st.wm2 = wmax2W + (size_t)par * WG2;
      const int32_t *zk = zpreW + (size_t)par * WG;
      const int32_t *qaC = qaS + par * QW * 32;
      int32_t *qaN = qaS + (par ^ 1) * QW * 32;
      const long long before = recBase;
      for (int r0 = qLo; r0 < qHi; r0 += kResMax) {
        const int nR = min(kResMax, qHi - r0);
        // fast path, one query per thread: the allele at f extends the match -- nothing to report, e stays, f takes the
        // FM step (the common case by far: a query leaves its block only where its best match ends)
        if (tid < nR) {
          const int t = r0 + tid;
          const int jj = qaC[t];
          const uint32_t x = (qwS[t >> 5] >> (t & 31)) & 1u;
          const int f = __ldcg(qfW + jj);
          const int fw = f >> 5;
          const uint32_t yw = st.yw(fw);
          const bool fast = ((yw >> (f & 31)) & 1u) == x;
          if (fast) {
            const int Z = __ldcg(zk + fw) + __popc(~yw & lanemask_lt(f & 31));
            int nf = x ? c + f - Z : Z;
            if (nf == M) nf = 0;
            qfW[jj] = nf;
          }
          resS[tid] = make_int4(0, 0, 0, jj);
          nmS[tid] = fast ? 0 : -1;  // -1: needs the scans below
        }
        __syncthreads();
        for (int r = warp; r < nR; r += 32) {
          if (nmS[r] != -1) continue;  // (uniform)
          const int t = r0 + r;
          const int jj = qaC[t];
          const uint32_t x = (qwS[t >> 5] >> (t & 31)) & 1u;
          int f = __ldcg(qfW + jj), e = __ldcg(qeW + jj);
          int repLo = 0, repHi = 0, repStart = 0;
          bool nm = false;
          if (ybit(st, f) != x) {
            int hi;
            if (first_allele_down(st, f + 1, e, x, lane, hi)) {
              f = hi;  // an alternate of the same length that extends (pbwtMatch.c:381-383)
            } else {
              repLo = f; repHi = hi; repStart = e;  // report [f, hi) with start e (pbwtMatch.c:385)
              // Reseed (pbwtMatch.c:389-411).  The reference widens [f, hi) tier by tier of divergence, upper side
              // first on ties, until a tier holds allele x.  The tier that reaches a position p above f has value
              // max d(p, f], below hi it has max d[hi, p]; so the winner is decided by the nearest carriers u (above)
              // and v (below): up if max d(u, f] <= max d[hi, v].  Up: the new f is the topmost carrier of that whole
              // tier; down: v itself.  No carrier at all: the "no match" event.
              const int u = prev_allele(st, f - 1, x, lane);
              const int v = next_allele(st, hi, x, lane);
              if (u < 0 && v >= M) {
                nm = true;
                e = k + 1;
              } else {
                const int tUp = u >= 0 ? max(range_max_d(st, u + 1, f, lane), qClamp) : 0x7fffffff;
                const int tDown = v < M ? max(range_max_d(st, hi, v, lane), qClamp) : 0x7fffffff;
                if (tUp <= tDown) {
                  const int q = block_start(st, u, tUp, lane);
                  f = first_allele_in(st, q, u + 1, x, lane);
                  e = tUp;
                } else {
                  f = v;
                  e = tDown;
                }
              }
            }
          }
          // FM step (pbwtCursorCalculateU + pbwtCursorMap, pbwtCore.c:510-519, pbwt.h:130; pbwtMatch.c:417-423)
          const int fw = f >> 5;
          const int Z = __ldcg(zk + fw) + __popc(~st.yw(fw) & lanemask_lt(f & 31));
          int nf = x ? c + f - Z : Z;
          if (nf == M) nf = 0;
          __syncwarp();  // every lane has read nmS[r] (the test at the top of the loop) before lane 0 replaces it
          if (lane == 0) {
            qfW[jj] = nf;
            qeW[jj] = e;
            resS[r] = make_int4(repLo, repHi, repStart, jj);
            nmS[r] = nm ? (int)(2u | x) : 0;
          }
        }
        __syncthreads();
        if (p.queryMode != 1) query_emit(st, nR, k);
        if (p.nomatch && tid < nR && nmS[tid]) {  // rare: global append, sorted by (site, position) afterwards
          const unsigned long long o = atomicAdd(&p.qCounts[1], 1ull);
          if ((long long)o < p.nomatchCap) {
            p.nomatch[o] = pbwtgpu_nomatch{resS[tid].w, nmS[tid] & 1, k};
            ga.nmKeys[o] = (unsigned long long)(k + colShift) * (unsigned)p.Q + (unsigned)(r0 + tid);
          }
        }
        __syncthreads();  // resS / nmS are reused by the next round
      }
      if (tid == 0 && ga.segCnt) {
        const size_t seg = (size_t)(k + colShift) * G + g;
        ga.segCnt[seg] = (int32_t)(recBase - before);
        ga.segSrc[seg] = before;
      }
      // query cursor step (pbwtCursorForwardsRead on uq, pbwtMatch.c:426), redundantly in every CTA
      int v = 0, inc = 0;
      if (tid < QW) {
        const int base = tid * 32;
        const uint32_t vm = (p.Q - base >= 32) ? kFull : (p.Q > base ? ((1u << (p.Q - base)) - 1u) : 0u);
        v = __popc(~qwS[tid] & vm);
        inc = v;
#pragma unroll
        for (int dl = 1; dl < 32; dl <<= 1) {
          const int t = __shfl_up_sync(kFull, inc, dl);
          if (lane >= dl) inc += t;
        }
        if (lane == 31) wsumC[warp] = inc;
      }
      __syncthreads();
      int cq = 0, wbase = 0;
      for (int j = 0; j < QW / 32; ++j) {
        const int t = wsumC[j];
        if (j < warp) wbase += t;
        cq += t;
      }
      if (tid < QW) qzS[tid] = wbase + inc - v;
      __syncthreads();
      for (int t = tid; t < p.Q; t += kThreads) {
        const uint32_t w = qwS[t >> 5];
        const uint32_t x = (w >> (t & 31)) & 1u;
        const int zb = qzS[t >> 5] + __popc(~w & lanemask_lt(t & 31));
        qaN[x ? cq + t - zb : zb] = qaC[t];
      }
    }
    tick(2);
    group_barrier();
    tick(3);
    cur ^= 1;
  }
  if (timing)
    for (int i = 0; i < 4; ++i) p.winNanos[i] = tAcc[i];

  // ---- state k1 ----
  const int32_t *aE = aW + (size_t)cur * MS;
  const int32_t *dE = WITH_D ? dW + (size_t)cur * (MS + 4) : nullptr;
  const bool finalScan = !p.perWindowSites && k1 == nTotal && ((MATCH != 0 && p.reportAtN && (MATCH == 1 || k1 > k0)) || (QUERY && p.queryMode != 1));
  if (SCAN && finalScan) {  // word maxima of the final d, published to the group
    const int parE = (k1 - k0) & 1;
    for (int w = warp; w < chunkWords; w += 32) {
      const int i = p0 + w * 32 + lane;
      const int dv = i == 0 ? nTotal + 1 : (i < M ? __ldcg(dE + i) : 0);
      const int wm = __reduce_max_sync(kFull, dv);
      if (lane == 0) wmaxW[(size_t)parE * WG + w0g + w] = wm;
    }
    group_barrier();
    const GlobalState st{aE, dE, ySortedW + (size_t)parE * WG, wmaxW + (size_t)parE * WG, M, nTotal + 1};
    const long long before = recBase;
    if (MATCH == 2) {  // the reference scans state N with the stale alleles of site N-1 (pbwtCore.c:551-556)
      const GlobalState sl{aE, dE, ySortedW + (size_t)(parE ^ 1) * WG, wmaxW + (size_t)parE * WG, M, nTotal + 1};
      uint32_t myMask = 0;
      int tw = 0;
      for (int e = 0; e < EW; ++e) {
        const int gw = w0g + warp * EW + e, base = gw * 32;
        if (base >= M) break;
        const uint32_t cw = sl.yw(gw);
        const uint32_t vm = (M - base >= 32) ? kFull : ((1u << (M - base)) - 1u);
        const uint32_t m = long_candidate_mask(sl, cw, gw == 0 ? (~cw & 1u) : sl.yw(gw - 1) >> 31, base, vm, nTotal - p.longT, lane);
        if (lane == e) myMask = m;
        tw += __popc(m);
      }
      if (lane < EW) cmaskS[warp * EW + lane] = myMask;
      if (lane == 0) wsumC[warp] = tw;
      __syncthreads();
      long_rounds(sl, nTotal, nTotal - p.longT, M, EW, wsumC, cmaskS, resS, sink, recBase, ensure_chunks, tid, lane, warp, p0);
    }
    if (MATCH == 1) {  // k == N: nothing extends, every block is reported (alleles are not consulted when live == false)
      int tw = 0;
      if (lane < EW) {
        const int base = p0 + (warp * EW + lane) * 32;
        const uint32_t vm = (M - base >= 32) ? kFull : (M > base ? ((1u << (M - base)) - 1u) : 0u);
        cmaskS[warp * EW + lane] = vm;
        tw = __popc(vm);
      }
      tw = __reduce_add_sync(kFull, tw);
      if (lane == 0) wsumC[warp] = tw;
      __syncthreads();
      match_rounds(st, nTotal, false, M, EW, wsumC, cmaskS, resS, sink, recBase, ensure_chunks, tid, lane, warp, p0);
    }
    if (QUERY) {  // matches still open at N (pbwtMatch.c:430-436): a[f], then every following i with d[i] <= e
      const int32_t *qaC = qaS + parE * QW * 32;
      for (int r0 = qLo; r0 < qHi; r0 += kResMax) {
        const int nR = min(kResMax, qHi - r0);
        for (int r = warp; r < nR; r += 32) {
          const int jj = qaC[r0 + r];
          const int f = __ldcg(qfW + jj), e = __ldcg(qeW + jj);
          const int hi = block_end(st, f + 1, e, M, lane);
          if (lane == 0) resS[r] = make_int4(f, hi, e, jj);
        }
        __syncthreads();
        query_emit(st, nR, nTotal);
        __syncthreads();
      }
    }
    if (tid == 0 && ga.segCnt) {
      const size_t seg = (size_t)p.N * G + g;
      ga.segCnt[seg] = (int32_t)(recBase - before);
      ga.segSrc[seg] = before;
    }
  }

  // ---- window end ----
  const bool writeEnd = !p.endOnlyLast || k1 == nTotal;
  const size_t endSlot = p.endOnlyLast ? 0 : (size_t)win;
  if (p.aEnd && writeEnd)
    for (int i = p0 + tid; i < min(M, p0 + chunk); i += kThreads) p.aEnd[endSlot * M + i] = __ldcg(aE + i);
  if (WITH_D && p.dEnd && writeEnd) {
    for (int i = p0 + tid; i < min(M, p0 + chunk); i += kThreads)
      p.dEnd[endSlot * (M + 1) + i] = i == 0 ? k1 + 1 : __ldcg(dE + i);
    if (g == 0 && tid == 0) p.dEnd[endSlot * (M + 1) + M] = k1 + 1;
  }
}

}  // namespace

// dynamic shared memory: per-position arrays (chunk-proportional) + fixed part; colWords = staged bit column
static size_t smem_fixed_words(bool match, bool query, int Q, int colWords) {
  const bool scan = match || query;
  const size_t QW = query ? (size_t)((Q + 31) / 32 + 31) / 32 * 32 : 0;
  return (scan ? kResMax * 4 + kTabMirror : 0) + (query ? kResMax : 0) + QW * 2 + QW * 32 * 2 + (size_t)colWords;
}
static bool stages_a(SweepMode mode, bool with_d, bool match, bool query) { return mode == kGather && !with_d && !match && !query; }  // (STAGE_A of the kernel)
static size_t smem_bytes(int chunk, bool with_d, bool match, bool query, int Q, int colWords, bool stageA) {
  const size_t cw = chunk / 32;
  const bool scan = match || query;
  const size_t words = cw + (with_d ? chunk : 0) + (stageA ? chunk : 0) + (scan ? cw : 0) + (match ? cw : 0) + (query ? cw : 0) +
                       smem_fixed_words(match, query, Q, colWords);
  return words * 4;
}
constexpr size_t kSmemBudget = 232448 - 2048;  // 227 KB per CTA minus the static arrays
// largest chunk (multiple of 1024, <= kMaxChunk) whose arrays fit beside the fixed part; 0 = does not fit
static int max_chunk(bool with_d, bool match, bool query, int Q, int colWords, bool stageA) {
  const size_t fixed = smem_fixed_words(match, query, Q, colWords) * 4;
  if (fixed + 4096 > kSmemBudget) return 0;
  const bool scan = match || query;
  const double perPos = (with_d ? 4.0 : 0.0) + (stageA ? 4.0 : 0.0) + 4.0 * (1 + (scan ? 1 : 0) + (match ? 1 : 0) + (query ? 1 : 0)) / 32.0;
  long long c = (long long)((kSmemBudget - fixed) / perPos);
  c = c / 1024 * 1024;
  return (int)std::min<long long>(c, kMaxChunk);
}

template <bool CLUSTER>
static void *pick_kernel(SweepMode mode, bool with_d, int match, bool query) {
  if (query) return (void *)sweep_global_kernel<kSorted, true, 0, true, CLUSTER>;
  if (mode == kGather)
    return match == 2 ? (void *)sweep_global_kernel<kGather, true, 2, false, CLUSTER>
           : match    ? (void *)sweep_global_kernel<kGather, true, 1, false, CLUSTER>
           : with_d   ? (void *)sweep_global_kernel<kGather, true, 0, false, CLUSTER>
                      : (void *)sweep_global_kernel<kGather, false, 0, false, CLUSTER>;
  return match == 2 ? (void *)sweep_global_kernel<kSorted, true, 2, false, CLUSTER>
         : match    ? (void *)sweep_global_kernel<kSorted, true, 1, false, CLUSTER>
         : with_d   ? (void *)sweep_global_kernel<kSorted, true, 0, false, CLUSTER>
                    : (void *)sweep_global_kernel<kSorted, false, 0, false, CLUSTER>;
}

static int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }
static int pow2_floor(int v) { int p = 1; while (p * 2 <= v) p <<= 1; return p; }

// clusters of G CTAs that can be co-resident (0 = this cluster size cannot be launched)
static int max_active_clusters(void *kern, int G, size_t smemBytes) {
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes) != cudaSuccess) return 0;
  if (G > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(G);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smemBytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// Plan.  A window of sites is swept by G CTAs, each owning `chunk` (multiple of 1024, <= kMaxChunk) sorted
// positions; W windows run concurrently.  Panels up to 16 * kMaxChunk haplotypes use one thread-block cluster per
// window (hardware barriers, exchange through distributed shared memory); wider ones a cooperative grid of
// W * G <= maxCtas CTAs synchronised through global memory.
void sweep_global_plan(int M, int maxCtas, int wantWindows, SweepMode mode, bool with_d, bool match, bool query, int Q,
                       bool allowCluster, int *G, int *chunk, int *W, int *cluster) {
  // (staging the gather-mode bit column in shared memory by TMA was measured neutral to slower -- the kernel is issue
  // bound and the gather hits L1 -- and is not done)
  const int colWords = 0;
  const bool stageA = stages_a(mode, with_d, match, query);
  const int chunkMax = max_chunk(with_d, match, query, Q, colWords, stageA);
  const int gmin = (M + chunkMax - 1) / chunkMax;
  int w = wantWindows > 0 ? wantWindows : 1;
  *cluster = 0;
  if (allowCluster && gmin <= 16) {
    const int gLo = pow2_ceil(gmin), gHi = std::max(gLo, std::min(16, pow2_floor(std::max(1, (M + 1023) / 1024))));
    int g = std::min(gHi, std::max(gLo, pow2_floor(std::max(1, maxCtas / w))));
    for (; g >= gLo; g >>= 1) {
      const int ch = ((M + g - 1) / g + 1023) / 1024 * 1024;
      const int n = max_active_clusters(pick_kernel<true>(mode, with_d, match, query), g,
                                        smem_bytes(ch, with_d, match, query, Q, colWords, stageA));
      if (n >= 1) {
        *G = g; *chunk = ch; *W = std::min(w, n); *cluster = 1;
        return;
      }
    }
  }
  if (w * gmin > maxCtas) w = maxCtas / gmin;
  if (w < 1) w = 1;
  int g = maxCtas / w;  // use every SM: more CTAs per window shorten the per-site critical path
  if (g < gmin) g = gmin;
  int ch = ((M + g - 1) / g + 1023) / 1024 * 1024;
  if (ch > chunkMax) ch = chunkMax;
  g = (M + ch - 1) / ch;
  *G = g; *chunk = ch; *W = w;
}

size_t sweep_global_state_ints(int G, int chunk, int W, bool with_d) {
  const size_t MS = (size_t)G * chunk;
  return (size_t)W * 2 * MS + (with_d ? (size_t)W * 2 * (MS + 4) : 0);
}

size_t sweep_global_column_words(int G, int chunk, int W) { return (size_t)W * 2 * G * (chunk / 32); }

static int launch_any(void *kern, bool cluster, const SweepArgs &a, const GlobalBuffers &b, int ctas, size_t smemBytes,
                      cudaStream_t stream) {
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
  if (err != cudaSuccess) return (int)err;
  SweepArgs args = a;
  GlobalArgs ga{};
  ga.G = b.G; ga.chunk = b.chunk; ga.aBuf = b.aBuf; ga.dBuf = b.dBuf; ga.partials = b.partials;
  ga.gsync = static_cast<GroupSync *>(b.gsync);
  ga.ySorted = b.ySorted; ga.wmaxG = b.wmaxG; ga.zpre = b.zpre; ga.segCnt = b.segCnt; ga.segSrc = b.segSrc;
  ga.qf = b.qf; ga.qe = b.qe; ga.nmKeys = b.nmKeys;
  ga.any1 = b.any1; ga.any0 = b.any0; ga.wmax2 = b.wmax2;
  void *params[] = {&args, &ga};
  if (!cluster) return (int)cudaLaunchCooperativeKernel(kern, dim3(ctas), dim3(kThreads), params, smemBytes, stream);
  if (b.G > 8) {
    err = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (err != cudaSuccess) return (int)err;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = b.G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelExC(&cfg, kern, params);
}

int launch_sweep_global(const SweepArgs &a, SweepMode mode, bool with_d, int with_match, const GlobalBuffers &b,
                        cudaStream_t stream) {
  if (with_match && !with_d) return (int)cudaErrorInvalidValue;
  if (with_match == 2 && a.longT < 1) return (int)cudaErrorInvalidValue;
  if (b.cluster && b.G > 16) return (int)cudaErrorInvalidValue;
  void *kern = b.cluster ? pick_kernel<true>(mode, with_d, with_match, false) : pick_kernel<false>(mode, with_d, with_match, false);
  const int colWords = 0;
  return launch_any(kern, b.cluster, a, b, a.nWindows * b.G,
                    smem_bytes(b.chunk, with_d, with_match != 0, false, 0, colWords, stages_a(mode, with_d, with_match != 0, false)), stream);
}

int query_sweep_global_max_q() { return kQCap; }

int launch_query_sweep_global(const SweepArgs &a, const GlobalBuffers &b, cudaStream_t stream) {
  if (a.Q < 1 || a.Q > kQCap) return (int)cudaErrorInvalidValue;
  if (b.cluster && b.G > 16) return (int)cudaErrorInvalidValue;
  void *kern = b.cluster ? pick_kernel<true>(kSorted, true, false, true) : pick_kernel<false>(kSorted, true, false, true);
  SweepArgs args = a;
  if (args.nWindows < 1) args.nWindows = 1;
  return launch_any(kern, b.cluster, args, b, args.nWindows * b.G, smem_bytes(b.chunk, true, false, true, a.Q, 0, false), stream);
}

size_t sweep_global_sync_bytes(int W) { return sizeof(GroupSync) * (size_t)W; }

}  // namespace pbwtgpu
