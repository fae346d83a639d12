// sweep_smem.cu -- per-site forward sweep with the prefix array a[] and divergence array d[]
// resident in one CTA's shared memory (M <= ~13k with d, ~25k without).
//
// Replaces, for every site k of a window of sites:
//   a2/a3  pbwtCursorForwardsA / pbwtCursorForwardsAD          pbwtCore.c:458-508
//   a6     the gather y[j] = x[a[j]] + pbwtCursorWriteForwards pbwtIO.c:572-574, pbwtCore.c:573-585
//   a5     pbwtCursorForwardsReadAD (sorted columns as input)  pbwtCore.c:543-557
//   a10    matchMaximalWithin's per-site scan                  pbwtMatch.c:115-142
// The sorted column y_k is handed to the pack3 encoder (pack3.cu) as a bit column.
//
// Layout: 1024 threads = 32 warps; warp w owns E consecutive 32-position words of the sorted order,
// lane l of word j owns position 32*j + l ("striped"), so
//   * one __ballot_sync per word IS the sorted bit column y_k (no transposition),
//   * the stable-partition offsets are popcounts of ballots (no shuffles),
//   * shared-memory accesses to a[]/d[] are conflict-free for every E, addressed lane base + immediate
//     (the site loop is unrolled by two so the double-buffer offsets are compile-time constants).
// Positions M..1024E-1 are pad elements: their a[] points at a column bit that is always 1, so they
// sort to themselves forever and no per-element validity predicate is needed.
// Divergence update, single segmented max-scan: d_{k+1}[new(i)] = d_k[i] unless
// i is the head of a run of equal alleles, in which case it is max(d_k[i], max d_k over the previous
// run); the first two run heads see the k+1 sentinel in d_k[0].  The sentinels d[0], d[M] are
// (re)written by their owner threads at the top of each site.
// Column bits are staged by 1-D TMA bulk copies (cp.async.bulk + mbarrier), kStages sites ahead.
// Two __syncthreads per site.
// matchMaximalWithin: block boundaries are found with a two-level search (per-word maxima of d, then
// within the word), reporters are queued per warp, and blocks of >= kBigBlock records are written by
// all 32 warps together.
#include <type_traits>

#include "match_scan.cuh"
#include "sweep.cuh"

namespace pbwtgpu {
namespace {

constexpr int kThreads = 1024;
constexpr int kStages = 3;

template <int E, bool WITH_D, bool MATCH, bool QUERY = false, bool P3 = false>
struct SmemLayout {  // offsets in 32-bit words; every array start is 16-byte aligned
  static constexpr int MP = kThreads * E;
  static constexpr int WORDS = 32 * E;
  static constexpr int STAGE = WORDS + 4;  // + one always-ones word the pad elements point at
  static constexpr int a0 = 0;
  static constexpr int a1 = a0 + MP;
  static constexpr int d0 = a1 + MP;
  static constexpr int d1 = d0 + (WITH_D ? MP + 4 : 0);
  static constexpr int ycol = d1 + (WITH_D ? MP + 4 : 0);
  static constexpr int xcol = ycol + WORDS;
  static constexpr int xs = xcol + kStages * STAGE;  // un-sorted column being assembled (XOUT kernels)
  static constexpr int wsum = xs + WORDS;
  static constexpr int p3s = wsum + 3 * 32;  // fused pack3: run heads emitted per site (four rotating counters), and the
                                             // allele before each warp's first position [32]
  static constexpr int mbar = p3s + (P3 ? 4 + 32 : 0);
  static constexpr int wmax = mbar + 8;  // room for up to 4 mbarriers, keeps 16-byte alignment
  static constexpr int cmask = wmax + (MATCH ? WORDS : 0);  // per-word maxima of d (skip structure for long blocks)
  static constexpr int res = cmask + (MATCH ? WORDS : 0);       // per-word candidate masks, then result slots
  static constexpr int ctab = res + (MATCH ? kResMax * 4 : 0);  // mirror of the newest kTabMirror chunk ids
  static constexpr int QCAP = E <= 12 ? 1024 : 512;  // queries per launch (one thread each)
  static constexpr int qf = ctab + ((MATCH || QUERY) ? kTabMirror : 0);  // per-query f (top of longest match)
  static constexpr int qe = qf + (QUERY ? QCAP : 0);                      // per-query e (its start)
  static constexpr int qa0 = qe + (QUERY ? QCAP : 0);                     // query prefix array, double buffered
  static constexpr int qa1 = qa0 + (QUERY ? QCAP : 0);
  static constexpr int zloc = qa1 + (QUERY ? QCAP : 0);                   // warp-local zero prefix of each word
  static constexpr int wsq = zloc + (QUERY ? WORDS : 0);                  // 3 x 32 per-warp sums of the query phase
  static constexpr int totalWords = wsq + (QUERY ? 96 : 0);
  static constexpr size_t bytes = (size_t)totalWords * 4;
  static constexpr int padIndex = 32 * WORDS;  // "haplotype" whose allele is 1 at every site
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

// accessor of the shared-memory resident state for the scans of match_scan.cuh
struct SmemState {
  const int32_t *sA, *sD;
  const uint32_t *yc;
  const int32_t *wm;
  __device__ __forceinline__ int d(int i) const { return sD[i]; }
  __device__ __forceinline__ uint32_t yw(int w) const { return yc[w]; }
  __device__ __forceinline__ int wmax(int w) const { return wm[w]; }
  __device__ __forceinline__ int a(int i) const { return sA[i]; }
};
__device__ __forceinline__ uint32_t ybit(const uint32_t *yc, int j) { return (yc[j >> 5] >> (j & 31)) & 1u; }

// MATCH: 0 = none, 1 = matchMaximalWithin (pbwtMatch.c:115-142), 2 = matchLongWithin2 (-longWithin L, pbwtMatch.c:85-113)
// XOUT: also un-sort every column into p.xOut (sorted mode, no d, no matches: the forward half of pbwtBuildReverse)
// P3: run detection of the pack3 encoder (pbwtCore.c:240-277) fused into the sweep.  The sorted column y_k never goes to
// global memory; instead every lane that owns a word of it appends the word's run heads -- (position << 1 | allele), a few
// dozen per site -- to the window's run list, slots drawn from a shared-memory counter.  No cross-warp prefix, no extra
// barrier: anything more (byte counts need the NEXT head, offsets need a scan over the warps) was measured to sit on the
// site's critical path, because PBWT order concentrates a column's heads in a few warps (+0.6 ms on 4.1 ms for the full
// in-kernel encoder, against 0.4 ms for encoding the sorted columns afterwards).  pack3.cu sorts each site's few heads
// and formats the bytes from the run lists (about 3 % of the sorted columns' volume).
template <int E, int MODE, bool WITH_D, int MATCH, bool QUERY = false, bool XOUT = false, bool P3 = false>
__global__ void __launch_bounds__(kThreads, 1) sweep_smem_kernel(const SweepArgs p) {
  using L = SmemLayout<E, WITH_D, MATCH != 0, QUERY, P3>;
  extern __shared__ __align__(16) uint32_t smem[];
  int32_t *const sI = reinterpret_cast<int32_t *>(smem);
  int32_t *const wsumZ = sI + L::wsum;
  uint32_t *const wsumR = smem + L::wsum + 32;
  int32_t *const wsumC = sI + L::wsum + 64;
  uint64_t *const mbar = reinterpret_cast<uint64_t *>(smem + L::mbar);
  int32_t *const wmaxS = sI + L::wmax;
  uint32_t *const cmaskS = smem + L::cmask;
  int4 *const resS = reinterpret_cast<int4 *>(smem + L::res);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = p.M, wpc = p.wpc;
  const int nTotal = p.nTotal ? p.nTotal : p.N;
  const int win = blockIdx.x;
  long long tStart = 0;
  if (p.winNanos && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tStart));
  // site labels [k0, k1); column of label k is k + colShift of this window's source block
  int k0, k1, colShift;
  const uint32_t *src;
  if (p.perWindowSites) {
    k0 = p.kLabelBase;
    k1 = k0 + p.perWindowSites;
    colShift = -k0;
    src = p.src + (size_t)win * p.srcWindowStride;
  } else {
    window_range(win, p.sitesPerWindow, p.firstWindowSites, p.secondWindowSites, p.N, k0, k1);
    k0 += p.siteOffset;
    k1 += p.siteOffset;
    colShift = -p.siteOffset;
    src = p.src;
  }
  const uint32_t colBytes = (uint32_t)wpc * 4u;
  const int wordBase = warp * E;
  const int laneOff = wordBase * 32 + lane;  // position of (word 0 of this warp, this lane)
  int32_t *const sLane = sI + laneOff;
  const uint32_t ltMask = lanemask_lt(lane), laneBit = 1u << lane;
  // the thread owning position M (a pad element, or thread 0 when there is none) maintains d[M]
  const int ownerM = M < L::MP ? ((M >> 5) / E) * 32 + (M & 31) : 0;

  // ---- window start state (pbwtNakedCursorCreate, pbwtCore.c:402-418) ----
  for (int i = tid; i < L::MP; i += kThreads)
    sI[L::a0 + i] = i < M ? (p.aStart ? p.aStart[(p.aStartShared ? 0 : (size_t)win * M) + i] : i) : L::padIndex;
  if (WITH_D)
    for (int i = tid; i < L::MP + 4; i += kThreads) {
      int v = 0;
      if (i > 0 && i < M && p.dStart) v = p.dStart[(p.dStartShared ? 0 : (size_t)win * (M + 1)) + i];
      sI[L::d0 + i] = v;
      sI[L::d1 + i] = 0;
    }
  for (int i = tid; i < kStages * L::STAGE; i += kThreads) smem[L::xcol + i] = kFull;  // pad bits are ones
  if (XOUT)
    for (int i = tid; i < L::WORDS; i += kThreads) smem[L::xs + i] = 0u;
  if (QUERY)
    for (int i = tid; i < L::QCAP; i += kThreads) {
      sI[L::qf + i] = 0;  // pbwtMatch.c:368-369: f and d start at zero
      sI[L::qe + i] = 0;
      sI[L::qa0 + i] = i < p.Q ? (p.qaStart ? p.qaStart[i] : i) : 0;
    }
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&mbar[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    fence_proxy_async();
    for (int s = 0; s < kStages; ++s)
      if (k0 + s < k1) {
        mbar_expect_tx(&mbar[s], colBytes);
        tma_load_1d(smem + L::xcol + s * L::STAGE, src + (size_t)(k0 + s + colShift) * wpc, colBytes, &mbar[s]);
      }
  }

  int32_t *const p3Cnt = sI + L::p3s;  // run heads of the site being swept / of the three sites before it (rotating)
  uint32_t *const p3Pb = smem + L::p3s + 4;  // allele before each warp's first position at the site being swept
  uint32_t *const p3Dst = P3 ? p.p3Out + (size_t)win * p.p3Cap : nullptr;
  int p3Base = 0;  // run heads of this window so far (uniform)
  if (P3 && tid < 4) p3Cnt[tid] = 0;  // (published by the barrier that follows the mbarrier set-up)

  int32_t *const ctabS = sI + L::ctab;
  RecSink sink;
  sink.pool = (MATCH || QUERY) ? reinterpret_cast<int4 *>(p.recPool) : nullptr;
  sink.tab = (MATCH || QUERY) && p.chunkTab ? p.chunkTab + (size_t)win * p.poolChunks : nullptr;
  sink.mirror = ctabS;
  sink.nChunks = 0;
  sink.cap = p.poolChunks;
  sink.hist = p.lenHist;
  long long recBase = 0;  // records produced so far by this window (uniform)
  long long nmBase = 0;   // "no match" events so far (query sweep)
  uint32_t qWord = (QUERY && tid < p.Q && k0 < k1) ? p.qcols[(size_t)(k0 + colShift) * p.qwpc + warp] : 0u;

  // make sure chunks cover record offsets below upTo (uniform call; may contain a barrier)
  auto ensure_chunks = [&](long long upTo) {
    const int need = (int)((upTo + kRecChunk - 1) >> kRecChunkLog);
    if (need > sink.nChunks) {
      if (tid == 0)
        for (int c2 = sink.nChunks; c2 < need; ++c2) {
          int id = atomicAdd(p.chunkCounter, 1);
          if (id >= p.poolChunks) id = -1;  // pool exhausted: records are dropped, the host reruns larger
          if (c2 < p.poolChunks) p.chunkTab[(size_t)win * p.poolChunks + c2] = id;  // (the table row ends there)
          ctabS[c2 & (kTabMirror - 1)] = id;
        }
      sink.nChunks = need;
      __syncthreads();
    }
  };

  // candidates of a state -> records (match_scan.cuh); holds the caller's "S2" barrier
  auto match_rounds = [&](const int32_t *sA, const int32_t *sD, const uint32_t *yc, int k, bool live) {
    const SmemState st{sA, sD, yc, wmaxS};
    if (MATCH == 2) pbwtgpu::long_rounds(st, k, k - p.longT, M, E, wsumC, cmaskS, resS, sink, recBase, ensure_chunks, tid, lane, warp);
    else pbwtgpu::match_rounds(st, k, live, M, E, wsumC, cmaskS, resS, sink, recBase, ensure_chunks, tid, lane, warp);
  };

  // -------- one site; CUR selects which half of the double buffers holds state k --------
  auto site = [&](auto curTag, const int k) {
    constexpr int CUR = decltype(curTag)::value;
    constexpr int aC = CUR ? L::a1 : L::a0, aN = CUR ? L::a0 : L::a1;
    constexpr int dC = CUR ? L::d1 : L::d0, dN = CUR ? L::d0 : L::d1;
    const int it = k - k0;
    const int stage = it % kStages;
    mbar_wait(&mbar[stage], (uint32_t)(it / kStages) & 1u);
    const uint32_t *col = smem + L::xcol + stage * L::STAGE;
    uint32_t *yc = smem + L::ycol;
    if (WITH_D) {  // sentinels of state k (pbwtCore.c:414,507), written by the threads that read them below
      if (tid == 0) sI[dC] = k + 1;
      if (tid == ownerM) sI[dC + M] = k + 1;
    }

    // ---------------- phase A: gather, ballots, warp-local scans ----------------
    uint32_t B[E];
    int32_t dn[E];
    uint32_t pb = 0;  // allele just before this warp's first position
    if (warp > 0) {
      if (MODE == kGather) {
        const int hp = sI[aC + wordBase * 32 - 1];
        pb = (col[hp >> 5] >> (hp & 31)) & 1u;
      } else pb = col[wordBase - 1] >> 31;
    }
    const uint32_t pbFirst = pb;
    uint32_t nb = 1u;  // MATCH == 1: allele just after this warp's last position (pad positions carry ones); looked up here
                       // so that the two dependent loads are long done when the candidate test at the end of the phase asks
    if (MATCH == 1 && warp < 31) {
      if (MODE == kGather) {
        const int hn = sI[aC + (wordBase + E) * 32];
        nb = (col[hn >> 5] >> (hn & 31)) & 1u;
      } else nb = col[wordBase + E] & 1u;
    }
    const bool storeY = lane == 0 && (MATCH || QUERY || P3 || p.yOut != nullptr);  // the sorted column, when someone consumes it
    int carry = 0;         // max d since the last run head, over completed words (uniform)
    int lmax = 0;          // per-lane max of d over head-free words since then (reduced lazily)
    bool anyHead = false;  // has this warp seen a run head
    // the warp's first run head additionally needs the carry entering the warp (patched after S1)
    int patchLane = 0, pDn = 0, pIdx = 0, pZ = 0;
    uint32_t pY = 0;
    int zw = 0;  // zeros in this warp so far
    // MATCH == 1, held by lane e for word e of this warp: the word's run heads, and those among them whose upward scan is
    // known to extend (see candidate_mask)
    uint32_t myH = 0, myExt = 0;
    uint32_t longMask = 0;  // MATCH == 2: lane e holds the interior heads of word e
    int longCnt = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int h = sLane[aC + e * 32];
      uint32_t Bw;
      if (MODE == kGather) Bw = __ballot_sync(kFull, (col[h >> 5] >> (h & 31)) & 1u);
      else Bw = col[wordBase + e];
      B[e] = Bw;
      if (XOUT && ((Bw >> lane) & 1u) && h < M) atomicOr(&smem[L::xs + (h >> 5)], 1u << (h & 31));  // x_k[a_k[i]] = y_k[i]
      if (e == 0 && warp == 0) pb = (~Bw) & 1u;  // sorted position 0 is always a run head
      const uint32_t H = Bw ^ ((Bw << 1) | pb);
      pb = Bw >> 31;
      if (storeY) yc[wordBase + e] = Bw;
      if (WITH_D) {
        const int dv = sLane[dC + e * 32];
        if (MATCH) {  // word maximum of d: lets the match scans skip over long blocks 1024 positions per step
          const int wm = __reduce_max_sync(kFull, dv);
          if (lane == 0) wmaxS[wordBase + e] = wm;
        }
        if (MATCH == 2) {  // interior heads: the allele changes here and the match with the neighbour above is long
          const uint32_t m = H ? (H & __ballot_sync(kFull, dv <= k - p.longT)) : 0u;
          if (lane == e) longMask = m;
          longCnt += __popc(m);
        }
        if (H == 0u) {
          dn[e] = dv;
          lmax = max(lmax, dv);
        } else {
          if (MATCH == 1 && lane == e) myH = H;  // (head-free words pay nothing for the bookkeeping of match candidates)
          carry = max(carry, __reduce_max_sync(kFull, lmax));
          lmax = 0;
          int v = dv;  // inclusive max since the nearest head at or below this lane
#ifndef PBWTGPU_SCAN_BY_MASK
          // One distance to the nearest head at or below the lane replaces per-stage window masks (measured on a B200:
          // pass 2 of config 2 4.33 -> 4.24 ms, and the 16-byte spill of the E = 10 variant disappears).  "No head in
          // (lane - dl, lane]" is dl <= dist; lanes with dl > lane get their own value back from the shuffle, so they
          // need no separate guard.  tests/test_segscan_model.py proves both formulations equal the definition.
          const int dist = lane - (31 - __clz((int)(H & (ltMask | laneBit))));  // lane + 1 when there is none
#pragma unroll
          for (int dl = 1; dl < 32; dl <<= 1) {
            const int t = __shfl_up_sync(kFull, v, dl);
            if (dl <= dist) v = max(v, t);
          }
#else
#pragma unroll
          for (int dl = 1; dl < 32; dl <<= 1) {
            const int t = __shfl_up_sync(kFull, v, dl);
            if (lane >= dl && ((H >> (lane - dl + 1)) & ((1u << dl) - 1u)) == 0u) v = max(v, t);
          }
#endif
          const int prev = __shfl_up_sync(kFull, v, 1);
          const bool isHead = H & laneBit;
          const bool headBelow = (H & ltMask) != 0u;
          const int prevMax = lane == 0 ? carry : (headBelow ? prev : max(carry, prev));
          dn[e] = isHead ? max(dv, prevMax) : dv;
          // a run head whose d is not exceeded anywhere in the previous run: its match extends upward into the run before
          // that (candidate_mask); the warp's first head is left out, its previous run may begin in the warps below
          if (MATCH == 1) {
            const uint32_t em = __ballot_sync(kFull, isHead && prevMax <= dv && (anyHead || lane != __ffs(H) - 1));
            if (lane == e) myExt = em;
          }
          if (!anyHead) {
            anyHead = true;
            patchLane = __ffs(H) - 1;
            pDn = dn[e];
            pIdx = laneOff + e * 32;
            pZ = zw + __popc(~Bw & ltMask);
            pY = Bw & laneBit;
          }
          carry = __shfl_sync(kFull, v, 31);
        }
      }
      if (QUERY && lane == 0) sI[L::zloc + wordBase + e] = zw;  // zeros of this warp before the word
      zw += __popc(~Bw);
    }
    if (WITH_D) carry = max(carry, __reduce_max_sync(kFull, lmax));
    if (lane == 0) {
      wsumZ[warp] = zw;
      if (WITH_D) wsumR[warp] = (anyHead ? 0x80000000u : 0u) | (uint32_t)carry;
      if (P3) p3Pb[warp] = pbFirst;  // (read back after S2 of this site, rewritten after S2 of this site at the earliest)
    }
    if (MATCH == 2) {
      if (lane < E) cmaskS[wordBase + lane] = longMask;
      if (lane == 0) wsumC[warp] = longCnt;
    }
    if (MATCH == 1 && !anyHead && nb == pb) {  // no run boundary among this warp's positions (most warps, most sites):
      if (lane == 0) wsumC[warp] = 0;          // its masks are never looked at (locate_candidate only reads owners)
    } else if (MATCH == 1) {  // candidate lanes of this warp's words (exact: the allele after the warp is looked up too)
      // All words at once, lane e for word e: the heads and tails of runs (tails: the position before a head, which may be
      // the next word's or the next warp's first) are tested exactly as the reference would on its first step
      // (pbwtMatch.c:124-129; candidate_mask is the same test with one vote per word).
      const uint32_t hNext = __shfl_down_sync(kFull, myH, 1);
      uint32_t myMask = 0;
      if (lane < E) {
        const int base = (wordBase + lane) * 32;
        const uint32_t vm = (M - base >= 32) ? kFull : (M > base ? ((1u << (M - base)) - 1u) : 0u);
        const uint32_t headAfter = lane < E - 1 ? (hNext & 1u) : (nb ^ pb);  // pb: this warp's last allele
        const uint32_t Hn = (myH >> 1) | (headAfter << 31);
        uint32_t bm = (myH | Hn) & vm;
        while (bm) {
          const int l = __ffs(bm) - 1;
          bm &= bm - 1;
          // (the sentinels d[0] = d[M] = k + 1 are written at the top of the site by the threads that own those positions,
          // possibly of another warp: read by others only after S1, known here)
          const int pos = base + l;
          const int di = pos == 0 ? k + 1 : sI[dC + pos], dn1 = pos + 1 == M ? k + 1 : sI[dC + pos + 1];
          const bool samePrev = !((myH >> l) & 1u) || ((myExt >> l) & 1u), sameNext = !((Hn >> l) & 1u);
          if (!((di <= dn1 && samePrev) || (di >= dn1 && sameNext))) myMask |= 1u << l;
        }
        cmaskS[wordBase + lane] = myMask;
      }
      const int tw = __reduce_add_sync(kFull, __popc(myMask));
      if (lane == 0) wsumC[warp] = p.debugNoCandidates ? 0 : tw;
    }
    __syncthreads();  // S1: ballots, warp sums, match candidates published; everyone is done with col[]

    if (XOUT && tid < wpc) {  // the un-sorted column is complete: hand it over (reverse site order) and clear it
      p.xOut[(size_t)(p.N - 1 - (k + colShift)) * wpc + tid] = smem[L::xs + tid];
      smem[L::xs + tid] = 0u;
    }
    if (tid == 0 && k + kStages < k1) {  // refill the column slot just consumed
      fence_proxy_async();
      mbar_expect_tx(&mbar[stage], colBytes);
      tma_load_1d(smem + L::xcol + stage * L::STAGE, src + (size_t)(k + kStages + colShift) * wpc, colBytes, &mbar[stage]);
    }

    // ---------------- cross-warp prefixes (each warp redundantly) ----------------
    // zeros below this warp / in the column, and the divergence carry entering this warp: single warp reductions (REDUX)
    // over the 32 per-warp sums instead of 5-step shuffle scans
    int zIn, c, zExLane = 0;  // zExLane (query sweep only): zeros in warps below warp index == lane
    {
      const int z = wsumZ[lane];
      if (QUERY) {
        int tot;
        zExLane = warp_excl_scan(z, lane, tot);
        zIn = __shfl_sync(kFull, zExLane, warp);
        c = tot;
      } else {
        c = __reduce_add_sync(kFull, z);
        zIn = __reduce_add_sync(kFull, lane < warp ? z : 0);
      }
    }
    int cIn = 0;
    if (WITH_D) {  // max of the warps' carries from the last warp below this one that saw a run head (from warp 0 if none did)
      const uint32_t x = wsumR[lane];
      const unsigned heads = __ballot_sync(kFull, x >> 31) & lanemask_lt(warp);
      const int from = heads ? 31 - __clz(heads) : 0;
      cIn = __reduce_max_sync(kFull, (lane >= from && lane < warp) ? (int)(x & 0x7fffffffu) : 0);
    }

    // ---------------- phase C: stable partition scatter ----------------
    {
      int zacc = zIn;
      const int onesLane = c + laneOff;  // a one at position i goes to c + i - (zeros before i)
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const uint32_t Bw = B[e];
        const int z = zacc + __popc(~Bw & ltMask);
        const int pos = (Bw & laneBit) ? onesLane + e * 32 - z : z;
        sI[aN + pos] = sLane[aC + e * 32];  // re-read (base + immediate) instead of holding E registers
        if (WITH_D) sI[dN + pos] = dn[e];
        zacc += __popc(~Bw);
      }
      if (WITH_D && anyHead && lane == patchLane && cIn > pDn) {  // same thread wrote this slot above
        const int z = zIn + pZ;
        sI[dN + (pY ? c + pIdx - z : z)] = cIn;
      }
    }
    if (p.yOut && tid < wpc) {
      const int base = tid * 32;
      const uint32_t vm = (M - base >= 32) ? kFull : (M > base ? ((1u << (M - base)) - 1u) : 0u);
      p.yOut[(size_t)(k + colShift) * wpc + tid] = yc[tid] & vm;
    }
    if (p.cOut && tid == 0) p.cOut[k + colShift] = c;

    // ---------------- matchMaximalWithin on state k (holds the S2 barrier) ----------------
    if (MATCH) match_rounds(sI + aC, sI + dC, yc, k, true);
    else if (!QUERY) __syncthreads();  // S2: scatter complete
    // ---------------- fused pack3: this site's run heads to the window's run list ----------------
    // After S2, in the shadow of the match rounds' record emission (the warps that write records are dealt by candidate
    // number, the heads sit wherever the minor alleles are): every lane that owns a word of y_k appends the word's run
    // heads; slots come from a shared-memory counter, so there is no cross-warp prefix and no extra barrier.  The count
    // of site k is final at S1 of site k + 1 and booked one site later.
    if (P3) {
      if (it > 0) {
        const int cnt = p3Cnt[(it - 1) & 3];  // the site before: everyone appended before this site's S1
        if (tid == 0) {
          p.p3SiteOff[k - 1 + colShift] = p3Base;
          p.p3SiteCnt[k - 1 + colShift] = cnt;
          p3Cnt[(it + 2) & 3] = 0;  // read for the last time two barriers ago; the site after next appends to it
        }
        p3Base += cnt;
      }
      if (lane < E && (!WITH_D || anyHead)) {  // (with d, phase A already knows whether a run starts among this warp's positions)
        const int wd = wordBase + lane, base = wd * 32;
        if (base < M) {
          const uint32_t cw = yc[wd];  // (own words: the next site's phase A of this very warp overwrites them, later)
          const uint32_t pw = lane > 0 ? yc[wd - 1] >> 31 : (warp == 0 ? (~cw & 1u) : p3Pb[warp]);  // position 0 always heads a run
          uint32_t h = cw ^ ((cw << 1) | pw);
          if (M - base < 32) h &= (1u << (M - base)) - 1u;  // pad positions (all ones) are not part of the column
          if (h) {
            int o = p3Base + atomicAdd(&p3Cnt[it & 3], __popc(h));
            do {
              const int j = __ffs(h) - 1;
              h &= h - 1;
              if (o < p.p3Cap) p3Dst[o] = ((uint32_t)(base + j) << 1) | ((cw >> j) & 1u);
              else *p.p3Overflow = 1;
              ++o;
            } while (h);
          }
        }
      }
      __syncwarp();  // lane 0 rewrites this warp's words of yc[] in the next site's phase A: order the reads above before it
                     // (the vote that precedes that store converges the warp but is no memory barrier; racecheck)
    }
    if (QUERY) {
      // ---------------- matchSequencesSweep on state k (pbwtMatch.c:376-426), one thread per query ----------------
      constexpr int qaC = CUR ? L::qa1 : L::qa0, qaN = CUR ? L::qa0 : L::qa1;
      const int32_t *sD = sI + dC;
      const bool active = tid < p.Q;
      const uint32_t x = active ? (qWord >> lane) & 1u : 1u;  // query allele at sorted query position tid
      if (active && k + 1 < k1) qWord = p.qcols[(size_t)(k + 1 + colShift) * p.qwpc + warp];  // prefetch next site
      const int jj = active ? sI[qaC + tid] : 0;
      int f = sI[L::qf + jj], e = sI[L::qe + jj];
      int repLo = 0, repHi = 0, repStart = 0;
      bool nm = false;
      if (active && ybit(yc, f) != x) {
        // an alternate of the same length further down that does extend? (pbwtMatch.c:381-383)
        int hi = f + 1;
        bool moved = false;
        while (hi < M && sD[hi] <= e) {
          if (ybit(yc, hi) == x) { f = hi; moved = true; break; }
          ++hi;
        }
        if (!moved) {
          repLo = f; repHi = hi; repStart = e;  // report [f, hi) with start e (pbwtMatch.c:385)
          // reseed: widen alternately up and down by divergence tiers (pbwtMatch.c:389-411)
          int lo = f;
          int dDown = hi < M ? sD[hi] : k;
          int dUp = sD[lo];
          for (;;) {
            if (dUp <= dDown) {
              int top = -1;
              while (lo > 0 && sD[lo] <= dUp) {
                --lo;
                if (ybit(yc, lo) == x) top = lo;
              }
              if (top >= 0) { f = top; e = dUp; break; }
              dUp = sD[lo];
            } else {
              bool hit = false;
              while (hi < M && sD[hi] <= dDown) {
                if (ybit(yc, hi) == x) { f = hi; e = dDown; hit = true; break; }
                ++hi;
              }
              if (hit) break;
              dDown = hi == M ? k : sD[hi];
              if (lo == 0 && hi == M) { nm = true; e = k + 1; break; }  // allele absent from the column
            }
          }
        }
      }
      {  // FM step: f -> position in state k+1 (pbwtCursorCalculateU + pbwtCursorMap, pbwtCore.c:510-519, pbwt.h:130)
        const int fw = f >> 5;
        const int Z = __shfl_sync(kFull, zExLane, fw / E) + sI[L::zloc + fw] + __popc(~yc[fw] & lanemask_lt(f & 31));
        int nf = x ? c + f - Z : Z;
        if (nf == M) nf = 0;  // pbwtMatch.c:421-422
        if (active) { sI[L::qf + jj] = nf; sI[L::qe + jj] = e; }
      }
      // per-warp sums: records, "no match" events, zeros of the query column
      const int cnt = repHi - repLo;
      int cntTot;
      const int cntEx = warp_excl_scan(cnt, lane, cntTot);
      const unsigned nmBal = __ballot_sync(kFull, nm), qzBal = __ballot_sync(kFull, active && x == 0u);
      int32_t *wsq = sI + L::wsq;
      if (lane == 0) { wsq[warp] = cntTot; wsq[32 + warp] = __popc(nmBal); wsq[64 + warp] = __popc(qzBal); }
      __syncthreads();  // S2: scatter complete; query sums published
      int recTot, nmTot, cq;
      const int recEx = __shfl_sync(kFull, warp_excl_scan(wsq[lane], lane, recTot), warp);
      const int nmEx = __shfl_sync(kFull, warp_excl_scan(wsq[32 + lane], lane, nmTot), warp);
      const int qzEx = __shfl_sync(kFull, warp_excl_scan(wsq[64 + lane], lane, cq), warp);
      if (sink.pool && recTot) {
        ensure_chunks(recBase + recTot);
        long long o = recBase + recEx + cntEx;
        for (int i = repLo; i < repHi; ++i, ++o) sink.put(o, make_int4(jj, sI[aC + i], repStart, k));
      }
      if (nm && p.nomatch) {
        const long long o = nmBase + nmEx + __popc(nmBal & ltMask);
        if (o < p.nomatchCap) p.nomatch[o] = pbwtgpu_nomatch{jj, (int)x, k};
      }
      recBase += recTot;
      nmBase += nmTot;
      if (active) {  // query cursor step (pbwtCursorForwardsRead on uq, pbwtMatch.c:426)
        const int zb = qzEx + __popc(qzBal & ltMask);
        sI[qaN + (x ? cq + tid - zb : zb)] = jj;
      }
    }
  };

  int k = k0;
  for (; k + 1 < k1; k += 2) {
    site(std::integral_constant<int, 0>{}, k);
    site(std::integral_constant<int, 1>{}, k + 1);
  }
  int cur = 0;
  if (k < k1) {
    site(std::integral_constant<int, 0>{}, k);
    cur = 1;
  }

  // ---------------- window end: state k1 ----------------
  const int aC = cur ? L::a1 : L::a0;
  const int dC = cur ? L::d1 : L::d0;
  if (WITH_D && tid == 0) {
    sI[dC] = k1 + 1;
    sI[dC + M] = k1 + 1;
  }
  __syncthreads();
  const bool writeEnd = !p.endOnlyLast || k1 == nTotal;
  const size_t endSlot = p.endOnlyLast ? 0 : (size_t)win;
  if (p.aEnd && writeEnd)
    for (int i = tid; i < M; i += kThreads) p.aEnd[endSlot * M + i] = sI[aC + i];
  if (WITH_D && p.dEnd && writeEnd)
    for (int i = tid; i <= M; i += kThreads) p.dEnd[endSlot * (M + 1) + i] = sI[dC + i];

  if (QUERY) {  // matches still open at N (pbwtMatch.c:430-436): a[f], then every following i with d[i] <= e
    const int32_t *sD = sI + dC;
    const bool active = tid < p.Q;
    const int jj = active ? sI[(cur ? L::qa1 : L::qa0) + tid] : 0;
    const int f = sI[L::qf + jj], e = sI[L::qe + jj];
    int hi = f + 1;
    if (active) while (hi < M && sD[hi] <= e) ++hi;
    const int cnt = active ? hi - f : 0;
    int cntTot;
    const int cntEx = warp_excl_scan(cnt, lane, cntTot);
    int32_t *wsq = sI + L::wsq;
    if (lane == 0) wsq[warp] = cntTot;
    __syncthreads();
    int recTot;
    const int recEx = __shfl_sync(kFull, warp_excl_scan(wsq[lane], lane, recTot), warp);
    if (sink.pool && recTot) {
      ensure_chunks(recBase + recTot);
      long long o = recBase + recEx + cntEx;
      for (int i = f; i < f + cnt; ++i, ++o) sink.put(o, make_int4(jj, sI[aC + i], e, nTotal));
    }
    recBase += recTot;
    if (tid == 0 && p.qCounts) { p.qCounts[0] = (unsigned long long)recBase; p.qCounts[1] = (unsigned long long)nmBase; }
  }
  if (MATCH && p.reportAtN && !p.perWindowSites && k1 == nTotal) {  // k == N: nothing extends, every block is reported
    const uint32_t *yc = smem + L::ycol;                           // alleles are not consulted when live == false
    for (int w = warp; w < L::WORDS; w += 32) {  // word maxima of the final d
      const int wm = __reduce_max_sync(kFull, sI[dC + w * 32 + lane]);
      if (lane == 0) wmaxS[w] = wm;
    }
    int tw = 0;
    if (MATCH == 2) {  // the reference scans state N with the stale alleles of site N-1 (pbwtCore.c:551-556): yc still holds them
      const SmemState st{sI + aC, sI + dC, yc, wmaxS};
      uint32_t myMask = 0;
      for (int e = 0; e < E; ++e) {
        const int wd = wordBase + e, base = wd * 32;
        if (base >= M) break;
        const uint32_t vm = (M - base >= 32) ? kFull : ((1u << (M - base)) - 1u);
        const uint32_t m = long_candidate_mask(st, yc[wd], wd > 0 ? yc[wd - 1] >> 31 : (~yc[wd] & 1u), base, vm, nTotal - p.longT, lane);
        if (lane == e) myMask = m;
        tw += __popc(m);
      }
      if (lane < E) cmaskS[wordBase + lane] = myMask;
    } else if (lane < E) {  // every valid position is a candidate
      const int base = (wordBase + lane) * 32;
      const uint32_t vm = (M - base >= 32) ? kFull : (M > base ? ((1u << (M - base)) - 1u) : 0u);
      cmaskS[wordBase + lane] = vm;
      tw = __popc(vm);
    }
    if (MATCH != 2) tw = __reduce_add_sync(kFull, tw);
    if (lane == 0) wsumC[warp] = tw;
    __syncthreads();
    match_rounds(sI + aC, sI + dC, yc, nTotal, false);
  }
  if (MATCH && p.recCount && tid == 0) p.recCount[win] = (unsigned long long)recBase;
  if (P3 && tid == 0) {  // the last site's heads (everyone appended before the barrier at the window's end)
    int total = p3Base;
    if (k1 > k0) {
      const int cnt = p3Cnt[(k1 - k0 - 1) & 3];
      p.p3SiteOff[k1 - 1 + colShift] = p3Base;
      p.p3SiteCnt[k1 - 1 + colShift] = cnt;
      total += cnt;
    }
    p.p3Count[win] = total;
  }
  if (p.winNanos && tid == 0) {
    long long tEnd;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tEnd));
    p.winNanos[win] = tEnd - tStart;
  }
}

template <int E, int MODE, bool WITH_D, int MATCH, bool P3 = false>
int launch_one(const SweepArgs &a, cudaStream_t stream) {
  using L = SmemLayout<E, WITH_D, MATCH != 0, false, P3>;
  static_assert(L::bytes <= 232448, "shared memory layout exceeds the 227 KB per-CTA limit");
  static_assert(kStages <= 4 && L::xcol % 4 == 0 && L::mbar % 2 == 0 && L::res % 4 == 0 && L::d0 % 4 == 0 && L::d1 % 4 == 0,
                "shared memory arrays must keep their alignment");
  auto kern = sweep_smem_kernel<E, MODE, WITH_D, MATCH, false, false, P3>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes);
  if (err != cudaSuccess) return (int)err;
  kern<<<a.nWindows, kThreads, L::bytes, stream>>>(a);
  return (int)cudaGetLastError();
}

template <int MODE, bool WITH_D, int MATCH, bool P3 = false>
int launch_e(const SweepArgs &a, cudaStream_t stream) {
  const int words = (a.M + 31) / 32;
  const int e = (words + 31) / 32;
  if (e <= 1) return launch_one<1, MODE, WITH_D, MATCH, P3>(a, stream);
  if (e <= 2) return launch_one<2, MODE, WITH_D, MATCH, P3>(a, stream);
  if (e <= 4) return launch_one<4, MODE, WITH_D, MATCH, P3>(a, stream);
  if (e <= 7) return launch_one<7, MODE, WITH_D, MATCH, P3>(a, stream);
  if (e <= 10) return launch_one<10, MODE, WITH_D, MATCH, P3>(a, stream);
  if (e <= 13) return launch_one<13, MODE, WITH_D, MATCH, P3>(a, stream);
  if constexpr (!WITH_D) {
    if (e <= 19) return launch_one<19, MODE, WITH_D, MATCH, P3>(a, stream);
    if (e <= 25) return launch_one<25, MODE, WITH_D, MATCH, P3>(a, stream);
  }
  return (int)cudaErrorInvalidValue;
}

}  // namespace

int sweep_smem_max_m(bool with_d) { return with_d ? 13 * 1024 : 25 * 1024; }
int query_sweep_max_m() { return 13 * 1024; }
int query_sweep_max_q(int M) { return M <= 10 * 1024 ? 1024 : 512; }

template <int E>
static int launch_query_one(const SweepArgs &a, cudaStream_t stream) {
  using L = SmemLayout<E, true, false, true>;
  static_assert(L::bytes <= 232448, "shared memory layout exceeds the 227 KB per-CTA limit");
  auto kern = sweep_smem_kernel<E, kSorted, true, 0, true>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes);
  if (err != cudaSuccess) return (int)err;
  kern<<<1, kThreads, L::bytes, stream>>>(a);
  return (int)cudaGetLastError();
}

int launch_query_sweep_smem(const SweepArgs &a, cudaStream_t stream) {
  if (a.M < 1 || a.M > query_sweep_max_m() || a.Q < 1 || a.Q > query_sweep_max_q(a.M)) return (int)cudaErrorInvalidValue;
  const int e = ((a.M + 31) / 32 + 31) / 32;
  if (e <= 1) return launch_query_one<1>(a, stream);
  if (e <= 2) return launch_query_one<2>(a, stream);
  if (e <= 4) return launch_query_one<4>(a, stream);
  if (e <= 7) return launch_query_one<7>(a, stream);
  if (e <= 10) return launch_query_one<10>(a, stream);
  return launch_query_one<13>(a, stream);
}

template <int E>
static int launch_xout_one(const SweepArgs &a, cudaStream_t stream) {
  using L = SmemLayout<E, false, false>;
  auto kern = sweep_smem_kernel<E, kSorted, false, 0, false, true>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes);
  if (err != cudaSuccess) return (int)err;
  kern<<<a.nWindows, kThreads, L::bytes, stream>>>(a);
  return (int)cudaGetLastError();
}

static int launch_xout(const SweepArgs &a, cudaStream_t stream) {
  const int e = ((a.M + 31) / 32 + 31) / 32;
  if (e <= 1) return launch_xout_one<1>(a, stream);
  if (e <= 2) return launch_xout_one<2>(a, stream);
  if (e <= 4) return launch_xout_one<4>(a, stream);
  if (e <= 7) return launch_xout_one<7>(a, stream);
  if (e <= 10) return launch_xout_one<10>(a, stream);
  if (e <= 13) return launch_xout_one<13>(a, stream);
  if (e <= 19) return launch_xout_one<19>(a, stream);
  return launch_xout_one<25>(a, stream);
}

int launch_sweep_smem(const SweepArgs &a, SweepMode mode, bool with_d, int match_mode, cudaStream_t stream) {
  if (a.M < 1 || a.M > sweep_smem_max_m(with_d || match_mode)) return (int)cudaErrorInvalidValue;
  if (a.xOut) {
    if (mode != kSorted || with_d || match_mode) return (int)cudaErrorInvalidValue;
    return launch_xout(a, stream);
  }
  if (match_mode == 2 && a.longT < 1) return (int)cudaErrorInvalidValue;
  if (a.p3Out) {  // pack3 run detection fused into the sweep: builds from panel columns
    if (mode != kGather || match_mode == 2 || !a.p3Count || !a.p3SiteOff || !a.p3SiteCnt || !a.p3Overflow || a.p3Cap < 1)
      return (int)cudaErrorInvalidValue;
    if (match_mode) return launch_e<kGather, true, 1, true>(a, stream);
    if (with_d) return launch_e<kGather, true, 0, true>(a, stream);
    return launch_e<kGather, false, 0, true>(a, stream);
  }
  if (mode == kGather) {
    if (match_mode == 2) return launch_e<kGather, true, 2>(a, stream);
    if (match_mode) return launch_e<kGather, true, 1>(a, stream);
    if (with_d) return launch_e<kGather, true, 0>(a, stream);
    return launch_e<kGather, false, 0>(a, stream);
  } else {
    if (match_mode == 2) return launch_e<kSorted, true, 2>(a, stream);
    if (match_mode) return launch_e<kSorted, true, 1>(a, stream);
    if (with_d) return launch_e<kSorted, true, 0>(a, stream);
    return launch_e<kSorted, false, 0>(a, stream);
  }
}

}  // namespace pbwtgpu
