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
//   * shared-memory accesses to a[]/d[] are conflict-free for every E, addressed base + immediate.
// Positions M..1024E-1 are pad elements: their a[] points at a column bit that is always 1, so they
// sort to themselves forever and no per-element validity predicate is needed.
// Divergence update, single segmented max-scan: d_{k+1}[new(i)] = d_k[i] unless
// i is the head of a run of equal alleles, in which case it is max(d_k[i], max d_k over the previous
// run); the first two run heads see the k+1 sentinel in d_k[0].
// Column bits are staged by 1-D TMA bulk copies (cp.async.bulk + mbarrier), kStages sites ahead.
// Two __syncthreads per site.  matchMaximalWithin's block scans are warp-cooperative (32 positions
// per step) and its records are written 32 at a time.
#include "sweep.cuh"

namespace pbwtgpu {
namespace {

constexpr int kThreads = 1024;
constexpr int kStages = 4;

template <int E, bool WITH_D>
struct SmemLayout {  // offsets in 32-bit words; every array start is 16-byte aligned
  static constexpr int MP = kThreads * E;
  static constexpr int WORDS = 32 * E;
  static constexpr int STAGE = WORDS + 4;  // + one always-ones word the pad elements point at
  static constexpr int a0 = 0;
  static constexpr int a1 = a0 + MP;
  static constexpr int d0 = a1 + MP;
  static constexpr int d1 = d0 + (WITH_D ? MP + 4 : 0);
  static constexpr int ycol = d1 + (WITH_D ? MP + 4 : 0);
  static constexpr int xcol = ycol + 2 * WORDS;
  static constexpr int wsum = xcol + kStages * STAGE;
  static constexpr int mbar = wsum + 3 * 32;
  static constexpr int totalWords = mbar + 2 * kStages;
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

// number of positions in [lo, hi) whose allele equals yi, counted from the sorted bit column
__device__ __forceinline__ bool any_same_allele(const uint32_t *yc, int lo, int hi, uint32_t yi, int lane) {
  if (lo >= hi) return false;
  const int w0 = lo >> 5, w1 = (hi - 1) >> 5;
  bool hit = false;
  for (int w = w0 + lane; w <= w1; w += 32) {
    uint32_t bits = yi ? yc[w] : ~yc[w];
    if (w == w0) bits &= kFull << (lo & 31);
    if (w == w1 && (hi & 31)) bits &= (1u << (hi & 31)) - 1u;
    hit |= bits != 0u;
  }
  return __any_sync(kFull, hit);
}

// One pass of matchMaximalWithin's per-site scan (pbwtMatch.c:122-137) over this warp's words.
//   EMIT = false: count; returns the warp's record total and a bitmask of words that report.
//   EMIT = true : redo the reporting words and write their records from offset off, in callback
//                 order: i ascending, upward block (start d[i]) then downward block (start d[i+1]).
// For a position i the upward block is [q, i) with q the nearest index <= i whose d exceeds d[i], the
// downward block (i, r) with r the nearest index > i whose d exceeds d[i+1]; i reports nothing when a
// member of either block carries the same allele (the match extends).  Only heads/tails of allele runs
// can fail that test on their first neighbour, so interior positions are never scanned (live sites).
template <int E, bool EMIT>
__device__ __forceinline__ int match_pass(const int32_t *sA, const int32_t *sD, const uint32_t *yc, int M, int k,
                                          bool live, int warp, int lane, unsigned &wordMask, pbwtgpu_match *out,
                                          long long cap, long long off) {
  int warpTotal = 0;
  unsigned found = 0;
#pragma unroll 1
  for (int e = 0; e < E; ++e) {
    if (EMIT && !((wordMask >> e) & 1u)) continue;
    const int wd = warp * E + e;
    const int base = wd * 32;
    if (base >= M) break;
    const uint32_t cw = yc[wd];
    const uint32_t vm = (M - base >= 32) ? kFull : ((1u << (M - base)) - 1u);
    uint32_t cand = vm;
    if (live) {
      const uint32_t pb = wd > 0 ? (yc[wd - 1] >> 31) : (cw & 1u);
      const uint32_t nw = (base + 32 < M) ? yc[wd + 1] : 0u;
      const uint32_t H = cw ^ ((cw << 1) | pb);
      const uint32_t Hn = (H >> 1) | (((nw & 1u) ^ (cw >> 31)) << 31);
      const uint32_t bm = (H | Hn) & vm;
      if (!bm) continue;
      bool mine = false;
      if ((bm >> lane) & 1u) {
        const int di = sD[base + lane], dn1 = sD[base + lane + 1];
        const bool samePrev = !((H >> lane) & 1u), sameNext = !((Hn >> lane) & 1u);
        mine = !((di <= dn1 && samePrev) || (di >= dn1 && sameNext));
      }
      cand = __ballot_sync(kFull, mine);
      if (!cand) continue;
    }
    int wordTotal = 0;
    while (cand) {
      const int j = __ffs(cand) - 1;
      cand &= cand - 1;
      const int i = base + j;
      const uint32_t yi = (cw >> j) & 1u;
      const int di = sD[i], dn1 = sD[i + 1];
      int lo = i, hi = i + 1;
      bool extends = false;
      if (di <= dn1) {  // nearest q <= i with d[q] > d[i]  (d[0] = k+1 always stops the search)
        int q = i;
        for (;;) {
          const int idx = q - lane;
          const bool stop = idx <= 0 || sD[idx] > di;
          const unsigned sm = __ballot_sync(kFull, stop);
          if (sm) { q -= __ffs(sm) - 1; break; }
          q -= 32;
        }
        lo = max(q, 0);
        extends = live && any_same_allele(yc, lo, i, yi, lane);
      }
      if (!extends && di >= dn1) {  // nearest r > i with d[r] > d[i+1]  (d[M] = k+1 always stops)
        int r = i + 1;
        for (;;) {
          const int idx = r + lane;
          const bool stop = idx >= M || sD[idx] > dn1;
          const unsigned sm = __ballot_sync(kFull, stop);
          if (sm) { r += __ffs(sm) - 1; break; }
          r += 32;
        }
        hi = min(r, M);
        extends = live && any_same_allele(yc, i + 1, hi, yi, lane);
      }
      if (extends) continue;
      const int cnt = (i - lo) + (hi - i - 1);
      if (EMIT && cnt) {
        const long long o = off + wordTotal;
        const int ai = sA[i];
        for (int t = lane; t < cnt; t += 32) {
          const int jj = t < i - lo ? lo + t : t + lo + 1;  // skip i itself
          if (o + t < cap)
            reinterpret_cast<int4 *>(out)[o + t] = make_int4(ai, sA[jj], t < i - lo ? di : dn1, k);
        }
      }
      wordTotal += cnt;
    }
    if (wordTotal) {
      found |= 1u << e;
      off += wordTotal;
      warpTotal += wordTotal;
    }
  }
  if (!EMIT) wordMask = found;
  return warpTotal;
}

template <int E, int MODE, bool WITH_D, bool MATCH>
__global__ void __launch_bounds__(kThreads, 1) sweep_smem_kernel(const SweepArgs p) {
  using L = SmemLayout<E, WITH_D>;
  extern __shared__ __align__(16) uint32_t smem[];
  int32_t *const sI = reinterpret_cast<int32_t *>(smem);
  int32_t *const wsumZ = sI + L::wsum;
  uint32_t *const wsumR = smem + L::wsum + 32;
  int32_t *const wsumC = sI + L::wsum + 64;
  uint64_t *const mbar = reinterpret_cast<uint64_t *>(smem + L::mbar);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = p.M, wpc = p.wpc;
  const int win = blockIdx.x;
  const int k0 = win * p.sitesPerWindow;
  const int k1 = min(p.N, k0 + p.sitesPerWindow);
  const uint32_t colBytes = (uint32_t)wpc * 4u;
  const int wordBase = warp * E;
  const int laneOff = wordBase * 32 + lane;  // position of (word 0 of this warp, this lane)
  // the thread owning position M (a pad element) also maintains the d[M] sentinel
  const int ownerM = M < L::MP ? ((M >> 5) / E) * 32 + (M & 31) : 0;

  // ---- window start state (pbwtNakedCursorCreate, pbwtCore.c:402-418) ----
  for (int i = tid; i < L::MP; i += kThreads)
    sI[L::a0 + i] = i < M ? (p.aStart ? p.aStart[(size_t)win * M + i] : i) : L::padIndex;
  if (WITH_D)
    for (int i = tid; i < L::MP + 4; i += kThreads) {
      int v = 0;
      if (i == 0 || i == M) v = k0 + 1;
      else if (i < M && p.dStart) v = p.dStart[(size_t)win * (M + 1) + i];
      sI[L::d0 + i] = v;
      sI[L::d1 + i] = 0;
    }
  for (int i = tid; i < kStages * L::STAGE; i += kThreads) smem[L::xcol + i] = kFull;  // pad bits are ones
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
        tma_load_1d(smem + L::xcol + s * L::STAGE, p.src + (size_t)(k0 + s) * wpc, colBytes, &mbar[s]);
      }
  }

  pbwtgpu_match *recOut = MATCH && p.rec ? p.rec + (size_t)win * p.recCap : nullptr;
  long long recBase = 0;  // records produced so far by this window (uniform)
  int cur = 0;

  for (int k = k0; k < k1; ++k) {
    const int it = k - k0;
    const int stage = it % kStages;
    mbar_wait(&mbar[stage], (uint32_t)(it / kStages) & 1u);
    const uint32_t *col = smem + L::xcol + stage * L::STAGE;
    const int aC = cur ? L::a1 : L::a0, aN = cur ? L::a0 : L::a1;
    const int dC = cur ? L::d1 : L::d0, dN = cur ? L::d0 : L::d1;
    uint32_t *yc = smem + L::ycol + (it & 1) * L::WORDS;

    // ---------------- phase A: gather, ballots, warp-local scans ----------------
    uint32_t B[E];
    int32_t hv[E];
    int32_t dn[E];
    uint32_t pb = 0;  // allele just before this warp's first position
    if (warp > 0) {
      if (MODE == kGather) {
        const int hp = sI[aC + wordBase * 32 - 1];
        pb = (col[hp >> 5] >> (hp & 31)) & 1u;
      } else pb = col[wordBase - 1] >> 31;
    }
    int zw = 0;            // zeros in this warp so far
    int carry = 0;         // max d since the last run head seen by this warp (or over all of it)
    bool anyHead = false;  // has this warp seen a run head
    // the warp's first run head additionally needs the carry entering the warp (patched after S1)
    int patchLane = 0, pDn = 0, pIdx = 0, pZ = 0;
    uint32_t pY = 0;
    uint32_t myY = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int idx = laneOff + e * 32;
      hv[e] = sI[aC + idx];
      if (MODE == kGather) B[e] = __ballot_sync(kFull, (col[hv[e] >> 5] >> (hv[e] & 31)) & 1u);
      else B[e] = col[wordBase + e];
      if (e == 0 && warp == 0) pb = (~B[0]) & 1u;  // sorted position 0 is always a run head
      const uint32_t H = B[e] ^ ((B[e] << 1) | pb);
      pb = B[e] >> 31;
      if (lane == e) myY = B[e];
      if (WITH_D) {
        const int dv = sI[dC + idx];
        if (H == 0u) {
          dn[e] = dv;
          carry = max(carry, __reduce_max_sync(kFull, dv));
        } else {
          int v = dv;  // inclusive max since the nearest head at or below this lane
#pragma unroll
          for (int dl = 1; dl < 32; dl <<= 1) {
            const int t = __shfl_up_sync(kFull, v, dl);
            if (lane >= dl && ((H >> (lane - dl + 1)) & ((1u << dl) - 1u)) == 0u) v = max(v, t);
          }
          const int prev = __shfl_up_sync(kFull, v, 1);
          const bool isHead = (H >> lane) & 1u;
          const bool headBelow = (H & lanemask_lt(lane)) != 0u;
          const int prevMax = lane == 0 ? carry : (headBelow ? prev : max(carry, prev));
          dn[e] = isHead ? max(dv, prevMax) : dv;
          if (!anyHead) {
            anyHead = true;
            patchLane = __ffs(H) - 1;
            pDn = dn[e];
            pIdx = idx;
            pZ = zw + __popc(~B[e] & lanemask_lt(lane));
            pY = (B[e] >> lane) & 1u;
          }
          carry = __shfl_sync(kFull, v, 31);
        }
      }
      zw += __popc(~B[e]);
    }
    if (lane < E) {
      const int base = (wordBase + lane) * 32;
      const uint32_t vm = (M - base >= 32) ? kFull : (M > base ? ((1u << (M - base)) - 1u) : 0u);
      yc[wordBase + lane] = myY & vm;
    }
    if (lane == 0) {
      wsumZ[warp] = zw;
      if (WITH_D) wsumR[warp] = (anyHead ? 0x80000000u : 0u) | (uint32_t)carry;
    }
    __syncthreads();  // S1: ballots, warp sums published; everyone is done with col[]

    if (tid == 0 && k + kStages < k1) {  // refill the column slot just consumed
      fence_proxy_async();
      mbar_expect_tx(&mbar[stage], colBytes);
      tma_load_1d(smem + L::xcol + stage * L::STAGE, p.src + (size_t)(k + kStages) * wpc, colBytes, &mbar[stage]);
    }

    // ---------------- cross-warp prefixes (each warp redundantly) ----------------
    int zIn, c;
    {
      const int z = wsumZ[lane];
      int tot;
      const int ex = warp_excl_scan(z, lane, tot);
      zIn = __shfl_sync(kFull, ex, warp);
      c = tot;
    }
    int cIn = 0;
    if (WITH_D) {
      uint32_t x = wsumR[lane];
#pragma unroll
      for (int dl = 1; dl < 32; dl <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, x, dl);
        if (lane >= dl && !(x >> 31)) x = (t & 0x80000000u) | max(t & 0x7fffffffu, x);
      }
      const uint32_t inc = __shfl_sync(kFull, x, warp > 0 ? warp - 1 : 0);
      cIn = warp > 0 ? (int)(inc & 0x7fffffffu) : 0;
    }

    // ---------------- phase C: stable partition scatter ----------------
    {
      int zacc = zIn;
      const int onesBase = c + laneOff;  // position of a one = c + i - zeros before i
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int z = zacc + __popc(~B[e] & lanemask_lt(lane));
        const int pos = ((B[e] >> lane) & 1u) ? onesBase + e * 32 - z : z;
        sI[aN + pos] = hv[e];
        if (WITH_D) sI[dN + pos] = pos == 0 ? k + 2 : dn[e];
        zacc += __popc(~B[e]);
      }
      if (WITH_D) {
        if (anyHead && lane == patchLane && cIn > pDn) {  // same thread wrote this slot above
          const int z = zIn + pZ;
          const int pos = pY ? c + pIdx - z : z;
          if (pos != 0) sI[dN + pos] = cIn;
        }
        if (tid == ownerM) sI[dN + M] = k + 2;  // after this thread's own pad write to slot M
      }
    }
    if (p.yOut && tid < wpc) p.yOut[(size_t)k * wpc + tid] = yc[tid];
    if (p.cOut && tid == 0) p.cOut[k] = c;

    // ---------------- matchMaximalWithin on state k ----------------
    unsigned wordMask = 0;
    if (MATCH) {
      const int wt = match_pass<E, false>(sI + aC, sI + dC, yc, M, k, true, warp, lane, wordMask, nullptr, 0, 0);
      if (lane == 0) wsumC[warp] = wt;
    }
    __syncthreads();  // S2: scatter complete; match counts published
    if (MATCH) {
      const int cw = wsumC[lane];
      int tot;
      const int ex = warp_excl_scan(cw, lane, tot);
      const int mine = __shfl_sync(kFull, ex, warp);
      if (wordMask && recOut)
        match_pass<E, true>(sI + aC, sI + dC, yc, M, k, true, warp, lane, wordMask, recOut, p.recCap, recBase + mine);
      recBase += tot;
    }
    cur ^= 1;
  }

  // ---------------- window end ----------------
  const int aC = cur ? L::a1 : L::a0;
  const int dC = cur ? L::d1 : L::d0;
  if (p.aEnd)
    for (int i = tid; i < M; i += kThreads) p.aEnd[(size_t)win * M + i] = sI[aC + i];
  if (WITH_D && p.dEnd)
    for (int i = tid; i <= M; i += kThreads) p.dEnd[(size_t)win * (M + 1) + i] = sI[dC + i];

  if (MATCH && p.reportAtN && k1 == p.N) {  // k == N: nothing extends, every block is reported
    const uint32_t *yc = smem + L::ycol;    // alleles are not consulted when live == false
    unsigned wordMask = 0;
    const int wt = match_pass<E, false>(sI + aC, sI + dC, yc, M, p.N, false, warp, lane, wordMask, nullptr, 0, 0);
    __syncthreads();  // previous readers of wsumC are done
    if (lane == 0) wsumC[warp] = wt;
    __syncthreads();
    const int cw = wsumC[lane];
    int tot;
    const int ex = warp_excl_scan(cw, lane, tot);
    const int mine = __shfl_sync(kFull, ex, warp);
    if (wordMask && recOut)
      match_pass<E, true>(sI + aC, sI + dC, yc, M, p.N, false, warp, lane, wordMask, recOut, p.recCap, recBase + mine);
    recBase += tot;
  }
  if (MATCH && p.recCount && tid == 0) p.recCount[win] = (unsigned long long)recBase;
}

template <int E, int MODE, bool WITH_D, bool MATCH>
int launch_one(const SweepArgs &a, cudaStream_t stream) {
  using L = SmemLayout<E, WITH_D>;
  auto kern = sweep_smem_kernel<E, MODE, WITH_D, MATCH>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes);
  if (err != cudaSuccess) return (int)err;
  kern<<<a.nWindows, kThreads, L::bytes, stream>>>(a);
  return (int)cudaGetLastError();
}

template <int MODE, bool WITH_D, bool MATCH>
int launch_e(const SweepArgs &a, cudaStream_t stream) {
  const int words = (a.M + 31) / 32;
  const int e = (words + 31) / 32;
  if (e <= 1) return launch_one<1, MODE, WITH_D, MATCH>(a, stream);
  if (e <= 2) return launch_one<2, MODE, WITH_D, MATCH>(a, stream);
  if (e <= 4) return launch_one<4, MODE, WITH_D, MATCH>(a, stream);
  if (e <= 7) return launch_one<7, MODE, WITH_D, MATCH>(a, stream);
  if (e <= 10) return launch_one<10, MODE, WITH_D, MATCH>(a, stream);
  if (e <= 13) return launch_one<13, MODE, WITH_D, MATCH>(a, stream);
  if constexpr (!WITH_D) {
    if (e <= 19) return launch_one<19, MODE, WITH_D, MATCH>(a, stream);
    if (e <= 25) return launch_one<25, MODE, WITH_D, MATCH>(a, stream);
  }
  return (int)cudaErrorInvalidValue;
}

}  // namespace

int sweep_smem_max_m(bool with_d) { return with_d ? 13 * 1024 : 25 * 1024; }

int launch_sweep_smem(const SweepArgs &a, SweepMode mode, bool with_d, bool with_match, cudaStream_t stream) {
  if (a.M < 1 || a.M > sweep_smem_max_m(with_d || with_match)) return (int)cudaErrorInvalidValue;
  if (mode == kGather) {
    if (with_match) return launch_e<kGather, true, true>(a, stream);
    if (with_d) return launch_e<kGather, true, false>(a, stream);
    return launch_e<kGather, false, false>(a, stream);
  } else {
    if (with_match) return launch_e<kSorted, true, true>(a, stream);
    if (with_d) return launch_e<kSorted, true, false>(a, stream);
    return launch_e<kSorted, false, false>(a, stream);
  }
}

}  // namespace pbwtgpu
