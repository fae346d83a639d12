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
//   * shared-memory accesses to a[]/d[] are conflict-free for every E.
// Divergence update, single segmented max-scan: d_{k+1}[new(i)] = d_k[i] unless
// i is the head of a run of equal alleles, in which case it is max(d_k[i], max d_k over the previous
// run); the first two run heads see the k+1 sentinel in d_k[0].
// Column bits are staged by 1-D TMA bulk copies (cp.async.bulk + mbarrier), kStages sites ahead.
// Two __syncthreads per site.
#include "sweep.cuh"

namespace pbwtgpu {
namespace {

constexpr int kThreads = 1024;
constexpr int kStages = 4;

template <int E, bool WITH_D>
struct SmemLayout {  // offsets in 32-bit words; every array start is 16-byte aligned
  static constexpr int MP = kThreads * E;
  static constexpr int WORDS = 32 * E;
  static constexpr int a0 = 0;
  static constexpr int a1 = a0 + MP;
  static constexpr int d0 = a1 + MP;
  static constexpr int d1 = d0 + (WITH_D ? MP + 4 : 0);
  static constexpr int ycol = d1 + (WITH_D ? MP + 4 : 0);
  static constexpr int xcol = ycol + 2 * WORDS;
  static constexpr int wsum = xcol + kStages * WORDS;
  static constexpr int mbar = wsum + 3 * 32;
  static constexpr int totalWords = mbar + 2 * kStages;
  static constexpr size_t bytes = (size_t)totalWords * 4;
};

__device__ __forceinline__ int ybit(const uint32_t *yc, int j) { return (yc[j >> 5] >> (j & 31)) & 1; }

// The d-bounded neighbour scan of matchMaximalWithin for sorted position i (pbwtMatch.c:123-129).
// Returns the number of records i reports (0 when some visited neighbour carries the same allele,
// i.e. the match extends); (lo..i) and (i..hi) are the reported blocks.
__device__ __forceinline__ int match_block(const int32_t *dC, const uint32_t *yc, int M, int i, int yi, int di,
                                           int dn1, bool live, int &lo, int &hi) {
  lo = i;
  hi = i + 1;
  if (di <= dn1) {
    while (lo > 0 && dC[lo] <= di) {
      --lo;
      if (live && ybit(yc, lo) == yi) return 0;
    }
  }
  if (di >= dn1) {
    while (hi < M && dC[hi] <= dn1) {
      if (live && ybit(yc, hi) == yi) return 0;
      ++hi;
    }
  }
  return (i - lo) + (hi - i - 1);
}

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

// One pass of the per-site match scan over this warp's words.
//   EMIT = false: count; returns the warp's record total and a bitmask of words that report.
//   EMIT = true : recompute the reporting words and write records from offset off (callback order:
//                 i ascending, upward block then downward block, pbwtMatch.c:133-134).
template <int E, bool EMIT>
__device__ __forceinline__ int match_pass(const int32_t *aC, const int32_t *dC, const uint32_t *yc, int M, int k,
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
    uint32_t H = 0, Hn = 0, bm = vm;
    if (live) {  // only heads/tails of allele runs can fail to extend on their first neighbour
      const uint32_t pb = wd > 0 ? (yc[wd - 1] >> 31) : (cw & 1u);
      const uint32_t nw = (base + 32 < M) ? yc[wd + 1] : 0u;
      H = cw ^ ((cw << 1) | pb);
      Hn = (H >> 1) | (((nw & 1u) ^ (cw >> 31)) << 31);
      bm = (H | Hn) & vm;
      if (!bm) continue;
    }
    int cnt = 0, lo = 0, hi = 0, di = 0, dn1 = 0;
    const int i = base + lane;
    if ((bm >> lane) & 1u) {
      di = dC[i];
      dn1 = dC[i + 1];
      bool skip = false;
      if (live) {
        const bool samePrev = !((H >> lane) & 1u), sameNext = !((Hn >> lane) & 1u);
        skip = (di <= dn1 && samePrev) || (di >= dn1 && sameNext);
      }
      if (!skip) cnt = match_block(dC, yc, M, i, (cw >> lane) & 1u, di, dn1, live, lo, hi);
    }
    const unsigned cm = __ballot_sync(kFull, cnt > 0);
    if (!cm) continue;
    found |= 1u << e;
    if (!EMIT) {
      warpTotal += __reduce_add_sync(kFull, cnt);
    } else {
      int total;
      const int pre = warp_excl_scan(cnt, lane, total);
      if (cnt > 0) {
        long long o = off + pre;
        const int ai = aC[i];
        for (int j = lo; j < i; ++j, ++o)
          if (o < cap) reinterpret_cast<int4 *>(out)[o] = make_int4(ai, aC[j], di, k);
        for (int j = i + 1; j < hi; ++j, ++o)
          if (o < cap) reinterpret_cast<int4 *>(out)[o] = make_int4(ai, aC[j], dn1, k);
      }
      off += total;
      warpTotal += total;
    }
  }
  if (!EMIT) wordMask = found;
  return warpTotal;
}

template <int E, int MODE, bool WITH_D, bool MATCH>
__global__ void __launch_bounds__(kThreads, 1) sweep_smem_kernel(const SweepArgs p) {
  using L = SmemLayout<E, WITH_D>;
  extern __shared__ __align__(16) uint32_t smem[];
  int32_t *aBuf[2] = {reinterpret_cast<int32_t *>(smem + L::a0), reinterpret_cast<int32_t *>(smem + L::a1)};
  int32_t *dBuf[2] = {reinterpret_cast<int32_t *>(smem + L::d0), reinterpret_cast<int32_t *>(smem + L::d1)};
  uint32_t *ycolBase = smem + L::ycol;
  uint32_t *xcolBase = smem + L::xcol;
  int32_t *wsumZ = reinterpret_cast<int32_t *>(smem + L::wsum);
  uint32_t *wsumR = smem + L::wsum + 32;
  int32_t *wsumC = reinterpret_cast<int32_t *>(smem + L::wsum + 64);
  uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + L::mbar);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = p.M, wpc = p.wpc;
  const int win = blockIdx.x;
  const int k0 = win * p.sitesPerWindow;
  const int k1 = min(p.N, k0 + p.sitesPerWindow);
  const uint32_t colBytes = (uint32_t)wpc * 4u;

  // ---- window start state (pbwtNakedCursorCreate, pbwtCore.c:402-418) ----
  for (int i = tid; i < M; i += kThreads) aBuf[0][i] = p.aStart ? p.aStart[(size_t)win * M + i] : i;
  if (WITH_D)
    for (int i = tid; i <= M; i += kThreads)
      dBuf[0][i] = p.dStart ? p.dStart[(size_t)win * (M + 1) + i] : ((i == 0 || i == M) ? k0 + 1 : 0);
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&mbar[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0)
    for (int s = 0; s < kStages; ++s)
      if (k0 + s < k1) {
        mbar_expect_tx(&mbar[s], colBytes);
        tma_load_1d(xcolBase + s * L::WORDS, p.src + (size_t)(k0 + s) * wpc, colBytes, &mbar[s]);
      }

  pbwtgpu_match *recOut = MATCH && p.rec ? p.rec + (size_t)win * p.recCap : nullptr;
  long long recBase = 0;  // records produced so far by this window (uniform)
  int cur = 0;

  for (int k = k0; k < k1; ++k) {
    const int it = k - k0;
    const int stage = it % kStages;
    mbar_wait(&mbar[stage], (uint32_t)(it / kStages) & 1u);
    const uint32_t *col = xcolBase + stage * L::WORDS;
    const int32_t *aC = aBuf[cur];
    int32_t *aN = aBuf[cur ^ 1];
    const int32_t *dC = dBuf[cur];
    int32_t *dN = dBuf[cur ^ 1];
    uint32_t *yc = ycolBase + (it & 1) * L::WORDS;

    // ---------------- phase A: gather, ballots, warp-local scans ----------------
    uint32_t B[E];
    int32_t hv[E];
    int32_t dn[E];
    uint32_t prevBit;  // allele at the position just before this warp's first
    {
      const int firstPos = warp * E * 32;
      if (warp == 0 || firstPos > M) prevBit = 2u;  // resolved below: position 0 is always a run head
      else if (MODE == kGather) {
        const int hp = aC[firstPos - 1];
        prevBit = (col[hp >> 5] >> (hp & 31)) & 1u;
      } else prevBit = col[(firstPos - 1) >> 5] >> 31;
    }
    int zw = 0;            // zeros in this warp so far
    int carry = 0;         // max d since the last run head seen by this warp (or over all of it)
    bool anyHead = false;  // has this warp seen a run head
    int patchE = -1, patchLane = 0;  // the first run head of the warp also needs the incoming carry
    uint32_t myY = 0;
    uint32_t pb = prevBit;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int wd = warp * E + e;
      const int i = wd * 32 + lane;
      const bool valid = i < M;
      hv[e] = valid ? aC[i] : 0;
      if (MODE == kGather) {
        const uint32_t bit = valid ? ((col[hv[e] >> 5] >> (hv[e] & 31)) & 1u) : 1u;
        B[e] = __ballot_sync(kFull, bit);
      } else {
        const int base = wd * 32;
        const uint32_t vm = (M - base >= 32) ? kFull : (M > base ? ((1u << (M - base)) - 1u) : 0u);
        B[e] = (base < M ? col[wd] : 0u) | ~vm;
      }
      if (pb == 2u) pb = (~B[e]) & 1u;  // forces a head at sorted position 0
      const uint32_t H = B[e] ^ ((B[e] << 1) | pb);
      pb = B[e] >> 31;
      zw += __popc(~B[e]);
      if (lane == e) myY = B[e];
      if (WITH_D) {
        const int dv = valid ? dC[i] : 0;
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
            patchE = e;
            patchLane = __ffs(H) - 1;
          }
          carry = __shfl_sync(kFull, v, 31);
        }
      }
    }
    if (lane < E) {
      const int wd = warp * E + lane;
      const int base = wd * 32;
      const uint32_t vm = (M - base >= 32) ? kFull : (M > base ? ((1u << (M - base)) - 1u) : 0u);
      yc[wd] = myY & vm;
    }
    if (lane == 0) {
      wsumZ[warp] = zw;
      if (WITH_D) wsumR[warp] = (anyHead ? 0x80000000u : 0u) | (uint32_t)carry;
    }
    __syncthreads();  // S1: ballots, warp sums published; everyone is done with col[]

    if (tid == 0 && k + kStages < k1) {  // refill the column slot just consumed
      fence_proxy_async();
      mbar_expect_tx(&mbar[stage], colBytes);
      tma_load_1d(xcolBase + stage * L::WORDS, p.src + (size_t)(k + kStages) * wpc, colBytes, &mbar[stage]);
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
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int i = (warp * E + e) * 32 + lane;
        if (i < M) {
          const uint32_t y = (B[e] >> lane) & 1u;
          const int z = zacc + __popc(~B[e] & lanemask_lt(lane));
          const int pos = y ? c + i - z : z;
          aN[pos] = hv[e];
          if (WITH_D) {
            int dv = dn[e];
            if (e == patchE && lane == patchLane) dv = max(dv, cIn);
            dN[pos] = pos == 0 ? k + 2 : dv;
          }
        }
        zacc += __popc(~B[e]);
      }
      if (WITH_D && tid == 0) dN[M] = k + 2;
    }
    if (p.yOut && tid < wpc) p.yOut[(size_t)k * wpc + tid] = tid < L::WORDS ? yc[tid] : 0u;
    if (p.cOut && tid == 0) p.cOut[k] = c;

    // ---------------- matchMaximalWithin on state k ----------------
    unsigned wordMask = 0;
    if (MATCH) {
      const int wt = match_pass<E, false>(aC, dC, yc, M, k, true, warp, lane, wordMask, nullptr, 0, 0);
      if (lane == 0) wsumC[warp] = wt;
    }
    __syncthreads();  // S2: scatter complete; match counts published
    if (MATCH) {
      const int cw = wsumC[lane];
      int tot;
      const int ex = warp_excl_scan(cw, lane, tot);
      const int mine = __shfl_sync(kFull, ex, warp);
      if (wordMask && recOut)
        match_pass<E, true>(aC, dC, yc, M, k, true, warp, lane, wordMask, recOut, p.recCap, recBase + mine);
      recBase += tot;
    }
    cur ^= 1;
  }

  // ---------------- window end ----------------
  const int32_t *aC = aBuf[cur];
  const int32_t *dC = dBuf[cur];
  if (p.aEnd)
    for (int i = tid; i < M; i += kThreads) p.aEnd[(size_t)win * M + i] = aC[i];
  if (WITH_D && p.dEnd)
    for (int i = tid; i <= M; i += kThreads) p.dEnd[(size_t)win * (M + 1) + i] = dC[i];

  if (MATCH && p.reportAtN && k1 == p.N) {  // k == N: nothing extends, every block is reported
    const uint32_t *yc = ycolBase;          // alleles are not consulted when live == false
    unsigned wordMask = 0;
    const int wt = match_pass<E, false>(aC, dC, yc, M, p.N, false, warp, lane, wordMask, nullptr, 0, 0);
    __syncthreads();  // previous readers of wsumC are done
    if (lane == 0) wsumC[warp] = wt;
    __syncthreads();
    const int cw = wsumC[lane];
    int tot;
    const int ex = warp_excl_scan(cw, lane, tot);
    const int mine = __shfl_sync(kFull, ex, warp);
    if (wordMask && recOut)
      match_pass<E, true>(aC, dC, yc, M, p.N, false, warp, lane, wordMask, recOut, p.recCap, recBase + mine);
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
