// combine_sort.cu -- the window-prefix combine (windows.cu) as a SORT instead of a sweep over pseudo-sites.
//
// For adjacent nodes I (earlier) and J (later), windows.cu defines
//     order(I+J)      = order(I) stably re-sorted by J's tie-group index,
//     divergence(I+J) = between different J-groups: J's divergence at the head of the later group,
//                       inside one J-group: the maximum of I's divergence over the I-positions in between,
// and computes it by sweeping (order(I), divergence(I)) over ceil(log2 M) pseudo-sites made of the bits of J's group
// index -- 14 dependent partitions per combine for 10,000 haplotypes.  Here the same node comes from
//   * an LSD radix sort (two 8-bit digits) of I's positions by the group index of the haplotype they hold, the elements
//     (group << 16 | I-position) resident in shared memory (stable, so members of a group keep their I-order), and
//   * one range-maximum query over I's divergence per neighbouring pair of the same group: blocks of 32 positions, a
//     sparse table over the block maxima, and for the two partial blocks the block's prefix / suffix maxima (two loads
//     instead of a walk over up to 31 positions each; up to 10,240 haplotypes, where the two extra arrays fit),
// followed in the same launch by what make_groups_kernel does for the new node (tie-group index of every haplotype,
// divergence at every group head).  One CTA per node; bit-identical to the sweep formulation
// (tests force both paths).  Panels up to 13,312 haplotypes in haplotype space; stored streams (position space) and
// wide panels keep the sweep.
#include "sweep.cuh"
#include "windows.cuh"

namespace pbwtgpu {
namespace {

constexpr int kThreads = 1024;
constexpr uint32_t kPadPos = 0xffffu;
constexpr int kLevels = 9;  // sparse table over up to 512 blocks of 32 positions

template <int E>
__global__ void __launch_bounds__(kThreads, 1) combine_sort_kernel(const CombineSortArgs c) {
  constexpr int MP = kThreads * E;
  constexpr int NB = 32 * E;  // blocks of 32 positions
  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t *buf0 = smem;  // (group << 16 | I-position), double buffered
  uint32_t *buf1 = buf0 + MP;
  int32_t *dS = reinterpret_cast<int32_t *>(buf1 + MP);  // I's divergence by I-position (0 at position 0 and beyond M)
  int32_t *st = dS + MP;                                 // [kLevels][NB] sparse table of block maxima
  uint16_t *hist = reinterpret_cast<uint16_t *>(st + kLevels * NB);  // [32 warps][256 digits]
  constexpr bool PS = E <= 10;  // prefix / suffix maxima inside every block of 32 positions
  int32_t *sufS = reinterpret_cast<int32_t *>(hist + 32 * 256);  // [MP] max of dS from the position to its block's end (PS)
  int32_t *preS = reinterpret_cast<int32_t *>(buf1);             // [MP] max of dS from its block's start to the position (PS;
                                                                 // the sort's second buffer, free once the sort is done)
  __shared__ int warpTot[32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = c.M;
  const int n = blockIdx.x;
  if (n >= c.cnt) {  // nodes that are already final: copied through to the other buffer set
    const int s = n - c.cnt;
    for (int i = tid; i < M; i += kThreads) {
      c.ordCopyDst[(size_t)s * M + i] = c.ordCopySrc[(size_t)s * M + i];
      c.gCopyDst[(size_t)s * M + i] = c.gCopySrc[(size_t)s * M + i];
      c.gfdCopyDst[(size_t)s * M + i] = c.gfdCopySrc[(size_t)s * M + i];
    }
    for (int i = tid; i <= M; i += kThreads) c.dvCopyDst[(size_t)s * (M + 1) + i] = c.dvCopySrc[(size_t)s * (M + 1) + i];
    return;
  }
  const int32_t *ordI = c.ordI + (c.sharedI ? 0 : (size_t)n * M);
  const int32_t *dvI = c.dvI + (c.sharedI ? 0 : (size_t)n * (M + 1));
  const int32_t *gJ = c.gJ + (size_t)n * M;
  const int32_t *gfdJ = c.gfdJ + (size_t)n * M;
  int32_t *ordOut = c.ordOut + (size_t)n * M;
  int32_t *dvOut = c.dvOut + (size_t)n * (M + 1);
  int32_t *gOut = c.gOut + (size_t)n * M;
  int32_t *gfdOut = c.gfdOut + (size_t)n * M;
  const uint32_t ltMask = lanemask_lt(lane);
  const int posBase = warp * E * 32 + lane;  // position of (word 0 of this warp, this lane)

  // ---- elements and I's divergence; block maxima (one warp load = one block of 32 positions) ----
  for (int i = tid; i < 32 * 256 / 2; i += kThreads) reinterpret_cast<uint32_t *>(hist)[i] = 0u;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = posBase + e * 32;
    uint32_t v = 0xffff0000u | kPadPos;
    int d = 0;
    if (i < M) {
      v = ((uint32_t)__ldg(gJ + __ldg(ordI + i)) << 16) | (uint32_t)i;
      if (i > 0) d = __ldg(dvI + i);
    }
    buf0[i] = v;
    dS[i] = d;
    const int bm = __reduce_max_sync(kFull, d);
    if (lane == 0) st[warp * E + e] = bm;
  }
  __syncthreads();

  // ---- LSD radix sort by the group index, two 8-bit digits (as window_sort.cu) ----
  uint32_t *cur = buf0, *nxt = buf1;
  for (int pass = 0; pass < 2; ++pass) {
    const int shift = 16 + 8 * pass;
    int off[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const uint32_t v = cur[posBase + e * 32];
      const uint32_t digit = (v >> shift) & 255u;
      // lanes holding the same digit, by eight independent votes: a warp's 32 positions hold haplotypes of many different
      // groups, and match.any takes time in proportion to the number of distinct values (combine 0.42 -> 0.36 ms per step;
      // the window sorts' digits are allele bits with few distinct values, there match.any is the faster form by 1.6x)
      unsigned peers = kFull;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const unsigned m = __ballot_sync(kFull, (digit >> b) & 1u);
        peers &= ((digit >> b) & 1u) ? m : ~m;
      }
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (lane == leader) {
        old = hist[warp * 256 + digit];
        hist[warp * 256 + digit] = (uint16_t)(old + __popc(peers));
      }
      __syncwarp();  // the next word's leaders may be other lanes
      old = __shfl_sync(kFull, old, leader);
      off[e] = (int)old + __popc(peers & ltMask);
    }
    __syncthreads();
    {  // exclusive scan of the 256 x 32 counters in (digit, warp) order
      const int dg = tid >> 2, w0 = (tid & 3) * 8;
      uint32_t cc[8], sum = 0;
#pragma unroll
      for (int u = 0; u < 8; ++u) { cc[u] = hist[(w0 + u) * 256 + dg]; sum += cc[u]; }
      uint32_t inc = sum;
#pragma unroll
      for (int dl = 1; dl < 32; dl <<= 1) {
        const uint32_t x = __shfl_up_sync(kFull, inc, dl);
        if (lane >= dl) inc += x;
      }
      if (lane == 31) warpTot[warp] = (int)inc;
      __syncthreads();
      uint32_t wsum = (uint32_t)warpTot[lane], winc = wsum;
#pragma unroll
      for (int dl = 1; dl < 32; dl <<= 1) {
        const uint32_t x = __shfl_up_sync(kFull, winc, dl);
        if (lane >= dl) winc += x;
      }
      uint32_t base = __shfl_sync(kFull, winc - wsum, warp) + inc - sum;
#pragma unroll
      for (int u = 0; u < 8; ++u) { hist[(w0 + u) * 256 + dg] = (uint16_t)base; base += cc[u]; }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const uint32_t v = cur[posBase + e * 32];
      const uint32_t digit = (v >> shift) & 255u;
      nxt[(uint32_t)hist[warp * 256 + digit] + (uint32_t)off[e]] = v;
    }
    __syncthreads();
    for (int i = tid; i < 32 * 256 / 2; i += kThreads) reinterpret_cast<uint32_t *>(hist)[i] = 0u;
    // sparse table level (the block maxima do not depend on the sort: built in the shadow of its barriers)
    for (int lv = 1 + 4 * pass; lv < min(kLevels, 5 + 4 * pass); ++lv) {
      __syncthreads();
      const int half = 1 << (lv - 1);
      for (int b = tid; b < NB; b += kThreads) st[lv * NB + b] = max(st[(lv - 1) * NB + b], b + half < NB ? st[(lv - 1) * NB + b + half] : 0);
    }
    uint32_t *sw = cur; cur = nxt; nxt = sw;
    __syncthreads();
  }

  if (PS) {  // (after two passes the elements are back in buf0)
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = posBase + e * 32;
      int up = dS[i], dn = up;
#pragma unroll
      for (int dl = 1; dl < 32; dl <<= 1) {
        const int a = __shfl_up_sync(kFull, up, dl), b = __shfl_down_sync(kFull, dn, dl);
        if (lane >= dl) up = max(up, a);
        if (lane + dl < 32) dn = max(dn, b);
      }
      preS[i] = up;
      sufS[i] = dn;
    }
    __syncthreads();
  }

  // max of I's divergence over I-positions [lo, hi], 1 <= lo <= hi
  auto range_max = [&](int lo, int hi) {
    const int bl = lo >> 5, bh = hi >> 5;
    int m = 0;
    if (bl == bh) {
      for (int i = lo; i <= hi; ++i) m = max(m, dS[i]);
      return m;
    }
    if (PS) m = max(sufS[lo], preS[hi]);
    else {
      for (int i = lo; i < (bl + 1) * 32; ++i) m = max(m, dS[i]);
      for (int i = bh * 32; i <= hi; ++i) m = max(m, dS[i]);
    }
    const int b0 = bl + 1, cntB = bh - b0;  // whole blocks [b0, bh)
    if (cntB > 0) {
      const int lv = 31 - __clz(cntB);
      m = max(m, max(st[lv * NB + b0], st[lv * NB + bh - (1 << lv)]));
    }
    return m;
  };

  // ---- the new node: order, divergence, tie groups (make_groups_kernel's job, windows.cu) ----
  int hB[E], dB[E];
  int nflag = 0;  // group heads among this warp's positions so far
  uint32_t flagsW[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int p = posBase + e * 32;
    const uint32_t v = cur[p];
    const uint32_t key = v >> 16;
    int h = 0, d = 0;
    if (p < M) {
      const int ip = (int)(v & 0xffffu);
      h = __ldg(ordI + ip);
      if (p > 0) {
        const uint32_t pv = cur[p - 1];
        d = (pv >> 16) != key ? __ldg(gfdJ + key) : range_max((int)(pv & 0xffffu) + 1, ip);
      }
    }
    hB[e] = h;
    dB[e] = d;
    flagsW[e] = __ballot_sync(kFull, p < M && (p == 0 || d != 0));
    nflag += __popc(flagsW[e]);
  }
  if (lane == 0) warpTot[warp] = nflag;
  __syncthreads();
  int rankBase;  // group heads in the warps below this one
  {
    const int wv = warpTot[lane];
    int wx = wv;
#pragma unroll
    for (int dl = 1; dl < 32; dl <<= 1) {
      const int t = __shfl_up_sync(kFull, wx, dl);
      if (lane >= dl) wx += t;
    }
    rankBase = __shfl_sync(kFull, wx - wv, warp);
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int p = posBase + e * 32;
    const bool flag = (flagsW[e] >> lane) & 1u;
    if (p < M) {
      const int rank = rankBase + __popc(flagsW[e] & ltMask) + (flag ? 1 : 0) - 1;  // inclusive count of group heads, minus one
      ordOut[p] = hB[e];
      dvOut[p] = p == 0 ? 0 : dB[e];
      gOut[hB[e]] = rank;
      if (flag) gfdOut[rank] = dB[e];
    }
    rankBase += __popc(flagsW[e]);
  }
  if (tid == 0) dvOut[M] = 0;
}

template <int E>
size_t smem_bytes_e() { return (size_t)kThreads * E * (E <= 10 ? 16 : 12) + (size_t)kLevels * 32 * E * 4 + 32 * 256 * 2; }

template <int E>
int launch_e(const CombineSortArgs &c, int ctas, cudaStream_t stream) {
  const size_t bytes = smem_bytes_e<E>();
  static_assert(32 * E <= (1 << kLevels), "sparse table too shallow");
  auto kern = combine_sort_kernel<E>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (err != cudaSuccess) return (int)err;
  kern<<<ctas, kThreads, bytes, stream>>>(c);
  return (int)cudaGetLastError();
}

}  // namespace

int combine_sort_max_m() { return 13 * 1024; }

int launch_combine_sort(const CombineSortArgs &c, cudaStream_t stream) {
  if (c.cnt < 0 || c.ncopy < 0 || c.cnt + c.ncopy == 0) return 0;
  if (c.M < 1 || c.M > combine_sort_max_m()) return (int)cudaErrorInvalidValue;
  const int ctas = c.cnt + c.ncopy;
  const int e = ((c.M + 31) / 32 + 31) / 32;
  if (e <= 1) return launch_e<1>(c, ctas, stream);
  if (e <= 2) return launch_e<2>(c, ctas, stream);
  if (e <= 4) return launch_e<4>(c, ctas, stream);
  if (e <= 7) return launch_e<7>(c, ctas, stream);
  if (e <= 10) return launch_e<10>(c, ctas, stream);
  return launch_e<13>(c, ctas, stream);
}

}  // namespace pbwtgpu
