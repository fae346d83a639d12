// window_sort.cu -- pass 1 of the window-parallel build as a SORT instead of a sweep.
//
// What pass 1 needs from a window of sites [k0, k1) is only the node it contributes to the combine scan
// (windows.cu): the order of the haplotypes after the window when it is swept on its own from the panel's start
// order, and the divergence of neighbours in that order.  Swept site by site (pbwtCursorForwardsAD,
// pbwtCore.c:485-508) that costs one stable partition per site.  But the result is simply
//     order      = the start order, stably sorted by the window's alleles read backwards (site k1-1 most significant),
//     divergence = 1 + the most recent site at which the two neighbours differ (0: identical over the window),
// so it is computed here by an LSD radix sort of the haplotypes' window keys, 8 sites per pass (one CTA per window,
// keys and ids resident in shared memory), followed by one comparison of neighbouring keys -- about 7x fewer
// instructions than the sweep, with bit-identical output.  Pass 2 (which must visit every state) still sweeps.
//
//   transpose_kernel   panel bit columns (site-major) -> per-window haplotype-major key rows (32 x 32 bit tiles, ballots)
//   window_sort_kernel LSD radix sort, 8-bit digits: per-warp histograms, block scan, stable scatter; neighbour LCP
#include "sweep.cuh"
#include "windows.cuh"

namespace pbwtgpu {
namespace {

constexpr int kThreads = 1024;
constexpr unsigned kPadId = 0xffffffffu;

// T[(win * M + h) * nw + q]: bit j = allele of haplotype h at site k0(win) + 32 q + j (0 beyond the window's end)
__global__ void __launch_bounds__(256) transpose_kernel(const uint32_t *panel, int wpc, int M, int N, int spw, int nw,
                                                        uint32_t *T) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.x * 8 + warp;  // word of the key row
  const int win = blockIdx.z;
  if (q >= nw) return;
  const int k0 = win * spw, k1 = min(N, k0 + spw);
  const int site = k0 + 32 * q + lane;
  const int hw0 = blockIdx.y * 8;  // 8 panel words = 256 haplotypes per CTA
  uint32_t in[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) in[u] = (site < k1 && hw0 + u < wpc) ? panel[(size_t)site * wpc + hw0 + u] : 0u;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const uint32_t b = __ballot_sync(kFull, (in[u] >> j) & 1u);
      if (lane == j) mine = b;
    }
    const int h = (hw0 + u) * 32 + lane;
    if (h < M) T[((size_t)win * M + h) * nw + q] = mine;
  }
}

template <int E>
__global__ void __launch_bounds__(kThreads, 1) window_sort_kernel(const uint32_t *T, int nw, int M, int N, int spw,
                                                                  int siteOffset, const int32_t *aStart, int32_t *ordOut,
                                                                  int32_t *dvOut) {
  constexpr int MP = kThreads * E;
  extern __shared__ __align__(16) uint32_t smem[];
  uint2 *buf0 = reinterpret_cast<uint2 *>(smem);  // (id, current key word), double buffered
  uint2 *buf1 = buf0 + MP;
  uint16_t *hist = reinterpret_cast<uint16_t *>(buf1 + MP);  // [32 warps][256 digits]
  __shared__ int warpTot[32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int win = blockIdx.x;
  const int k0 = win * spw, k1 = min(N, k0 + spw);
  const int nSites = k1 - k0;
  const int nDigits = (nSites + 7) / 8;
  const uint32_t *Tw = T + (size_t)win * M * nw;
  const uint32_t ltMask = lanemask_lt(lane);
  const int posBase = warp * E * 32 + lane;  // position of (word 0 of this warp, this lane)

  for (int i = tid; i < MP; i += kThreads) buf0[i] = make_uint2(i < M ? (uint32_t)(aStart ? aStart[i] : i) : kPadId, 0u);
  for (int i = tid; i < 32 * 256 / 2; i += kThreads) reinterpret_cast<uint32_t *>(hist)[i] = 0u;
  __syncthreads();

  uint2 *cur = buf0, *nxt = buf1;
  for (int t = 0; t < nDigits; ++t) {
    const int shift = (t & 3) * 8;
    uint16_t off[E];
    // ---- P1: digits, ranks inside the warp, per-warp histogram (words in order: stable) ----
#pragma unroll
    for (int e = 0; e < E; ++e) {
      uint2 v = cur[posBase + e * 32];
      if (shift == 0) {  // a new key word every 4 passes
        v.y = v.x == kPadId ? 0xffffffffu : __ldg(Tw + (size_t)v.x * nw + (t >> 2));
        cur[posBase + e * 32] = v;
      }
      const uint32_t digit = (v.y >> shift) & 255u;
      const unsigned peers = __match_any_sync(kFull, digit);
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (lane == leader) {
        old = hist[warp * 256 + digit];
        hist[warp * 256 + digit] = (uint16_t)(old + __popc(peers));
      }
      __syncwarp();  // the next word's leaders may be other lanes
      old = __shfl_sync(kFull, old, leader);
      off[e] = (uint16_t)(old + __popc(peers & ltMask));
    }
    __syncthreads();
    // ---- P2: exclusive scan of the 256 x 32 counters in (digit, warp) order ----
    {
      // thread tid covers digit tid / 4, warps (tid % 4) * 8 .. + 7
      const int dg = tid >> 2, w0 = (tid & 3) * 8;
      uint32_t c[8], sum = 0;
#pragma unroll
      for (int u = 0; u < 8; ++u) { c[u] = hist[(w0 + u) * 256 + dg]; sum += c[u]; }
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
      for (int u = 0; u < 8; ++u) { hist[(w0 + u) * 256 + dg] = (uint16_t)base; base += c[u]; }
    }
    __syncthreads();
    // ---- P3: stable scatter ----
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const uint2 v = cur[posBase + e * 32];
      const uint32_t digit = (v.y >> shift) & 255u;
      nxt[(uint32_t)hist[warp * 256 + digit] + off[e]] = v;
    }
    __syncthreads();
    for (int i = tid; i < 32 * 256 / 2; i += kThreads) reinterpret_cast<uint32_t *>(hist)[i] = 0u;
    uint2 *sw = cur; cur = nxt; nxt = sw;
    __syncthreads();
  }

  // ---- node: order, and divergence of neighbours = 1 + the most recent site at which their keys differ ----
  int32_t *ord = ordOut + (size_t)win * M;
  int32_t *dv = dvOut + (size_t)win * (M + 1);
  const int nWordsUsed = (nSites + 31) / 32;
  for (int i = tid; i < M; i += kThreads) {
    const uint32_t h = cur[i].x;
    ord[i] = (int32_t)h;
    int d = 0;
    if (i > 0) {
      const uint32_t *ra = Tw + (size_t)cur[i - 1].x * nw, *rb = Tw + (size_t)h * nw;
      for (int q = nWordsUsed - 1; q >= 0; --q) {
        const uint32_t x = __ldg(ra + q) ^ __ldg(rb + q);
        if (x) { d = siteOffset + k0 + 32 * q + (31 - __clz(x)) + 1; break; }
      }
    } else d = siteOffset + k1 + 1;
    dv[i] = d;
  }
  if (tid == 0) dv[M] = siteOffset + k1 + 1;
}

template <int E>
int launch_sort_e(const uint32_t *T, int nw, int M, int N, int spw, int siteOffset, const int32_t *aStart, int nodes,
                  int32_t *ordOut, int32_t *dvOut, cudaStream_t stream) {
  const size_t bytes = (size_t)kThreads * E * 16 + 32 * 256 * 2;
  static_assert((size_t)kThreads * E * 16 + 32 * 256 * 2 <= 232448, "shared memory layout exceeds the 227 KB per-CTA limit");
  auto kern = window_sort_kernel<E>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (err != cudaSuccess) return (int)err;
  kern<<<nodes, kThreads, bytes, stream>>>(T, nw, M, N, spw, siteOffset, aStart, ordOut, dvOut);
  return (int)cudaGetLastError();
}

}  // namespace

int window_sort_max_m() { return 12 * 1024; }
int window_sort_key_words(int spw) { return (spw + 31) / 32; }

int launch_window_sort(const uint32_t *panel, int wpc, int M, int N, int spw, int siteOffset, const int32_t *aStart, int nodes,
                       uint32_t *T, int32_t *ordOut, int32_t *dvOut, cudaStream_t stream) {
  if (nodes <= 0) return 0;
  if (M > window_sort_max_m()) return (int)cudaErrorInvalidValue;
  const int nw = window_sort_key_words(spw);
  {
    dim3 grid((nw + 7) / 8, (wpc + 7) / 8, nodes);
    transpose_kernel<<<grid, 256, 0, stream>>>(panel, wpc, M, N, spw, nw, T);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return (int)err;
  }
  const int e = ((M + 31) / 32 + 31) / 32;
  if (e <= 1) return launch_sort_e<1>(T, nw, M, N, spw, siteOffset, aStart, nodes, ordOut, dvOut, stream);
  if (e <= 2) return launch_sort_e<2>(T, nw, M, N, spw, siteOffset, aStart, nodes, ordOut, dvOut, stream);
  if (e <= 4) return launch_sort_e<4>(T, nw, M, N, spw, siteOffset, aStart, nodes, ordOut, dvOut, stream);
  if (e <= 7) return launch_sort_e<7>(T, nw, M, N, spw, siteOffset, aStart, nodes, ordOut, dvOut, stream);
  if (e <= 10) return launch_sort_e<10>(T, nw, M, N, spw, siteOffset, aStart, nodes, ordOut, dvOut, stream);
  return launch_sort_e<12>(T, nw, M, N, spw, siteOffset, aStart, nodes, ordOut, dvOut, stream);
}

}  // namespace pbwtgpu
