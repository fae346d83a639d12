// window_sort.cu -- pass 1 of the window-parallel build as a SORT instead of a sweep.
//
// What pass 1 needs from a window of sites [k0, k1) is only the node it contributes to the combine scan
// (windows.cu): the order of the haplotypes after the window when it is swept on its own from the panel's start
// order, and the divergence of neighbours in that order.  Swept site by site (pbwtCursorForwardsAD,
// pbwtCore.c:485-508) that costs one stable partition per site.  But the result is simply
//     order      = the start order, stably sorted by the window's alleles read backwards (site k1-1 most significant),
//     divergence = 1 + the most recent site at which the two neighbours differ (0: identical over the window),
// so it is computed here by an LSD radix sort of the haplotypes' window keys, 8 sites per pass (one CTA per window,
// everything resident in shared memory), followed by one top-down comparison of neighbouring keys -- about 7x fewer
// instructions than the sweep, with bit-identical output.  Pass 2 (which must visit every state) still sweeps.
//
//   transpose_kernel    panel bit columns (site-major) -> per window, per 16 sites, one uint16 per haplotype
//                       (32 x 32 bit tiles transposed with 5 shuffle stages)
//   window_sort_kernel  elements = (16 key bits << 16 | haplotype), double buffered in shared memory; the 16-site key
//                       column of the window is staged by TMA bulk copies (prefetched one column ahead); per 8-bit
//                       digit: per-warp histograms, block scan, stable scatter; then neighbour LCP, top column first
#include "sweep.cuh"
#include "windows.cuh"

namespace pbwtgpu {
namespace {

constexpr int kThreads = 1024;
constexpr uint32_t kPadId = 0xffffu;

__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane) {  // out(lane j) bit l = in(lane l) bit j
  uint32_t m = 0x0000ffffu;
#pragma unroll
  for (int j = 16; j; j >>= 1) {
    const uint32_t other = __shfl_xor_sync(kFull, x, j);
    x = (lane & j) ? ((x & (m << j)) | ((other >> j) & m)) : ((x & m) | ((other & m) << j));
    m ^= m << (j >> 1);
  }
  return x;
}

// T16[(win * nq + q) * mpad + h]: bit j = allele of haplotype h at site k0(win) + 16 q + j (0 beyond the window's end)
__global__ void __launch_bounds__(256) transpose_kernel(const uint32_t *panel, int wpc, int M, int N, int spw, int first, int second,
                                                        int nq, int mpad, uint16_t *T16) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int qw = blockIdx.x;  // 32 sites = two key columns
  const int win = blockIdx.z;
  int k0, k1;
  window_range(win, spw, first, second, N, k0, k1);
  const int site = k0 + 32 * qw + lane;
  const int hw0 = (blockIdx.y * 8 + warp) * 4;  // 4 panel words = 128 haplotypes per warp
  uint32_t in[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) in[u] = (site < k1 && hw0 + u < wpc) ? panel[(size_t)site * wpc + hw0 + u] : 0u;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const uint32_t mine = transpose32(in[u], lane);  // lane = haplotype (hw0+u)*32 + lane, bits = the 32 sites
    const int h = (hw0 + u) * 32 + lane;
    if (h < M) {
      T16[((size_t)win * nq + 2 * qw) * mpad + h] = (uint16_t)(mine & 0xffffu);
      if (2 * qw + 1 < nq) T16[((size_t)win * nq + 2 * qw + 1) * mpad + h] = (uint16_t)(mine >> 16);
    }
  }
}

template <int E>
__global__ void __launch_bounds__(kThreads, 1) window_sort_kernel(const uint16_t *T16, int nq, int mpad, int M, int N, int spw,
                                                                  int first, int second, int siteOffset,
                                                                  const int32_t *aStart, int32_t *ordOut, int32_t *dvOut) {
  constexpr int MP = kThreads * E;
  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t *buf0 = smem;  // (key16 << 16 | id), double buffered
  uint32_t *buf1 = buf0 + MP;
  uint16_t *hist = reinterpret_cast<uint16_t *>(buf1 + MP);          // [32 warps][256 digits]
  uint16_t *stage = hist + 32 * 256;                                   // [2][mpad] key columns
  __shared__ int warpTot[32];
  __shared__ __align__(8) uint64_t bar[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int win = blockIdx.x;
  int k0, k1;
  window_range(win, spw, first, second, N, k0, k1);
  const int nSites = k1 - k0;
  const int nCols = (nSites + 15) / 16;  // key columns in use (<= nq)
  const uint16_t *Tw = T16 + (size_t)win * nq * mpad;
  const uint32_t ltMask = lanemask_lt(lane);
  const int posBase = warp * E * 32 + lane;  // position of (word 0 of this warp, this lane)
  const uint32_t colBytes = (uint32_t)mpad * 2u;

  // column loads: number n < nCols reads column n (the sort, least significant first), n >= nCols reads column
  // 2*nCols-1-n (the neighbour comparison, most significant first); load n lands in stage[n & 1]
  auto issue = [&](int n) {
    const int q = n < nCols ? n : 2 * nCols - 1 - n;
    fence_proxy_async();
    mbar_expect_tx(&bar[n & 1], colBytes);
    tma_load_1d(stage + (size_t)(n & 1) * mpad, Tw + (size_t)q * mpad, colBytes, &bar[n & 1]);
  };
  auto wait = [&](int n) { mbar_wait(&bar[n & 1], (uint32_t)(n >> 1) & 1u); };

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  for (int i = tid; i < MP; i += kThreads) buf0[i] = i < M ? (uint32_t)(aStart ? aStart[i] : i) : (0xffff0000u | kPadId);
  for (int i = tid; i < 32 * 256 / 2; i += kThreads) reinterpret_cast<uint32_t *>(hist)[i] = 0u;
  __syncthreads();
  if (tid == 0 && nCols > 0) issue(0);

  uint32_t *cur = buf0, *nxt = buf1;
  for (int n = 0; n < nCols; ++n) {
    if (tid == 0) issue(n + 1);  // stage[(n+1)&1] was last read two barriers ago; (load nCols = top column, for the LCP)
    wait(n);
    const uint16_t *col = stage + (size_t)(n & 1) * mpad;
    for (int half = 0; half < 2; ++half) {
      if (16 * n + 8 * half >= nSites) break;  // uniform
      const int shift = 16 + 8 * half;
      int off[E];
      // ---- P1: digits, ranks inside the warp, per-warp histogram (words in order: stable) ----
#pragma unroll
      for (int e = 0; e < E; ++e) {
        uint32_t v = cur[posBase + e * 32];
        if (half == 0) {  // a new key column every second pass
          const uint32_t id = v & 0xffffu;
          v = id | ((id == kPadId ? 0xffffu : (uint32_t)col[id]) << 16);
          cur[posBase + e * 32] = v;
        }
        const uint32_t digit = (v >> shift) & 255u;
        const unsigned peers = __match_any_sync(kFull, digit);
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
      // ---- P2: exclusive scan of the 256 x 32 counters in (digit, warp) order ----
      {
        const int dg = tid >> 2, w0 = (tid & 3) * 8;  // thread tid: digit tid / 4, warps (tid % 4) * 8 .. + 7
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
        const uint32_t v = cur[posBase + e * 32];
        const uint32_t digit = (v >> shift) & 255u;
        nxt[(uint32_t)hist[warp * 256 + digit] + (uint32_t)off[e]] = v;
      }
      __syncthreads();
      for (int i = tid; i < 32 * 256 / 2; i += kThreads) reinterpret_cast<uint32_t *>(hist)[i] = 0u;
      uint32_t *sw = cur; cur = nxt; nxt = sw;
      __syncthreads();
    }
  }

  // ---- node: order, and divergence of neighbours = 1 + the most recent site at which their keys differ ----
  int32_t *ord = ordOut + (size_t)win * M;
  int32_t *dv = dvOut + (size_t)win * (M + 1);
  int d[E];
  uint32_t idA[E], idB[E];  // haplotype just above / at each of this thread's positions
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = posBase + e * 32;
    idB[e] = cur[i] & 0xffffu;
    idA[e] = i > 0 ? cur[i - 1] & 0xffffu : kPadId;
    d[e] = (i > 0 && i < M) ? 0 : -1;  // 0 = not separated so far, -1 = nothing to compute
    if (i < M) ord[i] = (int32_t)idB[e];
  }
  for (int n = nCols; n < 2 * nCols; ++n) {
    if (tid == 0 && n + 1 < 2 * nCols) issue(n + 1);
    wait(n);
    const uint16_t *col = stage + (size_t)(n & 1) * mpad;
    const int q = 2 * nCols - 1 - n;
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (d[e] == 0) {
        const uint32_t x = (uint32_t)col[idA[e]] ^ (uint32_t)col[idB[e]];
        if (x) d[e] = siteOffset + k0 + 16 * q + (31 - __clz(x)) + 1;
      }
    __syncthreads();  // everyone is done with this column before its buffer is refilled
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = posBase + e * 32;
    if (i < M) dv[i] = i == 0 ? siteOffset + k1 + 1 : d[e];
  }
  if (tid == 0) dv[M] = siteOffset + k1 + 1;
}

size_t sort_smem_bytes(int E, int mpad) { return (size_t)kThreads * E * 8 + 32 * 256 * 2 + (size_t)mpad * 4; }

template <int E>
int launch_sort_e(const uint16_t *T16, int nq, int mpad, int M, int N, int spw, int first, int second, int siteOffset,
                  const int32_t *aStart, int nodes, int32_t *ordOut, int32_t *dvOut, cudaStream_t stream) {
  const size_t bytes = sort_smem_bytes(E, mpad);
  auto kern = window_sort_kernel<E>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (err != cudaSuccess) return (int)err;
  kern<<<nodes, kThreads, bytes, stream>>>(T16, nq, mpad, M, N, spw, first, second, siteOffset, aStart, ordOut, dvOut);
  return (int)cudaGetLastError();
}

}  // namespace

int window_sort_max_m() { return 13 * 1024; }
static int key_cols(int spw) { return (spw + 15) / 16; }
static int padded_m(int M) { return (M + 7) / 8 * 8; }  // whole 16-byte units per key column (TMA bulk copy)
size_t window_sort_scratch_words(int M, int spw, int nodes) {
  return ((size_t)nodes * key_cols(spw) * padded_m(M) * 2 + 3) / 4;
}

int launch_window_sort(const uint32_t *panel, int wpc, int M, int N, int spw, int first, int second, int siteOffset,
                       const int32_t *aStart, int nodes, uint32_t *scratch, int32_t *ordOut, int32_t *dvOut, cudaStream_t stream) {
  if (nodes <= 0) return 0;
  if (M > window_sort_max_m()) return (int)cudaErrorInvalidValue;
  const int nq = key_cols(spw), mpad = padded_m(M);
  uint16_t *T16 = reinterpret_cast<uint16_t *>(scratch);
  {
    dim3 grid((spw + 31) / 32, (wpc + 31) / 32, nodes);
    transpose_kernel<<<grid, 256, 0, stream>>>(panel, wpc, M, N, spw, first, second, nq, mpad, T16);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return (int)err;
  }
  const int e = ((M + 31) / 32 + 31) / 32;
  if (e <= 1) return launch_sort_e<1>(T16, nq, mpad, M, N, spw, first, second, siteOffset, aStart, nodes, ordOut, dvOut, stream);
  if (e <= 2) return launch_sort_e<2>(T16, nq, mpad, M, N, spw, first, second, siteOffset, aStart, nodes, ordOut, dvOut, stream);
  if (e <= 4) return launch_sort_e<4>(T16, nq, mpad, M, N, spw, first, second, siteOffset, aStart, nodes, ordOut, dvOut, stream);
  if (e <= 7) return launch_sort_e<7>(T16, nq, mpad, M, N, spw, first, second, siteOffset, aStart, nodes, ordOut, dvOut, stream);
  if (e <= 10) return launch_sort_e<10>(T16, nq, mpad, M, N, spw, first, second, siteOffset, aStart, nodes, ordOut, dvOut, stream);
  return launch_sort_e<13>(T16, nq, mpad, M, N, spw, first, second, siteOffset, aStart, nodes, ordOut, dvOut, stream);
}

}  // namespace pbwtgpu
