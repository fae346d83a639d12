// window_sort_wide.cu -- pass 1 as a sort (see window_sort.cu) for panels wider than one CTA's shared memory:
// one window = one thread-block CLUSTER of 16 CTAs.  The haplotype ids live in the cluster's distributed shared
// memory (CTA r holds sorted positions [r*MPl, (r+1)*MPl)); per 8-site digit every CTA ranks its own elements
// (per-warp histograms, as in window_sort.cu), the per-CTA digit totals are exchanged through DSMEM, and the stable
// scatter writes each id straight into the destination CTA's shared memory (mapa + st.shared::cluster).  Two
// hardware cluster barriers per pass.  The 8-site digit column of the window (one byte per haplotype) is staged in
// every CTA by a TMA bulk copy that overlaps the scan and scatter of the previous pass.
// Divergences come from a separate pass over the row-major copy of the keys (neighbour LCP, most recent site first).
#include "sweep.cuh"
#include "windows.cuh"

namespace pbwtgpu {
namespace {

constexpr int kThreads = 1024;
constexpr int kCluster = 16;
constexpr uint32_t kPadId = 0xffffffu;

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

// T8[(win * nq8 + q) * mpad + h]: bit j = allele of haplotype h at site k0(win) + 8 q + j (digit columns);
// Trow[(win * M + h) * nw + w]: bit j = allele at site k0(win) + 32 w + j (rows, for the neighbour comparison)
__global__ void __launch_bounds__(256) transpose_wide_kernel(const uint32_t *panel, int wpc, int M, int N, int spw, int nq8,
                                                             int mpad, int nw, uint8_t *T8, uint32_t *Trow) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int qw = blockIdx.x;  // 32 sites
  const int win = blockIdx.z;
  const int k0 = win * spw, k1 = min(N, k0 + spw);
  const int site = k0 + 32 * qw + lane;
  const int hw0 = (blockIdx.y * 8 + warp) * 4;  // 4 panel words = 128 haplotypes per warp
  uint32_t in[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) in[u] = (site < k1 && hw0 + u < wpc) ? panel[(size_t)site * wpc + hw0 + u] : 0u;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const uint32_t mine = transpose32(in[u], lane);
    const int h = (hw0 + u) * 32 + lane;
    if (h < M) {
      Trow[((size_t)win * M + h) * nw + qw] = mine;
      if (T8) {
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (4 * qw + b < nq8) T8[((size_t)win * nq8 + 4 * qw + b) * mpad + h] = (uint8_t)(mine >> (8 * b));
      }
    }
  }
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_u32(const void *localSmem, unsigned ctaRank, uint32_t v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(localSmem)), "r"(ctaRank));
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(v) : "memory");
}

template <int E>
__global__ void __launch_bounds__(kThreads, 1) window_sort_cluster_kernel(const uint8_t *T8, int nq8, int mpad, int M, int N,
                                                                          int spw, const int32_t *aStart, int32_t *ordOut) {
  constexpr int MPl = kThreads * E;  // sorted positions held by one CTA
  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t *buf0 = smem, *buf1 = buf0 + MPl;                             // haplotype ids, double buffered
  uint16_t *hist = reinterpret_cast<uint16_t *>(buf1 + MPl);            // [32 warps][256 digits]
  uint32_t *ctaTot = reinterpret_cast<uint32_t *>(hist + 32 * 256);     // [16 CTAs][256 digits], written by the owners
  uint32_t *digBase = ctaTot + kCluster * 256;                          // [256] global base of this CTA's elements per digit
  uint8_t *stage = reinterpret_cast<uint8_t *>(digBase + 256);          // [mpad] digit column
  __shared__ uint32_t warpTot[32];
  __shared__ __align__(8) uint64_t bar;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int win = blockIdx.x / kCluster;
  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int k0 = win * spw, k1 = min(N, k0 + spw);
  const int nCols = (k1 - k0 + 7) / 8;
  const uint8_t *Tw = T8 + (size_t)win * nq8 * mpad;
  const uint32_t ltMask = lanemask_lt(lane);
  const int posBase = warp * E * 32 + lane;
  const uint32_t colBytes = (uint32_t)mpad;

  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  for (int i = tid; i < MPl; i += kThreads) {
    const int gp = (int)rank * MPl + i;
    buf0[i] = gp < M ? (uint32_t)(aStart ? aStart[gp] : gp) : kPadId;
  }
  for (int i = tid; i < 32 * 256 / 2; i += kThreads) reinterpret_cast<uint32_t *>(hist)[i] = 0u;
  __syncthreads();
  if (tid == 0 && nCols > 0) {
    fence_proxy_async();
    mbar_expect_tx(&bar, colBytes);
    tma_load_1d(stage, Tw, colBytes, &bar);
  }
  cluster_sync_all();  // every CTA of the cluster is running before anyone touches a peer's shared memory

  uint32_t *cur = buf0, *nxt = buf1;
  for (int n = 0; n < nCols; ++n) {
    mbar_wait(&bar, (uint32_t)n & 1u);
    int off[E];
    uint32_t dig[E];
    // ---- P1: digits, ranks inside the warp, per-warp histogram (words in order: stable) ----
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const uint32_t id = cur[posBase + e * 32];
      const uint32_t digit = id == kPadId ? 255u : (uint32_t)stage[id];
      dig[e] = digit;
      const unsigned peers = __match_any_sync(kFull, digit);
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (lane == leader) {
        old = hist[warp * 256 + digit];
        hist[warp * 256 + digit] = (uint16_t)(old + __popc(peers));
      }
      __syncwarp();
      old = __shfl_sync(kFull, old, leader);
      off[e] = (int)old + __popc(peers & ltMask);
    }
    __syncthreads();
    if (tid == 0 && n + 1 < nCols) {  // the digit column is consumed: fetch the next one behind the scan and scatter
      fence_proxy_async();
      mbar_expect_tx(&bar, colBytes);
      tma_load_1d(stage, Tw + (size_t)(n + 1) * mpad, colBytes, &bar);
    }
    // ---- P2a: per digit, this CTA's total and the exclusive prefix over its warps; totals go to every CTA ----
    if (tid < 256) {
      uint32_t run = 0;
#pragma unroll 8
      for (int w = 0; w < 32; ++w) {
        const uint32_t c = hist[w * 256 + tid];
        hist[w * 256 + tid] = (uint16_t)run;
        run += c;
      }
#pragma unroll
      for (int c = 0; c < kCluster; ++c) st_cluster_u32(&ctaTot[rank * 256 + tid], (unsigned)c, run);
    }
    cluster_sync_all();
    // ---- P2b: global base of this CTA's elements of each digit ----
    if (tid < 256) {
      uint32_t colSum = 0, before = 0;
#pragma unroll
      for (int c = 0; c < kCluster; ++c) {
        const uint32_t t = ctaTot[c * 256 + tid];
        colSum += t;
        if (c < (int)rank) before += t;
      }
      uint32_t inc = colSum;
#pragma unroll
      for (int dl = 1; dl < 32; dl <<= 1) {
        const uint32_t x = __shfl_up_sync(kFull, inc, dl);
        if (lane >= dl) inc += x;
      }
      if (lane == 31) warpTot[warp] = inc;
      asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 warps that hold the digits
      uint32_t wb = 0;
      for (int w = 0; w < warp; ++w) wb += warpTot[w];
      digBase[tid] = wb + inc - colSum + before;
    }
    __syncthreads();
    // ---- P3: stable scatter into the owning CTA's shared memory ----
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const uint32_t id = cur[posBase + e * 32];
      const uint32_t pos = digBase[dig[e]] + (uint32_t)hist[warp * 256 + dig[e]] + (uint32_t)off[e];
      st_cluster_u32(&nxt[pos % MPl], pos / MPl, id);
    }
    cluster_sync_all();  // every id of the pass has landed
    for (int i = tid; i < 32 * 256 / 2; i += kThreads) reinterpret_cast<uint32_t *>(hist)[i] = 0u;
    uint32_t *sw = cur; cur = nxt; nxt = sw;
    __syncthreads();
  }
  int32_t *ord = ordOut + (size_t)win * M;
  for (int i = tid; i < MPl; i += kThreads) {
    const int gp = (int)rank * MPl + i;
    if (gp < M) ord[gp] = (int32_t)cur[i];
  }
}

// ---- very wide panels (beyond the cluster's shared memory): the same sort with the elements in global memory ----
// One window = a group of G co-resident CTAs (cooperative launch, as in sweep_global.cu), CTA g ranks the elements at
// sorted positions [g*chunk, (g+1)*chunk).  Elements are (id, 32 key bits) pairs in an L2-resident double buffer;
// the key word is refreshed from the row-major keys every fourth pass; per-CTA digit totals go through a small
// global table; two group barriers per pass.
struct SortGroupSync {
  unsigned int arrive;
  unsigned int pad[31];
};

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(kThreads, 1) window_sort_global_kernel(const uint32_t *Trow, int nw, int M, int N, int spw,
                                                                         const int32_t *aStart, uint2 *bufAll,
                                                                         uint32_t *ctaTotAll, SortGroupSync *gsync, int G,
                                                                         int chunk, int32_t *ordOut) {
  extern __shared__ __align__(16) uint32_t smem[];
  uint16_t *hist = reinterpret_cast<uint16_t *>(smem);                    // [32 warps][256 digits]
  uint32_t *digBase = smem + 32 * 256 / 2;                                // [256]
  uint16_t *offS = reinterpret_cast<uint16_t *>(digBase + 256);           // [chunk] rank of each element inside its warp
  __shared__ uint32_t warpTot[32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int win = blockIdx.x / G, g = blockIdx.x % G;
  const int k0 = win * spw, k1 = min(N, k0 + spw);
  const int nDigits = (k1 - k0 + 7) / 8;
  const size_t MS = (size_t)G * chunk;
  uint2 *cur = bufAll + (size_t)win * 2 * MS, *nxt = cur + MS;
  uint32_t *ctaTot = ctaTotAll + (size_t)win * G * 256;
  SortGroupSync *gs = gsync + win;
  const uint32_t *Tw = Trow + (size_t)win * M * nw;
  const uint32_t ltMask = lanemask_lt(lane);
  const int EW = chunk / 1024;                     // words per warp
  const int posBase = g * chunk + warp * EW * 32;  // first sorted position of this warp
  unsigned epoch = 0;

  auto group_barrier = [&]() {
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(&gs->arrive, 1u);
      const unsigned target = (epoch + 1) * (unsigned)G;
      while (ld_volatile_u32(&gs->arrive) < target) {}
      __threadfence();
    }
    ++epoch;
    __syncthreads();
  };

  for (int i = g * chunk + tid; i < (g + 1) * chunk; i += kThreads)
    if (i < M) cur[i] = make_uint2((uint32_t)(aStart ? aStart[i] : i), 0u);
  for (int i = tid; i < 32 * 256 / 2; i += kThreads) smem[i] = 0u;
  group_barrier();

  for (int t = 0; t < nDigits; ++t) {
    const int shift = (t & 3) * 8;
    // ---- P1: digits, ranks inside the warp, per-warp histogram (words in order: stable) ----
    for (int e0 = 0; e0 < EW; e0 += 4) {
      uint2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = posBase + (e0 + u) * 32 + lane;
        v[u] = (e0 + u < EW && i < M) ? __ldcg(cur + i) : make_uint2(kPadId, 0xffffffffu);
      }
      if (shift == 0) {  // a new key word every fourth pass
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = posBase + (e0 + u) * 32 + lane;
          if (e0 + u < EW && i < M) {
            v[u].y = __ldg(Tw + (size_t)v[u].x * nw + (t >> 2));
            cur[i].y = v[u].y;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (e0 + u >= EW) break;
        const int i = posBase + (e0 + u) * 32 + lane;
        const bool valid = i < M;
        const uint32_t digit = valid ? (v[u].y >> shift) & 255u : 256u;
        const unsigned peers = __match_any_sync(kFull, digit);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && lane == leader) {
          old = hist[warp * 256 + digit];
          hist[warp * 256 + digit] = (uint16_t)(old + __popc(peers));
        }
        __syncwarp();
        old = __shfl_sync(kFull, old, leader);
        offS[(warp * EW + e0 + u) * 32 + lane] = (uint16_t)(old + __popc(peers & ltMask));
      }
    }
    __syncthreads();
    // ---- P2a: per digit, this CTA's total and the exclusive prefix over its warps; totals to the group's table ----
    if (tid < 256) {
      uint32_t run = 0;
#pragma unroll 8
      for (int w = 0; w < 32; ++w) {
        const uint32_t c = hist[w * 256 + tid];
        hist[w * 256 + tid] = (uint16_t)run;
        run += c;
      }
      ctaTot[g * 256 + tid] = run;
    }
    group_barrier();
    // ---- P2b: global base of this CTA's elements of each digit ----
    if (tid < 256) {
      uint32_t colSum = 0, before = 0;
      for (int c = 0; c < G; ++c) {
        const uint32_t x = __ldcg(ctaTot + c * 256 + tid);
        colSum += x;
        if (c < g) before += x;
      }
      uint32_t inc = colSum;
#pragma unroll
      for (int dl = 1; dl < 32; dl <<= 1) {
        const uint32_t x = __shfl_up_sync(kFull, inc, dl);
        if (lane >= dl) inc += x;
      }
      if (lane == 31) warpTot[warp] = inc;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      uint32_t wb = 0;
      for (int w = 0; w < warp; ++w) wb += warpTot[w];
      digBase[tid] = wb + inc - colSum + before;
    }
    __syncthreads();
    // ---- P3: stable scatter ----
    for (int e0 = 0; e0 < EW; e0 += 4) {
      uint2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = posBase + (e0 + u) * 32 + lane;
        v[u] = (e0 + u < EW && i < M) ? __ldcg(cur + i) : make_uint2(kPadId, 0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = posBase + (e0 + u) * 32 + lane;
        if (e0 + u < EW && i < M) {
          const uint32_t digit = (v[u].y >> shift) & 255u;
          const uint32_t pos = digBase[digit] + (uint32_t)hist[warp * 256 + digit] + (uint32_t)offS[(warp * EW + e0 + u) * 32 + lane];
          nxt[pos] = v[u];
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < 32 * 256 / 2; i += kThreads) smem[i] = 0u;
    group_barrier();  // every element of the pass has landed (and the totals table may be rewritten)
    uint2 *sw = cur; cur = nxt; nxt = sw;
  }
  int32_t *ord = ordOut + (size_t)win * M;
  for (int i = g * chunk + tid; i < (g + 1) * chunk; i += kThreads)
    if (i < M) ord[i] = (int32_t)__ldcg(cur + i).x;
}

// divergence of neighbours in the sorted order: 1 + the most recent site at which their key rows differ
__global__ void __launch_bounds__(256) neighbour_lcp_kernel(const uint32_t *Trow, int nw, int M, int N, int spw, int siteOffset,
                                                            const int32_t *ordAll, int32_t *dvAll) {
  const int win = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > M) return;
  const int k0 = win * spw, k1 = min(N, k0 + spw);
  int32_t *dv = dvAll + (size_t)win * (M + 1);
  if (i == 0 || i == M) { dv[i] = siteOffset + k1 + 1; return; }
  const int32_t *ord = ordAll + (size_t)win * M;
  const uint32_t *ra = Trow + ((size_t)win * M + ord[i - 1]) * nw, *rb = Trow + ((size_t)win * M + ord[i]) * nw;
  int d = 0;
  for (int q = (k1 - k0 + 31) / 32 - 1; q >= 0; --q) {
    const uint32_t x = __ldg(ra + q) ^ __ldg(rb + q);
    if (x) { d = siteOffset + k0 + 32 * q + (31 - __clz(x)) + 1; break; }
  }
  dv[i] = d;
}

size_t wide_smem_bytes(int E, int mpad) { return (size_t)kThreads * E * 8 + 32 * 256 * 2 + kCluster * 256 * 4 + 256 * 4 + (size_t)mpad; }

template <int E>
int launch_wide_e(bool queryOnly, const uint8_t *T8, int nq8, int mpad, int M, int N, int spw, const int32_t *aStart, int nodes,
                  int32_t *ordOut, cudaStream_t stream) {
  const size_t bytes = wide_smem_bytes(E, mpad);
  void *kern = (void *)window_sort_cluster_kernel<E>;
  if (bytes > 232448 - 1024) return (int)cudaErrorInvalidValue;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (err == cudaSuccess) err = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  if (err != cudaSuccess) { cudaGetLastError(); return (int)err; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(std::max(nodes, 1) * kCluster);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (queryOnly) {
    int n = 0;
    err = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (err != cudaSuccess) { cudaGetLastError(); return (int)err; }
    return n >= 1 ? 0 : (int)cudaErrorInvalidValue;
  }
  void *params[] = {&T8, &nq8, &mpad, &M, &N, &spw, &aStart, &ordOut};
  return (int)cudaLaunchKernelExC(&cfg, kern, params);
}

int dispatch_wide(bool queryOnly, const uint8_t *T8, int nq8, int mpad, int M, int N, int spw, const int32_t *aStart, int nodes,
                  int32_t *ordOut, cudaStream_t stream) {
  const int e = (M + kCluster * kThreads - 1) / (kCluster * kThreads);
  if (e <= 1) return launch_wide_e<1>(queryOnly, T8, nq8, mpad, M, N, spw, aStart, nodes, ordOut, stream);
  if (e <= 2) return launch_wide_e<2>(queryOnly, T8, nq8, mpad, M, N, spw, aStart, nodes, ordOut, stream);
  if (e <= 4) return launch_wide_e<4>(queryOnly, T8, nq8, mpad, M, N, spw, aStart, nodes, ordOut, stream);
  if (e <= 8) return launch_wide_e<8>(queryOnly, T8, nq8, mpad, M, N, spw, aStart, nodes, ordOut, stream);
  return (int)cudaErrorInvalidValue;
}

int digit_cols(int spw) { return (spw + 7) / 8; }
int row_words(int spw) { return (spw + 31) / 32; }
int padded_m16(int M) { return (M + 15) / 16 * 16; }  // whole 16-byte units per digit column (TMA bulk copy)

}  // namespace

// plan of the global-memory variant: G CTAs of `chunk` positions per window, all W windows co-resident
static void global_sort_plan(int M, int maxCtas, int W, int *G, int *chunk) {
  int g = std::max(1, maxCtas / std::max(W, 1));
  int ch = ((M + g - 1) / g + 1023) / 1024 * 1024;
  if (ch > 32768) ch = 32768;
  g = (M + ch - 1) / ch;
  *G = g; *chunk = ch;
}

bool window_sort_global_supported(int M, int W, int maxCtas) {
  int G, chunk;
  global_sort_plan(M, maxCtas, W, &G, &chunk);
  return M >= 2 && (long long)W * G <= maxCtas && (long long)M < (long long)kPadId;
}

size_t window_sort_global_scratch_words(int M, int spw, int W, int maxCtas) {
  int G, chunk;
  global_sort_plan(M, maxCtas, W, &G, &chunk);
  const size_t MS = (size_t)G * chunk;
  return (size_t)W * M * row_words(spw) + (size_t)W * 2 * MS * 2 + (size_t)W * G * 256 + (size_t)W * 32 + 256;
}

int launch_window_sort_global(const uint32_t *panel, int wpc, int M, int N, int spw, int siteOffset, const int32_t *aStart, int W,
                              int maxCtas, uint32_t *scratch, int32_t *ordOut, int32_t *dvOut, cudaStream_t stream) {
  if (W <= 0) return 0;
  int G, chunk;
  global_sort_plan(M, maxCtas, W, &G, &chunk);
  const int nw = row_words(spw);
  const size_t MS = (size_t)G * chunk;
  uint32_t *Trow = scratch;
  uint2 *buf = reinterpret_cast<uint2 *>(scratch + (((size_t)W * M * nw + 3) / 4) * 4);
  uint32_t *ctaTot = reinterpret_cast<uint32_t *>(buf + (size_t)W * 2 * MS);
  SortGroupSync *gsync = reinterpret_cast<SortGroupSync *>(ctaTot + (size_t)W * G * 256);
  {
    dim3 grid((spw + 31) / 32, (wpc + 31) / 32, W);
    transpose_wide_kernel<<<grid, 256, 0, stream>>>(panel, wpc, M, N, spw, 0, 0, nw, nullptr, Trow);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess) err = cudaMemsetAsync(gsync, 0, sizeof(SortGroupSync) * (size_t)W, stream);
    if (err != cudaSuccess) return (int)err;
  }
  const size_t smemBytes = 32 * 256 * 2 + 256 * 4 + (size_t)chunk * 2;
  void *kern = (void *)window_sort_global_kernel;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
  if (err != cudaSuccess) return (int)err;
  void *params[] = {&Trow, (void *)&nw, &M, &N, &spw, &aStart, &buf, &ctaTot, &gsync, &G, &chunk, &ordOut};
  err = cudaLaunchCooperativeKernel(kern, dim3(W * G), dim3(kThreads), params, smemBytes, stream);
  if (err != cudaSuccess) return (int)err;
  {
    dim3 grid((M + 1 + 255) / 256, W);
    neighbour_lcp_kernel<<<grid, 256, 0, stream>>>(Trow, nw, M, N, spw, siteOffset, ordOut, dvOut);
  }
  return (int)cudaGetLastError();
}

bool window_sort_wide_supported(int M, int spw) {
  if (M < 2 || (long long)M >= (long long)kPadId) return false;
  return dispatch_wide(true, nullptr, digit_cols(spw), padded_m16(M), M, 1, spw, nullptr, 1, nullptr, nullptr) == 0;
}

size_t window_sort_wide_scratch_words(int M, int spw, int nodes) {
  const size_t t8 = ((size_t)nodes * digit_cols(spw) * padded_m16(M) + 3) / 4;
  return t8 + (size_t)nodes * M * row_words(spw) + 64;
}

int launch_window_sort_wide(const uint32_t *panel, int wpc, int M, int N, int spw, int siteOffset, const int32_t *aStart, int nodes,
                            uint32_t *scratch, int32_t *ordOut, int32_t *dvOut, cudaStream_t stream) {
  if (nodes <= 0) return 0;
  const int nq8 = digit_cols(spw), mpad = padded_m16(M), nw = row_words(spw);
  uint8_t *T8 = reinterpret_cast<uint8_t *>(scratch);
  const size_t t8Words = (((size_t)nodes * nq8 * mpad + 3) / 4 + 3) / 4 * 4;
  uint32_t *Trow = scratch + t8Words;
  {
    dim3 grid((spw + 31) / 32, (wpc + 31) / 32, nodes);
    transpose_wide_kernel<<<grid, 256, 0, stream>>>(panel, wpc, M, N, spw, nq8, mpad, nw, T8, Trow);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return (int)err;
  }
  if (int rc = dispatch_wide(false, T8, nq8, mpad, M, N, spw, aStart, nodes, ordOut, stream)) return rc;
  {
    dim3 grid((M + 1 + 255) / 256, nodes);
    neighbour_lcp_kernel<<<grid, 256, 0, stream>>>(Trow, nw, M, N, spw, siteOffset, ordOut, dvOut);
  }
  return (int)cudaGetLastError();
}

}  // namespace pbwtgpu
