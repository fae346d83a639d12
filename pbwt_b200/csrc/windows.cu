// windows.cu -- exact (a, d) at the start of every site window, so that windows can be swept
// concurrently on different SMs (the site recurrence a_k -> a_{k+1} is otherwise strictly sequential).
//
// Idea.  The order at a window boundary is the stable sort of the order at the previous boundary by the
// window's own key (its alleles, most recent site most significant).  A window swept on its own from the
// panel's start order yields that key's rank structure: its final order, and its final divergence array
// whose zero entries mark haplotypes the window cannot tell apart ("tie groups").  Two adjacent windows
// I (earlier), J (later) combine associatively:
//     order(I+J)      = order(I) stably re-sorted by J's tie-group index,
//     divergence(I+J) = J's boundary divergence between different J-groups, and inside one J-group the
//                       maximum of I's divergence over the I-positions in between.
// That is exactly a PBWT sweep of order(I), divergence(I) over pseudo-sites made of the bits of J's group
// index (least significant first): the AD update rule carries the range maxima, and every pair separated by
// a pseudo-site ends up with a label >= kPseudoLabelBase, which marks it for replacement by J's boundary
// divergence.  So the combine reuses the sweep kernel, and an inclusive Hillis-Steele scan over the windows
// (log2 W rounds, all windows in parallel) yields the exact state at every boundary.
#include "windows.cuh"

#include <algorithm>
#include <climits>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

namespace pbwtgpu {
namespace {

__global__ void __launch_bounds__(1024) make_groups_kernel(const int32_t *ordAll, int32_t *dvAll, const int32_t *gJAll,
                                                           const int32_t *gfdJAll, int32_t *gAll, int32_t *gfdAll, int M) {
  __shared__ int warpSums[32];
  __shared__ int carryS;
  const int node = blockIdx.x;
  const int32_t *ord = ordAll + (size_t)node * M;
  int32_t *dv = dvAll + (size_t)node * (M + 1);
  const int32_t *gJ = gJAll ? gJAll + (size_t)node * M : nullptr;
  const int32_t *gfdJ = gfdJAll ? gfdJAll + (size_t)node * M : nullptr;
  int32_t *g = gAll + (size_t)node * M;
  int32_t *gfd = gfdAll + (size_t)node * M;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carryS = 0;
  __syncthreads();
  for (int base = 0; base < M; base += 1024) {
    const int i = base + tid;
    int h = 0, d = 0, flag = 0;
    if (i < M) {
      h = ord[i];
      d = i == 0 ? 0 : dv[i];
      if (d >= kPseudoLabelBase) {  // separated by a pseudo-site: different groups of the later window
        d = gfdJ[gJ[h]];
        dv[i] = d;
      }
      flag = i == 0 || d != 0;
    }
    int x = flag;
#pragma unroll
    for (int dl = 1; dl < 32; dl <<= 1) {
      const int t = __shfl_up_sync(kFull, x, dl);
      if (lane >= dl) x += t;
    }
    if (lane == 31) warpSums[warp] = x;
    __syncthreads();
    int wv = warpSums[lane], wx = wv;
#pragma unroll
    for (int dl = 1; dl < 32; dl <<= 1) {
      const int t = __shfl_up_sync(kFull, wx, dl);
      if (lane >= dl) wx += t;
    }
    const int warpOff = __shfl_sync(kFull, wx - wv, warp);
    const int blockTotal = __shfl_sync(kFull, wx, 31);
    const int carry = carryS;
    if (i < M) {
      const int rank = carry + warpOff + x - 1;  // inclusive count of group heads, minus one
      g[h] = rank;
      if (flag) gfd[rank] = d;
    }
    __syncthreads();
    if (tid == 0) carryS = carry + blockTotal;
    __syncthreads();
  }
}

// The same for panels of up to 16 Ki haplotypes in ONE pass: every thread keeps its (up to 16) positions in registers,
// the loads of all chunks are in flight together, and the ranks come from per-(chunk, warp) ballot counts with a single
// scan -- two block barriers instead of three per 1024 positions.
constexpr int kGroupsFastChunks = 16;

__global__ void __launch_bounds__(1024) make_groups_fast_kernel(const int32_t *ordAll, int32_t *dvAll, const int32_t *gJAll,
                                                                const int32_t *gfdJAll, int32_t *gAll, int32_t *gfdAll, int M) {
  __shared__ int wcnt[kGroupsFastChunks * 32];
  const int node = blockIdx.x;
  const int32_t *ord = ordAll + (size_t)node * M;
  int32_t *dv = dvAll + (size_t)node * (M + 1);
  const int32_t *gJ = gJAll ? gJAll + (size_t)node * M : nullptr;
  const int32_t *gfdJ = gfdJAll ? gfdJAll + (size_t)node * M : nullptr;
  int32_t *g = gAll + (size_t)node * M;
  int32_t *gfd = gfdAll + (size_t)node * M;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nChunks = (M + 1023) / 1024;
  int h[kGroupsFastChunks], d[kGroupsFastChunks];
#pragma unroll
  for (int c = 0; c < kGroupsFastChunks; ++c) {
    const int i = c * 1024 + tid;
    const bool valid = c < nChunks && i < M;
    h[c] = valid ? ord[i] : 0;
    d[c] = (valid && i > 0) ? dv[i] : 0;
  }
#pragma unroll
  for (int c = 0; c < kGroupsFastChunks; ++c)
    if (d[c] >= kPseudoLabelBase) {  // separated by a pseudo-site: different groups of the later window
      d[c] = gfdJ[gJ[h[c]]];
      dv[c * 1024 + tid] = d[c];
    }
#pragma unroll
  for (int c = 0; c < kGroupsFastChunks; ++c) {
    if (c < nChunks) {  // uniform
      const int i = c * 1024 + tid;
      const unsigned bal = __ballot_sync(kFull, i < M && (i == 0 || d[c] != 0));
      if (lane == 0) wcnt[c * 32 + warp] = __popc(bal);
    }
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan of the nChunks*32 counts in position order; lane l owns nChunks consecutive entries
    int sum = 0;
    for (int j = 0; j < nChunks; ++j) sum += wcnt[lane * nChunks + j];
    int inc = sum;
#pragma unroll
    for (int dl = 1; dl < 32; dl <<= 1) {
      const int t = __shfl_up_sync(kFull, inc, dl);
      if (lane >= dl) inc += t;
    }
    int run = inc - sum;
    for (int j = 0; j < nChunks; ++j) {
      const int v = wcnt[lane * nChunks + j];
      wcnt[lane * nChunks + j] = run;
      run += v;
    }
  }
  __syncthreads();
  const unsigned ltMask = (1u << lane) - 1u;
#pragma unroll
  for (int c = 0; c < kGroupsFastChunks; ++c) {
    const int i = c * 1024 + tid;
    if (c < nChunks) {  // uniform
      const bool flag = i < M && (i == 0 || d[c] != 0);
      const unsigned bal = __ballot_sync(kFull, flag);  // (recomputed rather than kept: 16 more registers would spill)
      if (i < M) {
        const int rank = wcnt[c * 32 + warp] + __popc(bal & ltMask) + (flag ? 1 : 0) - 1;  // inclusive count of group heads, minus one
        g[h[c]] = rank;
        if (flag) gfd[rank] = d[c];
      }
    }
  }
}

__global__ void __launch_bounds__(256) key_bits_kernel(const int32_t *gAll, int M, int bits, int wpc, uint32_t *out) {
  const int lane = threadIdx.x & 31;
  const int nWords = (M + 31) / 32;
  const int node = blockIdx.y;
  const int wd = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wd >= nWords) return;
  const int h = wd * 32 + lane;
  const int v = h < M ? gAll[(size_t)node * M + h] : 0;
  for (int b = 0; b < bits; ++b) {
    const uint32_t w = __ballot_sync(kFull, (v >> b) & 1);
    if (lane == 0) out[((size_t)node * bits + b) * wpc + wd] = w;
  }
}

// out[n][j] = table[n][idx[n][j]] (table shared by all n when tableShared): composition of permutations
__global__ void __launch_bounds__(256) compose_kernel(int32_t *out, const int32_t *table, int tableShared,
                                                      const int32_t *idx, int M) {
  const int n = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const int32_t *t = table + (tableShared ? 0 : (size_t)n * M);
  out[(size_t)n * M + j] = t[idx[(size_t)n * M + j]];
}

__global__ void iota_kernel(int32_t *out, int M) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < M) out[j] = j;
}

__global__ void __launch_bounds__(256) gather_records_kernel(const int4 *pool, const int *chunkTab, int poolChunks,
                                                             const long long *winOffset,
                                                             const unsigned long long *winCount, int4 *out) {
  const int w = blockIdx.y, c = blockIdx.x;
  const long long cnt = (long long)winCount[w];
  const long long first = (long long)c << kRecChunkLog;
  if (first >= cnt) return;
  const int id = chunkTab[(size_t)w * poolChunks + c];
  if (id < 0) return;
  const int n = (int)min((long long)kRecChunk, cnt - first);
  const int4 *srcp = pool + ((long long)id << kRecChunkLog);
  int4 *dst = out + winOffset[w] + first;
  for (int t = threadIdx.x; t < n; t += blockDim.x) dst[t] = srcp[t];
}


// Records of a wide-panel sweep: one segment per (column, CTA of the window's group), in callback order
// (column-major, CTAs in position order).  One thread per output record: binary search of its segment in the
// exclusive scan of the segment counts (segments differ in size by orders of magnitude and most are empty).
__global__ void __launch_bounds__(256) gather_segments_kernel(const int4 *pool, const int *chunkTab, int poolChunks,
                                                              const long long *segSrc, const long long *segDst,
                                                              long long nSeg, long long total, int G, int spw, int N, int W,
                                                              int4 *out) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (long long)gridDim.x * blockDim.x) {
    long long lo = 0, hi = nSeg - 1;  // last segment with segDst[seg] <= r (segDst[nSeg] = total > r)
    while (lo < hi) {
      const long long mid = (lo + hi + 1) >> 1;
      if (__ldg(segDst + mid) <= r) lo = mid; else hi = mid - 1;
    }
    const long long seg = lo;
    const int col = (int)(seg / G), g = (int)(seg % G);
    const int win = col >= N ? W - 1 : col / spw;
    const long long o = __ldg(segSrc + seg) + (r - __ldg(segDst + seg));
    const int id = __ldg(chunkTab + (size_t)(win * G + g) * poolChunks + (o >> kRecChunkLog));
    if (id >= 0) out[r] = pool[((long long)id << kRecChunkLog) + (o & (kRecChunk - 1))];
  }
}

struct IntToLL {
  __host__ __device__ long long operator()(int32_t v) const { return (long long)v; }
};

__global__ void permute_nomatch_kernel(const pbwtgpu_nomatch *in, const uint32_t *idx, long long n, pbwtgpu_nomatch *out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[idx[i]];
}
__global__ void iota_u32_kernel(uint32_t *out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t)i;
}

}  // namespace

int launch_make_groups(const int32_t *ord, int32_t *dv, const int32_t *gJ, const int32_t *gfdJ, int32_t *g, int32_t *gfd,
                       int M, int nodes, cudaStream_t stream) {
  if (nodes > 0) {
    if (M <= kGroupsFastChunks * 1024) make_groups_fast_kernel<<<nodes, 1024, 0, stream>>>(ord, dv, gJ, gfdJ, g, gfd, M);
    else make_groups_kernel<<<nodes, 1024, 0, stream>>>(ord, dv, gJ, gfdJ, g, gfd, M);
  }
  return (int)cudaGetLastError();
}

int launch_key_bits(const int32_t *g, int M, int nodes, int bits, int wpc, uint32_t *out, cudaStream_t stream) {
  if (nodes > 0) {
    dim3 grid(((M + 31) / 32 + 7) / 8, nodes);
    key_bits_kernel<<<grid, 256, 0, stream>>>(g, M, bits, wpc, out);
  }
  return (int)cudaGetLastError();
}

int launch_compose(int32_t *out, const int32_t *table, bool tableShared, const int32_t *idx, int M, int n,
                   cudaStream_t stream) {
  if (n > 0 && M > 0) {
    dim3 grid((M + 255) / 256, n);
    compose_kernel<<<grid, 256, 0, stream>>>(out, table, tableShared ? 1 : 0, idx, M);
  }
  return (int)cudaGetLastError();
}

int launch_iota(int32_t *out, int M, cudaStream_t stream) {
  if (M > 0) iota_kernel<<<(M + 255) / 256, 256, 0, stream>>>(out, M);
  return (int)cudaGetLastError();
}

int launch_gather_records(const pbwtgpu_match *pool, const int *chunkTab, int poolChunks, const long long *winOffset,
                          const unsigned long long *winCount, int nWindows, int maxChunksPerWindow, pbwtgpu_match *out,
                          cudaStream_t stream) {
  if (nWindows > 0 && maxChunksPerWindow > 0) {
    dim3 grid(maxChunksPerWindow, nWindows);
    gather_records_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const int4 *>(pool), chunkTab, poolChunks, winOffset,
                                                    winCount, reinterpret_cast<int4 *>(out));
  }
  return (int)cudaGetLastError();
}

size_t segment_scan_temp_bytes(long long nSeg) {
  size_t bytes = 0;
  cub::TransformInputIterator<long long, IntToLL, const int32_t *> it(nullptr, IntToLL());
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, (long long *)nullptr, (int)nSeg);
  return bytes;
}

int launch_segment_scan(const int32_t *segCnt, long long nSeg, long long *segDst, void *temp, size_t tempBytes,
                        cudaStream_t stream) {
  if (nSeg > INT_MAX) return (int)cudaErrorInvalidValue;  // callers refuse such shapes up front (PBWTGPU_ERR_UNSUPPORTED)
  cub::TransformInputIterator<long long, IntToLL, const int32_t *> it(segCnt, IntToLL());
  return (int)cub::DeviceScan::ExclusiveSum(temp, tempBytes, it, segDst, (int)nSeg, stream);
}

int launch_gather_segments(const pbwtgpu_match *pool, const int *chunkTab, int poolChunks, const long long *segSrc,
                           const long long *segDst, long long nSeg, long long total, int G, int spw, int N, int W,
                           pbwtgpu_match *out, cudaStream_t stream) {
  if (nSeg > 0 && total > 0) {
    const long long blocks = std::min<long long>((total + 255) / 256, 148LL * 64);
    gather_segments_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const int4 *>(pool), chunkTab, poolChunks,
                                                                 segSrc, segDst, nSeg, total, G, spw, N, W,
                                                                 reinterpret_cast<int4 *>(out));
  }
  return (int)cudaGetLastError();
}

int sort_nomatch(const unsigned long long *keys, const pbwtgpu_nomatch *events, long long n, pbwtgpu_nomatch *out,
                 cudaStream_t stream) {
  if (n <= 0) return 0;
  if (n > INT_MAX) return (int)cudaErrorInvalidValue;
  unsigned long long *keysOut = nullptr;
  uint32_t *idxIn = nullptr, *idxOut = nullptr;
  void *temp = nullptr;
  size_t tempBytes = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, keys, keysOut, idxIn, idxOut, (int)n, 0, 64, stream);
  if (e == cudaSuccess) e = cudaMalloc(&keysOut, sizeof(unsigned long long) * n);
  if (e == cudaSuccess) e = cudaMalloc(&idxIn, sizeof(uint32_t) * n);
  if (e == cudaSuccess) e = cudaMalloc(&idxOut, sizeof(uint32_t) * n);
  if (e == cudaSuccess) e = cudaMalloc(&temp, tempBytes);
  if (e == cudaSuccess) {
    iota_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(idxIn, n);
    e = cub::DeviceRadixSort::SortPairs(temp, tempBytes, keys, keysOut, idxIn, idxOut, (int)n, 0, 64, stream);
  }
  if (e == cudaSuccess) {
    permute_nomatch_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(events, idxOut, n, out);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cudaFree(keysOut); cudaFree(idxIn); cudaFree(idxOut); cudaFree(temp);
  return (int)e;
}

}  // namespace pbwtgpu
