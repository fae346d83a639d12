// pack3.cu -- the reference's 3-level run-length codec on the device.
//
// Encode (pack3 / pack3Add / pack3arrayAdd, pbwtCore.c:240-277): the sweep kernels hand over the
// sorted columns y_k as bit columns; columns are independent, so one warp encodes one column.
// Runs are found with word-wide bit tricks (a run ends where a bit differs from its successor),
// per-lane byte counts are prefix-summed inside the warp, and column byte offsets come from an
// exclusive scan of the per-column byte counts (runs never span columns: pack3 is called per column).
//
// Decode (unpack3, pbwtCore.c:279-305): a 64-bit prefix sum of the run lengths coded by every byte
// gives each run's absolute symbol position; column boundaries fall where that position crosses a
// multiple of M.  Runs of 1s are OR-ed into a zeroed bit-column buffer.
#include "pack3.cuh"
#include "windows.cuh"

namespace pbwtgpu {
namespace {

// Visit the runs of one column that END inside the 32 words of the chunk [w0, w0+32) handled by
// this warp (lane = word).  carryEnd = last run-end position before the chunk (-1 at column start).
// f(runValue, runLength, ordinalWithinLane) is called per run, in order, by the owning lane.
template <typename F>
__device__ __forceinline__ void visit_runs(const uint32_t *col, int M, int nWords, int w0, int lane, int &carryEnd,
                                           F &&f) {
  const int wd = w0 + lane;
  const int base = wd * 32;
  uint32_t w = 0, vm = 0;
  if (wd < nWords) {
    w = col[wd];
    vm = (M - base >= 32) ? kFull : ((1u << (M - base)) - 1u);
  }
  uint32_t nb = __shfl_down_sync(kFull, w, 1) & 1u;  // first bit of the next word
  if (lane == 31) nb = (wd + 1 < nWords) ? (col[wd + 1] & 1u) : 0u;
  uint32_t ends = (w ^ ((w >> 1) | (nb << 31))) & vm;
  if (wd < nWords && M - 1 >= base && M - 1 < base + 32) ends |= 1u << ((M - 1) & 31);
  // last end position at or before each lane's word, exclusive over lanes
  int lastEnd = ends ? base + 31 - __clz(ends) : -1;
  int x = lastEnd;
#pragma unroll
  for (int dl = 1; dl < 32; dl <<= 1) {
    const int t = __shfl_up_sync(kFull, x, dl);
    if (lane >= dl) x = max(x, t);
  }
  int prevEnd = __shfl_up_sync(kFull, x, 1);
  if (lane == 0) prevEnd = -1;
  prevEnd = max(prevEnd, carryEnd);
  carryEnd = max(carryEnd, __shfl_sync(kFull, x, 31));
  int ord = 0;
  while (ends) {
    const int j = __ffs(ends) - 1;
    ends &= ends - 1;
    const int pos = base + j;
    f((int)((w >> j) & 1u), pos - prevEnd, ord++);
    prevEnd = pos;
  }
}

constexpr int kWarpsPerBlock = 8;

// pack3 bytes of one bit column (global or shared memory), by one warp
__device__ __forceinline__ int count_column(const uint32_t *col, int M, int lane) {
  const int nWords = (M + 31) / 32;
  int carryEnd = -1, bytes = 0;
  for (int w0 = 0; w0 < nWords; w0 += 32)
    visit_runs(col, M, nWords, w0, lane, carryEnd, [&](int, int n, int) { bytes += p3_run_bytes(n); });
  return __reduce_add_sync(kFull, bytes);
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) p3_count_kernel(const uint32_t *ycols, int wpc, int M, int N,
                                                                        int32_t *colBytes) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (k >= N) return;
  const int bytes = count_column(ycols + (size_t)k * wpc, M, lane);
  if (lane == 0) colBytes[k] = bytes;
}

// the bytes of one bit column (global or shared memory), by one warp
__device__ __forceinline__ void emit_column(const uint32_t *col, int M, uint8_t *out, int lane) {
  const int nWords = (M + 31) / 32;
  int carryEnd = -1;
  int chunkBase = 0;
  for (int w0 = 0; w0 < nWords; w0 += 32) {
    // two passes over the chunk: byte counts per lane, then emission at the prefix offsets
    int c0 = carryEnd, bytes = 0;
    visit_runs(col, M, nWords, w0, lane, c0, [&](int, int n, int) { bytes += p3_run_bytes(n); });
    int x = bytes;
#pragma unroll
    for (int dl = 1; dl < 32; dl <<= 1) {
      const int t = __shfl_up_sync(kFull, x, dl);
      if (lane >= dl) x += t;
    }
    const int total = __shfl_sync(kFull, x, 31);
    int off = chunkBase + x - bytes;
    visit_runs(col, M, nWords, w0, lane, carryEnd, [&](int v, int n, int) { off += p3_run_emit(out + off, v, n); });
    chunkBase += total;
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) p3_emit_kernel(const uint32_t *ycols, int wpc, int M, int N,
                                                                       const int64_t *colBase, uint8_t *yz) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (k >= N) return;
  emit_column(ycols + (size_t)k * wpc, M, yz + colBase[k], lane);
}

// Exclusive scan of int32 counts into int64 offsets, out[n] = total. Single block.
__global__ void __launch_bounds__(1024) scan_i32_to_i64_kernel(const int32_t *in, int64_t *out, int n) {
  __shared__ long long warpSums[32];
  __shared__ long long carryS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carryS = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + tid;
    const long long v = i < n ? in[i] : 0;
    long long x = v;
#pragma unroll
    for (int dl = 1; dl < 32; dl <<= 1) {
      const long long t = __shfl_up_sync(kFull, x, dl);
      if (lane >= dl) x += t;
    }
    if (lane == 31) warpSums[warp] = x;
    __syncthreads();
    long long wv = warpSums[lane];
    long long wx = wv;
#pragma unroll
    for (int dl = 1; dl < 32; dl <<= 1) {
      const long long t = __shfl_up_sync(kFull, wx, dl);
      if (lane >= dl) wx += t;
    }
    const long long warpOff = __shfl_sync(kFull, wx - wv, warp);
    const long long blockTotal = __shfl_sync(kFull, wx, 31);
    const long long carry = carryS;
    if (i < n) out[i] = carry + warpOff + x - v;
    __syncthreads();
    if (tid == 0) carryS = carry + blockTotal;
    __syncthreads();
  }
  if (tid == 0) out[n] = carryS;
}

// ---- encode from run lists (the sweep kernel's fused run detection, sweep.cuh: SweepArgs::p3Out) ----
// window of column k for the window layout of sweep.cuh:window_range
__device__ __forceinline__ int window_of(int k, int spw, int first, int second) {
  if (first <= 0) return k / spw;
  if (k < first) return 0;
  if (k < first + second) return 1;
  return 2 + (k - first - second) / spw;
}

constexpr int kRunWarps = 8;
constexpr int kP3SortedRunsPerSite = 256;  // more heads than this in a site: rebuild its column instead of ranking them

// The sorted column of a site rebuilt in shared memory from its run heads, for sites with too many heads to rank: the
// heads' positions are a bitmap (that IS their sorted order), and since the allele flips at every head but the first,
// y[p] = y[0] ^ parity(heads in [1, p]): a prefix-xor inside each word plus a carry across the words.
__device__ __forceinline__ void column_from_heads(const uint32_t *e, int cnt, int M, uint32_t *col, int lane) {
  const int nWords = (M + 31) / 32;
  for (int w = lane; w < nWords; w += 32) col[w] = 0u;
  __syncwarp();
  uint32_t y0 = 0u;
  for (int i = lane; i < cnt; i += 32) {
    const uint32_t h = e[i];
    const int pos = (int)(h >> 1);
    if (pos == 0) y0 = h & 1u;
    else atomicOr(&col[pos >> 5], 1u << (pos & 31));
  }
  y0 = __reduce_or_sync(kFull, y0);
  __syncwarp();
  const int per = (nWords + 31) / 32;  // consecutive words per lane
  const int wLo = min(lane * per, nWords), wHi = min(wLo + per, nWords);
  uint32_t par = 0u;
  for (int w = wLo; w < wHi; ++w) par ^= __popc(col[w]) & 1u;
  uint32_t inc = par;  // parity of the heads in the words of the lanes up to and including this one
#pragma unroll
  for (int dl = 1; dl < 32; dl <<= 1) {
    const uint32_t t = __shfl_up_sync(kFull, inc, dl);
    if (lane >= dl) inc ^= t;
  }
  uint32_t carry = (inc ^ par) ^ y0;  // allele just before this lane's first word
  for (int w = wLo; w < wHi; ++w) {
    uint32_t x = col[w];
    x ^= x << 1; x ^= x << 2; x ^= x << 4; x ^= x << 8; x ^= x << 16;  // bit j = parity of the heads at bits 0 .. j
    if (carry) x = ~x;
    col[w] = x;
    carry = x >> 31;
  }
  __syncwarp();
}

// One warp per site.  Its run heads arrive in no particular order.  Few heads (the rule once the PBWT order has formed): rank
// them by position -- each lane counts the heads below its own --, leave them sorted in place, and count the bytes of
// the runs between them.  Many heads (the first sites of a panel, dense columns): rebuild the column and count as the
// encoder over columns does.
__global__ void __launch_bounds__(kRunWarps * 32) p3_count_runs_kernel(uint32_t *runs, long long cap, const int32_t *siteOff,
                                                                        const int32_t *siteCnt, int M, int N, int spw, int first,
                                                                        int second, int32_t *colBytes, int *overflow) {
  __shared__ uint32_t buf[kRunWarps][kP3MaxRunsPerSite];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.x * kRunWarps + warp;
  if (k >= N) return;
  const int cnt = siteCnt[k];
  if ((long long)siteOff[k] + cnt > cap) {  // the sweep dropped heads of this window
    if (lane == 0) { *overflow = 1; colBytes[k] = 0; }
    return;
  }
  uint32_t *e = runs + (size_t)window_of(k, spw, first, second) * cap + siteOff[k];
  uint32_t *so = buf[warp];
  if (cnt > kP3SortedRunsPerSite) {
    column_from_heads(e, cnt, M, so, lane);
    const int bytes = count_column(so, M, lane);
    if (lane == 0) colBytes[k] = bytes;
    return;
  }
  uint32_t *raw = so + kP3SortedRunsPerSite;  // the list as it arrived, staged once (ranking reads it cnt times)
  for (int i = lane; i < cnt; i += 32) raw[i] = e[i];
  __syncwarp();
  for (int i = lane; i < cnt; i += 32) {
    const uint32_t mine = raw[i];
    int rank = 0;
    for (int j = 0; j < cnt; ++j) rank += raw[j] < mine;  // (positions are distinct; raw[j] is a broadcast)
    so[rank] = mine;
  }
  __syncwarp();
  int bytes = 0;
  for (int i = lane; i < cnt; i += 32) {
    const uint32_t v = so[i];
    e[i] = v;
    const int next = i + 1 < cnt ? (int)(so[i + 1] >> 1) : M;
    bytes += p3_run_bytes(next - (int)(v >> 1));
  }
  bytes = __reduce_add_sync(kFull, bytes);
  if (lane == 0) colBytes[k] = bytes;
}

__global__ void __launch_bounds__(kRunWarps * 32) p3_emit_runs_kernel(const uint32_t *runs, long long cap, const int32_t *siteOff,
                                                                       const int32_t *siteCnt, int M, int N, int spw, int first,
                                                                       int second, const int64_t *colBase, uint8_t *yz) {
  __shared__ uint32_t buf[kRunWarps][kP3MaxRunsPerSite];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.x * kRunWarps + warp;
  if (k >= N) return;
  const int cnt = siteCnt[k];
  const uint32_t *e = runs + (size_t)window_of(k, spw, first, second) * cap + siteOff[k];
  uint8_t *out = yz + colBase[k];
  if (cnt > kP3SortedRunsPerSite) {
    column_from_heads(e, cnt, M, buf[warp], lane);
    emit_column(buf[warp], M, out, lane);
    return;
  }
  int chunkBase = 0;
  for (int i0 = 0; i0 < cnt; i0 += 32) {  // 32 runs per step (sorted by the count pass): byte counts, prefix offsets, bytes
    const int i = i0 + lane;
    int v = 0, len = 0;
    if (i < cnt) {
      const uint32_t h = e[i];
      v = (int)(h & 1u);
      len = (i + 1 < cnt ? (int)(e[i + 1] >> 1) : M) - (int)(h >> 1);
    }
    const int nb = p3_run_bytes(len);
    int x = nb;
#pragma unroll
    for (int dl = 1; dl < 32; dl <<= 1) {
      const int t = __shfl_up_sync(kFull, x, dl);
      if (lane >= dl) x += t;
    }
    if (len) p3_run_emit(out + chunkBase + x - nb, v, len);
    chunkBase += __shfl_sync(kFull, x, 31);
  }
}

// ---- decode ---------------------------------------------------------------------------------
constexpr int kDecThreads = 256;
constexpr int kDecPerThread = 8;
constexpr int kDecTile = kDecThreads * kDecPerThread;

__global__ void __launch_bounds__(kDecThreads) p3_tile_sums_kernel(const uint8_t *yz, int64_t nyz, int64_t *tileSums) {
  __shared__ long long ws[kDecThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kDecTile;
  long long s = 0;
  for (int j = 0; j < kDecPerThread; ++j) {
    const int64_t b = base + (int64_t)j * kDecThreads + threadIdx.x;
    if (b < nyz) s += p3_byte_len(yz[b]);
  }
  s = __reduce_add_sync(kFull, (int)s);  // a tile holds < 2048 * 63488 symbols: fits in int32
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int i = 0; i < kDecThreads / 32; ++i) t += ws[i];
    tileSums[blockIdx.x] = t;
  }
}

// in-place exclusive scan of int64 values, out[n] = total. Single block.
__global__ void __launch_bounds__(1024) scan_i64_kernel(int64_t *data, int n) {
  __shared__ long long warpSums[32];
  __shared__ long long carryS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carryS = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + tid;
    const long long v = i < n ? data[i] : 0;
    long long x = v;
#pragma unroll
    for (int dl = 1; dl < 32; dl <<= 1) {
      const long long t = __shfl_up_sync(kFull, x, dl);
      if (lane >= dl) x += t;
    }
    if (lane == 31) warpSums[warp] = x;
    __syncthreads();
    long long wv = warpSums[lane];
    long long wx = wv;
#pragma unroll
    for (int dl = 1; dl < 32; dl <<= 1) {
      const long long t = __shfl_up_sync(kFull, wx, dl);
      if (lane >= dl) wx += t;
    }
    const long long warpOff = __shfl_sync(kFull, wx - wv, warp);
    const long long blockTotal = __shfl_sync(kFull, wx, 31);
    const long long carry = carryS;
    if (i < n) data[i] = carry + warpOff + x - v;
    __syncthreads();
    if (tid == 0) carryS = carry + blockTotal;
    __syncthreads();
  }
  if (tid == 0) data[n] = carryS;
}

__device__ __forceinline__ void or_bits(uint32_t *col, int lo, int hi) {  // set bits [lo, hi)
  int w = lo >> 5;
  const int wEnd = (hi - 1) >> 5;
  uint32_t m = kFull << (lo & 31);
  for (; w < wEnd; ++w) {
    atomicOr(&col[w], m);
    m = kFull;
  }
  const int r = hi & 31;
  if (r) m &= (1u << r) - 1u;
  atomicOr(&col[w], m);
}

__global__ void __launch_bounds__(kDecThreads) p3_expand_kernel(const uint8_t *yz, int64_t nyz, const int64_t *tileOff,
                                                                 int M, int N, int wpc, uint32_t *ycols,
                                                                 int32_t *zeroCounts) {
  __shared__ long long ws[kDecThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // blocked assignment: thread t owns bytes [t*P, (t+1)*P) of the tile
  const int64_t b0 = (int64_t)blockIdx.x * kDecTile + (int64_t)tid * kDecPerThread;
  int len[kDecPerThread];
  uint8_t val[kDecPerThread];
  long long s = 0;
#pragma unroll
  for (int j = 0; j < kDecPerThread; ++j) {
    const int64_t b = b0 + j;
    const uint8_t byte = b < nyz ? yz[b] : 0;
    len[j] = p3_byte_len(byte);
    val[j] = byte >> 7;
    s += len[j];
  }
  long long x = s;
#pragma unroll
  for (int dl = 1; dl < 32; dl <<= 1) {
    const long long t = __shfl_up_sync(kFull, x, dl);
    if (lane >= dl) x += t;
  }
  if (lane == 31) ws[warp] = x;
  __syncthreads();
  long long pos = tileOff[blockIdx.x] + x - s;
  for (int i = 0; i < warp; ++i) pos += ws[i];
#pragma unroll
  for (int j = 0; j < kDecPerThread; ++j) {
    if (len[j] && val[j]) {
      const long long k = pos / M;
      const int o = (int)(pos - k * M);
      if (k < N && o + len[j] <= M) or_bits(ycols + (size_t)k * wpc, o, o + len[j]);
    } else if (len[j] && zeroCounts) {
      const long long k = pos / M;
      if (k < N) atomicAdd(&zeroCounts[k], len[j]);
    }
    pos += len[j];
  }
}

// set the pad bits [M, 32*wpc) of every column to one: the sweep kernels treat positions >= M as
// permanent 1-alleles (sweep_smem.cu)
__global__ void p3_pad_ones_kernel(uint32_t *ycols, int M, int N, int wpc) {
  const int firstWord = M >> 5;
  const int padWords = wpc - firstWord;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)N * padWords) return;
  const int k = (int)(g / padWords), w = firstWord + (int)(g % padWords);
  const uint32_t m = (w == firstWord && (M & 31)) ? (kFull << (M & 31)) : kFull;
  ycols[(size_t)k * wpc + w] |= m;
}

// bytes-per-allele columns -> bit columns (for the standalone pack3 entry point)
__global__ void bytes_to_bits_kernel(const uint8_t *y, int M, int N, int wpc, uint32_t *ycols) {
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nWords = (M + 31) / 32;
  if (gw >= (long long)N * nWords) return;
  const int k = (int)(gw / nWords), wd = (int)(gw % nWords);
  const int i = wd * 32 + lane;
  const uint32_t bit = i < M ? (y[(size_t)k * M + i] & 1u) : 0u;
  const uint32_t w = __ballot_sync(kFull, bit);
  if (lane == 0) ycols[(size_t)k * wpc + wd] = w;
}

__global__ void bits_to_bytes_kernel(const uint32_t *ycols, int M, int N, int wpc, uint8_t *y) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)N * M) return;
  const int k = (int)(g / M), i = (int)(g % M);
  y[g] = (ycols[(size_t)k * wpc + (i >> 5)] >> (i & 31)) & 1u;
}

}  // namespace

static int scan_col_bytes(int32_t *colBytes, int64_t *colBase, int N, void *scratch, size_t scratchBytes, cudaStream_t stream);

int p3_count_columns(const uint32_t *ycols, int wpc, int M, int N, int32_t *colBytes, int64_t *colBase, void *scratch,
                     size_t scratchBytes, cudaStream_t stream) {
  if (N > 0) {
    p3_count_kernel<<<(N + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, stream>>>(ycols, wpc, M, N,
                                                                                                   colBytes);
  }
  return scan_col_bytes(colBytes, colBase, N, scratch, scratchBytes, stream);
}

// exclusive 64-bit scan of the per-column byte counts, colBase[N] = total
static int scan_col_bytes(int32_t *colBytes, int64_t *colBase, int N, void *scratch, size_t scratchBytes, cudaStream_t stream) {
  if (N < 4096 || !scratch) {  // short streams: one block is quicker than a library scan
    scan_i32_to_i64_kernel<<<1, 1024, 0, stream>>>(colBytes, colBase, N);
    return (int)cudaGetLastError();
  }
  // N counts plus a trailing zero (CUB)
  cudaError_t err = cudaMemsetAsync(colBytes + N, 0, sizeof(int32_t), stream);
  if (err != cudaSuccess) return (int)err;
  return launch_segment_scan(colBytes, (long long)N + 1, reinterpret_cast<long long *>(colBase), scratch, scratchBytes, stream);
}

size_t p3_scan_scratch_bytes(int N) { return segment_scan_temp_bytes((long long)N + 1); }

int p3_count_runs(uint32_t *runs, long long cap, const int32_t *siteOff, const int32_t *siteCnt, int M, int N, int spw, int first,
                  int second, int32_t *colBytes, int64_t *colBase, void *scratch, size_t scratchBytes, int *overflow, cudaStream_t stream) {
  if (N > 0) {
    p3_count_runs_kernel<<<(N + kRunWarps - 1) / kRunWarps, kRunWarps * 32, 0, stream>>>(runs, cap, siteOff, siteCnt, M, N, spw, first, second,
                                                                                       colBytes, overflow);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return (int)err;
  }
  return scan_col_bytes(colBytes, colBase, N, scratch, scratchBytes, stream);
}

int p3_emit_runs(const uint32_t *runs, long long cap, const int32_t *siteOff, const int32_t *siteCnt, int M, int N, int spw, int first,
                 int second, const int64_t *colBase, uint8_t *yz, cudaStream_t stream) {
  if (N > 0)
    p3_emit_runs_kernel<<<(N + kRunWarps - 1) / kRunWarps, kRunWarps * 32, 0, stream>>>(runs, cap, siteOff, siteCnt, M, N, spw, first, second,
                                                                                      colBase, yz);
  return (int)cudaGetLastError();
}

int p3_emit_columns(const uint32_t *ycols, int wpc, int M, int N, const int64_t *colBase, uint8_t *yz,
                    cudaStream_t stream) {
  if (N > 0)
    p3_emit_kernel<<<(N + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, stream>>>(ycols, wpc, M, N,
                                                                                                  colBase, yz);
  return (int)cudaGetLastError();
}

int64_t p3_decode_num_tiles(int64_t nyz) { return (nyz + kDecTile - 1) / kDecTile; }

int p3_decode_columns(const uint8_t *yz, int64_t nyz, int M, int N, int wpc, int64_t *tileScratch, uint32_t *ycols,
                      int32_t *zeroCounts, cudaStream_t stream) {
  const int64_t nTiles = p3_decode_num_tiles(nyz);
  cudaMemsetAsync(ycols, 0, (size_t)N * wpc * 4, stream);
  if (zeroCounts) cudaMemsetAsync(zeroCounts, 0, (size_t)N * 4, stream);
  if (nTiles == 0) {
    cudaMemsetAsync(tileScratch, 0, sizeof(int64_t), stream);
    return (int)cudaGetLastError();
  }
  p3_tile_sums_kernel<<<(unsigned)nTiles, kDecThreads, 0, stream>>>(yz, nyz, tileScratch);
  scan_i64_kernel<<<1, 1024, 0, stream>>>(tileScratch, (int)nTiles);
  p3_expand_kernel<<<(unsigned)nTiles, kDecThreads, 0, stream>>>(yz, nyz, tileScratch, M, N, wpc, ycols, zeroCounts);
  const long long padThreads = (long long)N * (wpc - (M >> 5));
  if (padThreads > 0) p3_pad_ones_kernel<<<(unsigned)((padThreads + 255) / 256), 256, 0, stream>>>(ycols, M, N, wpc);
  return (int)cudaGetLastError();
}

int bytes_to_bits(const uint8_t *y, int M, int N, int wpc, uint32_t *ycols, cudaStream_t stream) {
  const long long threads = (long long)N * ((M + 31) / 32) * 32;
  if (threads == 0) return 0;
  cudaMemsetAsync(ycols, 0, (size_t)N * wpc * 4, stream);
  bytes_to_bits_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(y, M, N, wpc, ycols);
  return (int)cudaGetLastError();
}

int bits_to_bytes(const uint32_t *ycols, int M, int N, int wpc, uint8_t *y, cudaStream_t stream) {
  const long long n = (long long)N * M;
  if (n == 0) return 0;
  bits_to_bytes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(ycols, M, N, wpc, y);
  return (int)cudaGetLastError();
}

}  // namespace pbwtgpu
