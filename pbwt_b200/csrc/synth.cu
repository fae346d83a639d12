// synth.cu -- counter-based synthetic "recursive-copy mosaic" panel generator (kernel K8).
// x_k[h] = XOR of flip(k, v) over the ancestor path h -> parent_k(h) -> ... -> 0, a pure function of
// (seed, k, h); bit-identical to oracle/pbwt_oracle.c:orc_generate (which memoises the same recursion).
#include "synth.cuh"

namespace pbwtgpu {
namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t hash3(uint64_t seed, uint64_t a, uint64_t b) {
  return mix64(seed + a * 0x9E3779B97F4A7C15ull + b * 0xD1B54A32D192ED03ull);
}

__global__ void __launch_bounds__(256) synth_kernel(uint32_t *xbits, int wpc, int M, int N, int k0, uint64_t seed,
                                                    uint32_t L, uint32_t flipThresh) {
  const int lane = threadIdx.x & 31;
  const int nWords = (M + 31) / 32;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= (long long)N * nWords) return;
  const int kk = (int)(gw / nWords), wd = (int)(gw % nWords);
  const uint32_t k = (uint32_t)(k0 + kk);
  uint32_t v = (uint32_t)(wd * 32 + lane);
  uint32_t acc = 0;
  if (v < (uint32_t)M) {
    while (v >= 1u) {
      acc ^= (uint32_t)((uint32_t)(hash3(seed ^ 0x5A5A5A5A5A5Aull, k, v) >> 32) < flipThresh);
      const uint32_t phase = (uint32_t)(hash3(seed ^ 0xA5A5A5A5ull, v, 0x1234567ull) >> 32) % L;
      const uint64_t r = hash3(seed, v, (k + phase) / L) >> 32;
      v = (uint32_t)((r * v) >> 32);
    }
  }
  const uint32_t w = __ballot_sync(kFull, acc & 1u);
  if (lane == 0) xbits[(size_t)kk * wpc + wd] = w;
}

}  // namespace

int launch_synth(uint32_t *xbits, int wpc, int M, int N, int k0, uint64_t seed, uint32_t L, uint32_t flipThresh,
                 cudaStream_t stream) {
  const long long threads = (long long)N * ((M + 31) / 32) * 32;
  if (threads == 0) return 0;
  cudaMemsetAsync(xbits, 0, (size_t)N * wpc * 4, stream);
  synth_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(xbits, wpc, M, N, k0, seed, L, flipThresh);
  return (int)cudaGetLastError();
}

}  // namespace pbwtgpu
