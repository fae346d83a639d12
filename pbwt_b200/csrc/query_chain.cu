// query_chain.cu -- the exact query state (f, e) of matchSequencesSweep (pbwtMatch.c:363-443) entering every site
// window, so that the windows of a query sweep can run concurrently (and on different GPUs).
//
// Per query the reference carries f (top of the block of panel haplotypes with the longest match ending here, in
// panel order) and e (where that match starts).  A window [K0, K1) swept on its own from the fresh state -- with
// every divergence tier that lies before K0 counted as K0, because the window knows nothing earlier -- ends in the
// CLAMPED state (f', e'): the top of the haplotypes that match the query on [e', K1) with e' = max(e, K0).  So
//   * e' > K0: the longest match starts inside the window; the clamped state IS the true one.
//   * e' = K0: the block B' = [f', h') at K1 holds the haplotypes that match the query across the whole window;
//     which of them match longest was decided at K0.  B' is ordered by the haplotypes' positions at K0 (the sort is
//     stable), P(i) = position at K0 of the haplotype at position i at K1 (the window's own permutation, position
//     space).  With the true state (f, e) at K0 and its block B = [f, h):
//       - some member of B' lies in B: the first such member is the new top, e stays;
//       - otherwise the nearest members above (j1) and below (j2) of B compete exactly like the reference's reseed
//         (pbwtMatch.c:389-411): Dm = max d_K0 (j1, f], Dp = max d_K0 [h, j2], the smaller one is the new e, up wins
//         ties and the new top is the first member of B' inside that tier.
// The chain is sequential over windows and independent per query: one warp per query, W - 1 cheap steps.
// tests/test_query_window_model.py proves the construction against the sequential sweep on the CPU.
#include "query_chain.cuh"

namespace pbwtgpu {
namespace {

constexpr int kNone = 0x7fffffff;

// divergence arrays hold interior values; positions 0 and M are sentinels above every site label
__device__ __forceinline__ int d_at(const int32_t *d, int i, int M) { return (i <= 0 || i >= M) ? kNone : __ldg(d + i); }

// first index >= from whose divergence exceeds t, or M
__device__ __forceinline__ int chain_block_end(const int32_t *d, int from, int t, int M, int lane) {
  for (int base = from; base < M; base += 32) {
    const int idx = base + lane;
    const unsigned m = __ballot_sync(kFull, idx >= M || d_at(d, idx, M) > t);
    if (m) return min(base + __ffs(m) - 1, M);
  }
  return M;
}
// nearest index <= from whose divergence exceeds t, or 0
__device__ __forceinline__ int chain_block_start(const int32_t *d, int from, int t, int M, int lane) {
  for (int top = from; top > 0; top -= 32) {
    const int idx = top - lane;
    const unsigned m = __ballot_sync(kFull, idx <= 0 || d_at(d, idx, M) > t);
    if (m) return max(top - (__ffs(m) - 1), 0);
  }
  return 0;
}
// max of d over [lo, hi], 1 <= lo <= hi < M
__device__ __forceinline__ int chain_range_max(const int32_t *d, int lo, int hi, int lane) {
  int m = 0;
  for (int i = lo + lane; i <= hi; i += 32) m = max(m, __ldg(d + i));
  return __reduce_max_sync(kFull, m);
}
// first index in [lo, hi) with P[index] >= t (P increasing there), or hi
__device__ __forceinline__ int chain_lower_bound(const int32_t *P, int lo, int hi, int t) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(P + mid) < t) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) query_chain_kernel(const QueryChainArgs c) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= c.Q) return;
  const int M = c.M;
  int f0 = c.fIn ? c.fIn[q] : 0;
  int e0 = c.eIn ? c.eIn[q] : c.firstLabel;  // pbwtMatch.c:368-369
  if (lane == 0 && c.W > 0) { c.fTrue[q] = f0; c.eTrue[q] = e0; }
  for (int w = 0; w < c.steps; ++w) {
    const int K0 = c.firstLabel + w * c.spw;
    const int fp = c.fClamped[(size_t)w * c.Q + q], ep = c.eClamped[(size_t)w * c.Q + q];
    if (ep > K0) {
      f0 = fp;
      e0 = ep;
    } else {
      const int32_t *d0 = c.dv + (size_t)w * (M + 1), *d1 = c.dv + (size_t)(w + 1) * (M + 1);
      const int32_t *P = c.nodeOrd + (size_t)w * M;
      const int hp = chain_block_end(d1, fp + 1, K0, M, lane);  // B' = [fp, hp) at K1
      const int h0 = chain_block_end(d0, f0 + 1, e0, M, lane);  // B  = [f0, h0) at K0
      const int i0 = chain_lower_bound(P, fp, hp, f0);
      if (i0 < hp && __ldg(P + i0) < h0) {
        f0 = i0;  // a member of B' inside B: e stays
      } else {
        const int j1 = i0 > fp ? __ldg(P + i0 - 1) : -1;
        const int j2 = i0 < hp ? __ldg(P + i0) : M;
        const int Dm = j1 >= 0 ? chain_range_max(d0, j1 + 1, f0, lane) : kNone;
        const int Dp = j2 < M ? chain_range_max(d0, h0, j2, lane) : kNone;
        if (Dm <= Dp) {
          const int t = chain_block_start(d0, j1, Dm, M, lane);
          f0 = chain_lower_bound(P, fp, hp, t);
          e0 = Dm;
        } else {
          f0 = i0;
          e0 = Dp;
        }
      }
    }
    if (lane == 0) {
      if (w + 1 < c.W) { c.fTrue[(size_t)(w + 1) * c.Q + q] = f0; c.eTrue[(size_t)(w + 1) * c.Q + q] = e0; }
      else { c.fOut[q] = f0; c.eOut[q] = e0; }
    }
  }
}

}  // namespace

int launch_query_chain(const QueryChainArgs &c, cudaStream_t stream) {
  if (c.Q < 1 || c.W < 1) return (int)cudaErrorInvalidValue;
  if (c.steps > c.W || (c.steps == c.W && !(c.fOut && c.eOut))) return (int)cudaErrorInvalidValue;
  const int warps = 8;
  query_chain_kernel<<<(c.Q + warps - 1) / warps, warps * 32, 0, stream>>>(c);
  return (int)cudaGetLastError();
}

}  // namespace pbwtgpu
