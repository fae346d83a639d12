// query_chain.cuh -- true query states at the window boundaries of a window-parallel query sweep (query_chain.cu).
#pragma once
#include "common.cuh"

namespace pbwtgpu {

struct QueryChainArgs {
  int Q, M, W;
  int steps;                 // boundaries to cross: W - 1 (one GPU) or W (a shard hands its end state on: fOut / eOut)
  int spw;                   // sites per window (equal windows)
  int firstLabel;            // label of the first site of window 0
  const int32_t *dv;         // [W + 1][M + 1] true divergence entering window w (slot w) ... leaving the last one
  const int32_t *nodeOrd;    // [steps][M] window w's own permutation: position entering w of the element at position i leaving it
  const int32_t *fClamped, *eClamped;  // [steps][Q] clamped end states of the windows
  const int32_t *fIn, *eIn;  // [Q] true state entering window 0, or NULL = fresh (f = 0, e = firstLabel)
  int32_t *fTrue, *eTrue;    // [W][Q] out: true state entering every window
  int32_t *fOut, *eOut;      // [Q] out: true state leaving the last window (steps == W only)
};
int launch_query_chain(const QueryChainArgs &c, cudaStream_t stream);

}  // namespace pbwtgpu
