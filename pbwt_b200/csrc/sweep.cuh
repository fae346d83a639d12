// sweep.cuh -- launcher interface of the per-site forward sweep kernels.
#pragma once
#include "common.cuh"

namespace pbwtgpu {

// One launch sweeps W windows of consecutive sites concurrently (one CTA per window). Window w
// covers sites [w*sitesPerWindow, min(N,(w+1)*sitesPerWindow)) and starts from the exact state
// (a,d) entering its first site.
struct SweepArgs {
  const uint32_t *src;  // gather mode: panel bit columns x_k; sorted mode: sorted columns y_k
  int wpc;              // words per column
  int M, N;
  int sitesPerWindow;
  int nWindows;
  const int32_t *aStart;  // [W][M]; NULL = identity (only valid for W == 1)
  const int32_t *dStart;  // [W][M+1]; NULL = fresh cursor: zeros, ends k0+1 (pbwtCore.c:413-414)
  int32_t *aEnd;          // [W][M] state after the window's last site, or NULL
  int32_t *dEnd;          // [W][M+1] or NULL
  uint32_t *yOut;         // [N][wpc] sorted columns y_k (input to the pack3 encoder), or NULL
  int32_t *cOut;          // [N] zero counts c_k, or NULL
  pbwtgpu_match *rec;     // match records; window w writes rec + w*recCap
  long long recCap;       // per-window capacity (records beyond it are counted, not stored)
  unsigned long long *recCount;  // [W] records produced by each window
  int reportAtN;          // the window that ends at N also reports the k == N blocks
};

enum SweepMode { kGather = 0, kSorted = 1 };

// largest M the shared-memory-resident kernel handles with / without the divergence array
int sweep_smem_max_m(bool with_d);
// returns cudaError_t as int
int launch_sweep_smem(const SweepArgs &args, SweepMode mode, bool with_d, bool with_match, cudaStream_t stream);

}  // namespace pbwtgpu
