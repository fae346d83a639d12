#pragma once
#include "common.cuh"
namespace pbwtgpu {
// fills N columns (sites k0 .. k0+N-1) of a [N][wpc] bit-column buffer
int launch_synth(uint32_t *xbits, int wpc, int M, int N, int k0, uint64_t seed, uint32_t L, uint32_t flipThresh,
                 cudaStream_t stream);
}
