// pack3.cuh -- device pack3 codec launchers (see pack3.cu).
#pragma once
#include "common.cuh"

namespace pbwtgpu {

// per-column byte counts (int32[N+1]) and their exclusive scan (int64[N+1], last = total bytes);
// scratch: p3_scan_scratch_bytes(N) bytes of device memory (NULL: a single-block scan is used)
size_t p3_scan_scratch_bytes(int N);
int p3_count_columns(const uint32_t *ycols, int wpc, int M, int N, int32_t *colBytes, int64_t *colBase, void *scratch,
                     size_t scratchBytes, cudaStream_t stream);
int p3_emit_columns(const uint32_t *ycols, int wpc, int M, int N, const int64_t *colBase, uint8_t *yz,
                    cudaStream_t stream);
// pack3 from the run lists the sweep kernel leaves (SweepArgs::p3Out): site k's run heads -- (position << 1 | allele), in
// no particular order -- are runs + win(k) * cap + siteOff[k], siteCnt[k] of them.  count: sorts every site's heads in
// place (sites with many heads: rebuilds the column in shared memory instead), byte counts to colBytes and their exclusive
// scan to colBase (as p3_count_columns); *overflow is set when the sweep dropped heads (the caller falls back to the
// encoder over sorted columns).  emit: the bytes.  M <= 32 * kP3MaxRunsPerSite.
constexpr int kP3MaxRunsPerSite = 1024;  // words of shared memory per site (a column of up to 32,768 haplotypes)
int p3_count_runs(uint32_t *runs, long long cap, const int32_t *siteOff, const int32_t *siteCnt, int M, int N, int spw, int first,
                  int second, int32_t *colBytes, int64_t *colBase, void *scratch, size_t scratchBytes, int *overflow, cudaStream_t stream);
int p3_emit_runs(const uint32_t *runs, long long cap, const int32_t *siteOff, const int32_t *siteCnt, int M, int N, int spw, int first,
                 int second, const int64_t *colBase, uint8_t *yz, cudaStream_t stream);
// decode: tileScratch needs p3_decode_num_tiles(nyz)+1 int64; tileScratch[nTiles] = symbols decoded
int64_t p3_decode_num_tiles(int64_t nyz);
int p3_decode_columns(const uint8_t *yz, int64_t nyz, int M, int N, int wpc, int64_t *tileScratch, uint32_t *ycols,
                      int32_t *zeroCounts, cudaStream_t stream);
int bytes_to_bits(const uint8_t *y, int M, int N, int wpc, uint32_t *ycols, cudaStream_t stream);
int bits_to_bytes(const uint32_t *ycols, int M, int N, int wpc, uint8_t *y, cudaStream_t stream);

}  // namespace pbwtgpu
