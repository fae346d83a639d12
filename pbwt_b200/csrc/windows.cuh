// windows.cuh -- exact window-start states for the window-parallel sweep (see windows.cu).
#pragma once
#include "common.cuh"
#include "sweep.cuh"

namespace pbwtgpu {

constexpr int kPseudoLabelBase = 1 << 29;  // site labels of the combine's pseudo-sites (above any real site)

// Tie groups of `nodes` prefix nodes.  ord/dv: [nodes][M] / [nodes][M+1] (order and divergence of each node, the
// divergence possibly still holding pseudo-site labels >= kPseudoLabelBase from a combine sweep, which are
// replaced by the later window's boundary divergence gfdJ[gJ[hap]]).  Writes dv in place, g [nodes][M] (group
// of each haplotype) and gfd [nodes][M] (divergence at the head of each group).
int launch_make_groups(const int32_t *ord, int32_t *dv, const int32_t *gJ, const int32_t *gfdJ, int32_t *g, int32_t *gfd,
                       int M, int nodes, cudaStream_t stream);
// Pseudo-site columns of a node: column b, bit h = bit b of g[h].  out: [nodes][bits][wpc]
int launch_key_bits(const int32_t *g, int M, int nodes, int bits, int wpc, uint32_t *out, cudaStream_t stream);
// out[n][j] = table[n or 0][idx[n][j]]: composition of permutations (position-space window combine, start order)
int launch_compose(int32_t *out, const int32_t *table, bool tableShared, const int32_t *idx, int M, int n, cudaStream_t stream);
int launch_iota(int32_t *out, int M, cudaStream_t stream);
// Gather the chunked match records of all windows into one contiguous array in window order.
int launch_gather_records(const pbwtgpu_match *pool, const int *chunkTab, int poolChunks, const long long *winOffset,
                          const unsigned long long *winCount, int nWindows, int maxChunksPerWindow, pbwtgpu_match *out,
                          cudaStream_t stream);

// Wide-panel sweeps write one record segment per (column, CTA): exclusive 64-bit scan of the segment counts
// (nSeg entries, the last one a zero so that segDst[nSeg-1] is the total) and the gather into callback order.
size_t segment_scan_temp_bytes(long long nSeg);
int launch_segment_scan(const int32_t *segCnt, long long nSeg, long long *segDst, void *temp, size_t tempBytes,
                        cudaStream_t stream);
// segDst: exclusive scan with nSeg + 1 entries (segDst[nSeg] = total)
int launch_gather_segments(const pbwtgpu_match *pool, const int *chunkTab, int poolChunks, const long long *segSrc,
                           const long long *segDst, long long nSeg, long long total, int G, int spw, int N, int W,
                           pbwtgpu_match *out, cudaStream_t stream);
// "no match" events of a wide query sweep arrive in arbitrary order: sort them by key (site, query-PBWT position)
int sort_nomatch(const unsigned long long *keys, const pbwtgpu_nomatch *events, long long n, pbwtgpu_nomatch *out,
                 cudaStream_t stream);

// The combine of adjacent nodes as a sort + range-maximum queries (combine_sort.cu) instead of a sweep over pseudo-sites:
// node n of cnt: I = (ordI, dvI)[n] (one node for all when sharedI), J's tie groups (gJ, gfdJ)[n] -> (ordOut, dvOut, gOut,
// gfdOut)[n].  ncopy further CTAs copy nodes that are already final from one buffer set to the other.
struct CombineSortArgs {
  int M, cnt, sharedI;
  const int32_t *ordI, *dvI, *gJ, *gfdJ;
  int32_t *ordOut, *dvOut, *gOut, *gfdOut;
  int ncopy;
  const int32_t *ordCopySrc, *dvCopySrc, *gCopySrc, *gfdCopySrc;
  int32_t *ordCopyDst, *dvCopyDst, *gCopyDst, *gfdCopyDst;
};
int combine_sort_max_m();
int launch_combine_sort(const CombineSortArgs &c, cudaStream_t stream);

// Pass 1 as a sort (window_sort.cu): nodes (order, divergence) of windows [w*spw, (w+1)*spw) of the panel columns,
// w < nodes, each swept on its own from aStart ([M], NULL = identity).  scratch: window_sort_scratch_words words.
// ordOut [nodes][M], dvOut [nodes][M+1]; labels are siteOffset + column + 1.
int window_sort_max_m();
size_t window_sort_scratch_words(int M, int spw, int nodes);
// `first` > 0: windows 0 and 1 have `first` and `second` sites (window_range, sweep.cuh).
int launch_window_sort(const uint32_t *panel, int wpc, int M, int N, int spw, int first, int second, int siteOffset,
                       const int32_t *aStart, int nodes, uint32_t *scratch, int32_t *ordOut, int32_t *dvOut, cudaStream_t stream);

// The same for wide panels (window_sort_wide.cu): one window = one cluster of 16 CTAs, ids in distributed shared memory.
bool window_sort_wide_supported(int M, int spw);
size_t window_sort_wide_scratch_words(int M, int spw, int nodes);
int launch_window_sort_wide(const uint32_t *panel, int wpc, int M, int N, int spw, int siteOffset, const int32_t *aStart, int nodes,
                            uint32_t *scratch, int32_t *ordOut, int32_t *dvOut, cudaStream_t stream);

// Panels too wide for the cluster (window_sort_wide.cu, global-memory variant): W windows, each sorted by a group of
// co-resident CTAs (cooperative launch, W * G <= maxCtas), elements in an L2-resident double buffer.
bool window_sort_global_supported(int M, int W, int maxCtas);
size_t window_sort_global_scratch_words(int M, int spw, int W, int maxCtas);
int launch_window_sort_global(const uint32_t *panel, int wpc, int M, int N, int spw, int siteOffset, const int32_t *aStart, int W,
                              int maxCtas, uint32_t *scratch, int32_t *ordOut, int32_t *dvOut, cudaStream_t stream);

}  // namespace pbwtgpu
