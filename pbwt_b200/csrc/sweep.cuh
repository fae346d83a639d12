// sweep.cuh -- launcher interface of the per-site forward sweep kernels.
#pragma once
#include "common.cuh"

namespace pbwtgpu {

constexpr int kRecChunkLog = 13;  // match records are allocated to windows in chunks of 8192 (128 KiB)
constexpr int kRecChunk = 1 << kRecChunkLog;

enum SweepMode { kGather = 0, kSorted = 1 };

// Columns [k0, k1) of window `win`.  Windows are spw sites long, except that windows 0 and 1 may be given `first` and
// `second` (<= spw) sites and the last window takes what remains: the edge windows of a panel carry extra match work
// (everything matches back to site 0 at the start, every block is reported at site N), so they get fewer sites.
__host__ __device__ inline void window_range(int win, int spw, int first, int second, int N, int &k0, int &k1) {
  if (first > 0) {
    k0 = win == 0 ? 0 : win == 1 ? first : first + second + (win - 2) * spw;
    k1 = win == 0 ? first : win == 1 ? first + second : k0 + spw;
  } else {
    k0 = win * spw;
    k1 = k0 + spw;
  }
  if (k1 > N) k1 = N;
}

// One launch sweeps W windows concurrently (one CTA per window), each starting from the exact
// state (a,d) entering its first site.
//   regular mode (perWindowSites == 0): window w covers sites [w*sitesPerWindow, min(N,(w+1)*sitesPerWindow))
//     of the shared column array src; site labels are the column indices.
//   per-window mode (perWindowSites > 0): window w sweeps columns [0, perWindowSites) of its own block
//     src + w*srcWindowStride; column j carries site label kLabelBase + j.  Used by the window-prefix
//     combine (windows.cu), whose "sites" are the key bits of a later window's tie groups.
struct SweepArgs {
  const uint32_t *src;  // gather mode: panel bit columns x_k; sorted mode: sorted columns y_k
  int wpc;              // words per column
  int M, N;
  int sitesPerWindow;
  int firstWindowSites;  // shared-memory kernel: > 0 gives window 0 this many sites (see window_range); 0 = like the others
  int secondWindowSites; // and window 1 this many (only with firstWindowSites > 0)
  int nWindows;
  int siteOffset;  // regular mode: global label of column 0 (a shard of a longer panel), 0 otherwise
  int nTotal;      // global site count (labels run to nTotal); 0 = N
  int perWindowSites;
  long long srcWindowStride;  // words
  int kLabelBase;
  const int32_t *aStart;  // [W][M] ([M] if aStartShared); NULL = identity
  int aStartShared;
  const int32_t *dStart;  // [W][M+1] interior values (sentinels are implied); NULL = fresh cursor (zeros)
  int dStartShared;       // dStart is one [M+1] array used by every window
  int32_t *aEnd;          // [W][M] state after the window's last site ([M], last window only, if endOnlyLast)
  int32_t *dEnd;          // [W][M+1]
  int endOnlyLast;
  uint32_t *yOut;         // [N][wpc] sorted columns y_k (input to the pack3 encoder), or NULL
  // pack3 run detection fused into the sweep (shared-memory kernel, gather mode): window w appends the run heads of its
  // columns, (position << 1 | allele), to p3Out + w * p3Cap -- site k's at offset p3SiteOff[k], p3SiteCnt[k] of them, in
  // no particular order -- and leaves their number in p3Count[w]; heads beyond p3Cap are dropped and p3Overflow[0] is set
  // (the host then falls back to yOut + the encoder over sorted columns).  pack3.cu turns the lists into bytes.
  uint32_t *p3Out;
  long long p3Cap;
  int32_t *p3Count;
  int32_t *p3SiteOff, *p3SiteCnt;  // [N]
  int *p3Overflow;
  int32_t *cOut;          // [N] zero counts c_k, or NULL
  uint32_t *xOut;         // sorted mode: [N][wpc] the UN-sorted columns x_k (bit a_k[i] = y_k[i]), written in REVERSE
                          // site order (row N-1-k): the input of the reverse build (pbwtCore.c:170-175); or NULL.
                          // Wide-panel kernel: must be zeroed by the caller (bits are OR-ed in)
  // match records: a pool of poolChunks chunks of kRecChunk records handed out through chunkCounter;
  // chunkTab[w*poolChunks + c] = pool chunk holding records [c*kRecChunk, (c+1)*kRecChunk) of window w
  pbwtgpu_match *recPool;  // NULL = count only
  int poolChunks;
  int *chunkCounter;
  int *chunkTab;
  unsigned long long *recCount;  // [W] records produced by each window
  int reportAtN;                 // the window that ends at N also reports the k == N blocks
  int longT;                     // -longWithin L: report matches of at least this many sites (match mode 2)
  long long *winNanos;           // [W] optional: wall time of each window's CTA (globaltimer, debugging aid)
  int debugNoCandidates;         // profiling aid: publish zero match candidates per site (sweep-side cost only)
  unsigned int *lenHist;         // -stats (pbwtMatch.c:130-131): [nTotal + 1] histogram of the reporters' match lengths, or NULL
  // query sweep (matchSequencesSweep, pbwtMatch.c:363-443): single window, sorted mode, with d
  const uint32_t *qcols;         // [N][qwpc] sorted query columns
  int qwpc, Q;
  const int32_t *qaStart;        // [Q] or NULL = identity
  pbwtgpu_nomatch *nomatch;      // "no match" events, in emission order
  long long nomatchCap;
  unsigned long long *qCounts;   // [0] records, [1] "no match" events
  // window-parallel query sweep (wide-panel kernel): W windows at once, query state [W][Q] in GlobalBuffers::qf / qe
  //   queryMode 0: one window from the fresh state (f = 0, e = first label)
  //   queryMode 1: CLAMPED pass: every window from the fresh state, divergence tiers below the window's first site
  //                count as that site, nothing is emitted; what is left in qf / qe is the window's end state
  //   queryMode 2: final pass: window w starts from the state already in qf / qe [w][.]; records and events are emitted
  int queryMode;
  int qaPerWindow;               // qaStart is [W][Q]: the query PBWT order entering every window
};


constexpr int kTabMirror = 64;  // newest chunk ids of a window mirrored in shared memory

#ifdef __CUDACC__
// Where a window's records go: chunks of kRecChunk records drawn from a global pool.
struct RecSink {
  int4 *pool;           // NULL = count only
  const int *tab;       // this window's chunk table (global)
  const int *mirror;    // shared-memory mirror of the newest kTabMirror entries
  int nChunks;          // chunks allocated so far (uniform)
  int cap;              // entries of the chunk table row (chunks beyond it do not exist: the pool overflowed)
  unsigned int *hist;   // -stats: histogram of the reporters' match lengths (device consumer of the matches), or NULL
  __device__ __forceinline__ void put(long long o, int4 rec) const {
    const int c = (int)(o >> kRecChunkLog);
    const int id = (nChunks - c <= kTabMirror) ? mirror[c & (kTabMirror - 1)] : (c < cap ? __ldcg(tab + c) : -1);
    if (id >= 0) pool[((long long)id << kRecChunkLog) + (o & (kRecChunk - 1))] = rec;
  }
};
#endif

// largest M the shared-memory-resident kernel handles with / without the divergence array
int sweep_smem_max_m(bool with_d);
// returns cudaError_t as int
// match_mode: 0 none, 1 matchMaximalWithin, 2 matchLongWithin2 with args.longT (both imply d)
int launch_sweep_smem(const SweepArgs &args, SweepMode mode, bool with_d, int match_mode, cudaStream_t stream);
// (args.p3Out != NULL selects the kernels with the fused pack3 encoder: gather mode, match_mode 0 or 1)
// query sweep: panel columns in args.src (sorted mode), queries in args.qcols; records go through the chunk pool
int query_sweep_max_m();
int query_sweep_max_q(int M);
int launch_query_sweep_smem(const SweepArgs &args, cudaStream_t stream);


// ---- wide panels: state in global memory, groups of cooperating CTAs (sweep_global.cu) ----
// plan: G CTAs per window and their chunk of sorted positions, W = windows swept concurrently, and whether each
// group is a thread-block cluster (M <= 16 * 32768) or part of a cooperative grid (W * G <= maxCtas)
void sweep_global_plan(int M, int maxCtas, int wantWindows, SweepMode mode, bool with_d, bool match, bool query, int Q,
                       bool allowCluster, int *G, int *chunk, int *W, int *cluster);
size_t sweep_global_state_ints(int G, int chunk, int W, bool with_d);  // a (and d) double buffers, in int32
size_t sweep_global_sync_bytes(int W);
size_t sweep_global_column_words(int G, int chunk, int W);  // words of ySorted / wmaxG / zpre (double buffered)
// device buffers of a wide-panel launch.  partials (W*G*4 uint32) and gsync (sweep_global_sync_bytes(W)) must be
// zeroed before every launch.
struct GlobalBuffers {
  int G = 1, chunk = 0;
  int cluster = 0;  // groups are thread-block clusters (sweep_global_plan)
  int32_t *aBuf = nullptr, *dBuf = nullptr;  // sweep_global_state_ints
  uint32_t *partials = nullptr;
  void *gsync = nullptr;
  // match scans only (maxWithin / query sweep): sweep_global_column_words each
  uint32_t *ySorted = nullptr;
  int32_t *wmaxG = nullptr;
  int32_t *zpre = nullptr;          // query sweep only
  uint32_t *any1 = nullptr, *any0 = nullptr;  // query sweep only: sweep_global_column_words / 32 each
  int32_t *wmax2 = nullptr;
  // record segments, one per (column, CTA of the window's group), (N+1)*G entries; chunk table rows are per CTA
  int32_t *segCnt = nullptr;
  long long *segSrc = nullptr;
  int32_t *qf = nullptr, *qe = nullptr;      // query sweep: [Q] each
  unsigned long long *nmKeys = nullptr;      // query sweep: [nomatchCap] sort keys of the "no match" events
};
int launch_sweep_global(const SweepArgs &a, SweepMode mode, bool with_d, int match_mode, const GlobalBuffers &b,
                        cudaStream_t stream);
// query sweep over a wide panel: one window, sorted mode, with d; every CTA of the group serves its share of the
// queries.  args.qCounts[1] counts the "no match" events (zero it before the launch).
int launch_query_sweep_global(const SweepArgs &a, const GlobalBuffers &b, cudaStream_t stream);
int query_sweep_global_max_q();

}  // namespace pbwtgpu
