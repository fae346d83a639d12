// match_scan.cuh -- warp-cooperative scans over one PBWT state (a_k, d_k, y_k), shared by the
// shared-memory sweep (sweep_smem.cu) and the wide-panel sweep (sweep_global.cu).
//
// Replaces the element-by-element while loops of matchMaximalWithin (pbwtMatch.c:122-137) and of
// matchSequencesSweep's alternate / reseed scans (pbwtMatch.c:381-411): 32 positions per step, and for
// long blocks 1024 positions per step through the per-word maxima of d.
//
// The state is reached through an accessor S (where the arrays live differs per kernel):
//   int      S::d(int i)      divergence d_k[i], i in [0, M]; d(0) = d(M) = k + 1 (pbwtCore.c:414,507)
//   uint32_t S::yw(int w)     word w of the sorted bit column y_k (bits beyond M are ones)
//   int      S::wmax(int w)   max of d over positions [32w, 32w + 32) (any value for words at or beyond M)
//   int      S::a(int i)      prefix array a_k[i]
#pragma once
#include "sweep.cuh"

namespace pbwtgpu {

template <class S>
__device__ __forceinline__ uint32_t ybit(const S &s, int j) { return (s.yw(j >> 5) >> (j & 31)) & 1u; }

// does any position in [lo, hi) carry allele yi?
template <class S>
__device__ __forceinline__ bool any_same_allele(const S &s, int lo, int hi, uint32_t yi, int lane) {
  if (lo >= hi) return false;
  const int w0 = lo >> 5, w1 = (hi - 1) >> 5;
  bool hit = false;
  for (int w = w0 + lane; w <= w1; w += 32) {
    uint32_t bits = yi ? s.yw(w) : ~s.yw(w);
    if (w == w0) bits &= kFull << (lo & 31);
    if (w == w1 && (hi & 31)) bits &= (1u << (hi & 31)) - 1u;
    hit |= bits != 0u;
  }
  return __any_sync(kFull, hit);
}

constexpr int kEarlySteps = 2;  // 32-position steps scanned with the reference's early exit before skipping by words

// nearest q <= from with d[q] > t (q = 0 when there is none: d[0] is the k+1 sentinel)
template <class S>
__device__ __forceinline__ int block_start(const S &s, int from, int t, int lane) {
  const int w = from >> 5;
  const int idx = (w << 5) + lane;
  const unsigned m = __ballot_sync(kFull, idx <= from && (idx <= 0 || s.d(idx) > t));
  if (m) return (w << 5) + 31 - __clz(m);
  int ws = w - 1;
  for (;;) {
    const int wi = ws - lane;
    const unsigned mw = __ballot_sync(kFull, wi < 0 || s.wmax(wi) > t);
    if (mw) { ws -= __ffs(mw) - 1; break; }
    ws -= 32;
  }
  if (ws < 0) return 0;
  const int idx2 = (ws << 5) + lane;
  const unsigned m2 = __ballot_sync(kFull, idx2 <= 0 || s.d(idx2) > t);
  return m2 ? (ws << 5) + 31 - __clz(m2) : 0;
}

// nearest r >= from with d[r] > t, or M
template <class S>
__device__ __forceinline__ int block_end(const S &s, int from, int t, int M, int lane) {
  const int w = from >> 5;
  const int idx = (w << 5) + lane;
  const unsigned m = __ballot_sync(kFull, idx >= from && (idx >= M || s.d(idx) > t));
  if (m) return min((w << 5) + __ffs(m) - 1, M);
  int ws = w + 1;
  for (;;) {
    const int wi = ws + lane;
    const unsigned mw = __ballot_sync(kFull, (wi << 5) >= M || s.wmax(wi) > t);
    if (mw) { ws += __ffs(mw) - 1; break; }
    ws += 32;
  }
  if ((ws << 5) >= M) return M;
  const int idx2 = (ws << 5) + lane;
  const unsigned m2 = __ballot_sync(kFull, idx2 >= M || s.d(idx2) > t);
  return m2 ? min((ws << 5) + __ffs(m2) - 1, M) : M;
}

// The reference's upward scan of position i (pbwtMatch.c:124-126):
//   lo = i; while (d[lo] <= d[i]) { --lo; if (y[lo] == y[i]) -> the match extends }
// First kEarlySteps x 32 positions exactly as the reference walks them (most candidates resolve here: a
// neighbour with the same allele, or the block's end); longer blocks are closed with the per-word maxima of d
// (1024 positions per step) and tested for the allele by masked popcounts.  Lane l of a step looks at
// p = top - l, visited iff no d[q] > di for q in (p, i].  Returns true when the match extends, else sets lo.
template <class S>
__device__ __forceinline__ bool scan_up(const S &s, int i, uint32_t yi, int di, bool live, int lane, int &lo) {
  int top = i - 1;
  for (int step = 0; step < kEarlySteps; ++step, top -= 32) {
    const int p = top - lane;
    const bool blk = p < -1 || (p + 1 < i && s.d(p + 1) > di);  // d[0] = k+1 blocks at p = -1
    const unsigned mBlk = __ballot_sync(kFull, blk);
    const unsigned before = mBlk ? ((1u << (__ffs(mBlk) - 1)) - 1u) : kFull;  // lanes visited in this step
    if (live) {
      const unsigned mSame = __ballot_sync(kFull, p >= 0 && ybit(s, p) == yi);
      if (mSame & before) return true;
    }
    if (mBlk) {
      lo = top - (__ffs(mBlk) - 1) + 1;
      return false;
    }
  }
  // positions (top, i) are visited and clean; nearest q <= top + 1 with d[q] > di closes the block
  lo = block_start(s, top + 1, di, lane);
  return live && any_same_allele(s, lo, top + 1, yi, lane);
}

// The downward scan (pbwtMatch.c:127-129): hi = i+1; while (d[hi] <= d[i+1]) { if (y[hi] == y[i]) extends; ++hi }
template <class S>
__device__ __forceinline__ bool scan_down(const S &s, int i, uint32_t yi, int dn1, int M, bool live, int lane, int &hi) {
  int bot = i + 1;
  for (int step = 0; step < kEarlySteps; ++step, bot += 32) {
    const int p = bot + lane;
    const bool blk = p >= M || s.d(p) > dn1;  // d[M] = k+1 blocks at p = M
    const unsigned mBlk = __ballot_sync(kFull, blk);
    const unsigned before = mBlk ? ((1u << (__ffs(mBlk) - 1)) - 1u) : kFull;
    if (live) {
      const unsigned mSame = __ballot_sync(kFull, p < M && ybit(s, p) == yi);
      if (mSame & before) return true;
    }
    if (mBlk) {
      hi = bot + __ffs(mBlk) - 1;
      return false;
    }
  }
  // positions (i, bot) are visited and clean; nearest r >= bot with d[r] > dn1 closes the block
  hi = block_end(s, bot, dn1, M, lane);
  return live && any_same_allele(s, bot, hi, yi, lane);
}

// write the records of reporter i: upward block [lo, i) with start di, then downward block (i, hi) with
// start dn1 (callback order, pbwtMatch.c:133-134); record t of the block is written by "lane" t0 + t*step
template <class S>
__device__ __forceinline__ void emit_block(const RecSink &out, long long o, const S &s, int i, int lo, int hi, int di,
                                           int dn1, int k, int t0, int step) {
  const int cnt = (i - lo) + (hi - i - 1);
  const int ai = s.a(i);
  const int nUp = i - lo;
  for (int t = t0; t < cnt; t += step) {
    const int jj = t < nUp ? lo + t : t + lo + 1;  // skips i itself
    out.put(o + t, make_int4(ai, s.a(jj), t < nUp ? di : dn1, k));
  }
}

// ---- matchMaximalWithin's per-site scan (pbwtMatch.c:122-137), block-wide and load-balanced ----
// For a position i the upward block is [q, i) with q the nearest index <= i whose d exceeds d[i], the
// downward block (i, r) with r the nearest index > i whose d exceeds d[i+1]; i reports nothing when a
// member of either block carries the same allele (the match extends).  Only heads/tails of allele runs
// can fail that test on their first neighbour ("candidates"); PBWT order clusters them in a few words, so
// they are numbered block-wide and dealt round-robin to the 32 warps, kResMax per round.

// Candidate lanes of one word of state k.  cw: the word's alleles, pb: allele before the word, nb: allele
// after it.  A lane is a candidate unless its first upward/downward neighbour already carries its allele in
// the direction(s) the reference scans (pbwtMatch.c:124-129).
// upExtends (per lane, optional knowledge of the caller): this lane heads a run and max d over the whole previous run is
// <= its own d, i.e. the upward scan would walk through the previous run -- all of the other allele -- and reach the
// run before it, which carries this lane's allele again: the match extends (pbwtMatch.c:124-126) without any scan.  The
// sweep has that maximum in hand: it is what the divergence update of a run head needs (pbwtCore.c:496-505).
template <class S>
__device__ __forceinline__ uint32_t candidate_mask(const S &s, uint32_t cw, uint32_t pb, uint32_t nb, int base, uint32_t vm,
                                                   int lane, bool upExtends = false) {
  const uint32_t H = cw ^ ((cw << 1) | pb);
  const uint32_t Hn = (H >> 1) | (((nb & 1u) ^ (cw >> 31)) << 31);
  const uint32_t bm = (H | Hn) & vm;
  bool mine = false;
  if ((bm >> lane) & 1u) {
    const int di = s.d(base + lane), dn1 = s.d(base + lane + 1);
    const bool samePrev = !((H >> lane) & 1u) || upExtends, sameNext = !((Hn >> lane) & 1u);
    mine = !((di <= dn1 && samePrev) || (di >= dn1 && sameNext));
  }
  return __ballot_sync(kFull, mine);
}

// position of the g-th candidate of the block (g < T).  candBase: lane l holds the number of candidates in
// warps < l; masks: per-word candidate masks, E (<= 32) words per warp.
__device__ __forceinline__ int locate_candidate(const uint32_t *masks, int E, int candBase, int g, int lane) {
  const int wv = 31 - __clz(__ballot_sync(kFull, candBase <= g));  // owning warp
  int r = g - __shfl_sync(kFull, candBase, wv);                    // rank inside that warp
  const uint32_t m = lane < E ? masks[wv * E + lane] : 0u;
  const int pc = __popc(m);
  int inc = pc;
#pragma unroll
  for (int dl = 1; dl < 32; dl <<= 1) {
    const int t = __shfl_up_sync(kFull, inc, dl);
    if (lane >= dl) inc += t;
  }
  const int e = __ffs(__ballot_sync(kFull, inc > r)) - 1;  // word holding it
  r -= __shfl_sync(kFull, inc - pc, e);
  const uint32_t me = __shfl_sync(kFull, m, e);
  uint32_t mm = me;
  for (int q = 0; q < r; ++q) mm &= mm - 1;  // drop the r lowest set bits (r is small)
  return (wv * E + e) * 32 + __ffs(mm) - 1;
}

constexpr int kResMax = 256;    // match candidates processed per round (multiple of 32)
constexpr int kBigBlock = 128;  // records; larger blocks are emitted by the whole CTA

// -------- candidates of a state -> records, in rounds of kResMax dealt round-robin to the 32 warps --------
// Called by all 1024 threads of a CTA.  Expects cmaskS[] (per-word candidate masks, E words per warp) and
// wsumC[] (per-warp candidate counts) to have been published before the preceding barrier.  Contains at least
// one __syncthreads.  ensure(upTo) makes sure record offsets below upTo have chunks (uniform call).
// posOffset: sorted position of bit 0 of cmaskS[0] (the masks of a CTA that owns a slice of the positions).
template <class S, class Ensure>
__device__ __forceinline__ void match_rounds(const S &s, int k, bool live, int M, int E, const int32_t *wsumC,
                                             const uint32_t *cmaskS, int4 *resS, RecSink &sink, long long &recBase,
                                             Ensure &&ensure, int tid, int lane, int warp, int posOffset = 0) {
  const int myCnt = wsumC[lane];
  const int T = __reduce_add_sync(kFull, myCnt);
  if (T == 0) {
    __syncthreads();
    return;
  }
  int candBase = 0;
  if (warp < T) {  // only warps that will own a candidate need the per-warp bases
    int x = myCnt;
#pragma unroll
    for (int dl = 1; dl < 32; dl <<= 1) {
      const int t = __shfl_up_sync(kFull, x, dl);
      if (lane >= dl) x += t;
    }
    candBase = x - myCnt;
  }
  for (int g0 = 0; g0 < T; g0 += kResMax) {
    const int nR = min(kResMax, T - g0);
    for (int g = g0 + warp; g < g0 + nR; g += 32) {
      const int i = posOffset + locate_candidate(cmaskS, E, candBase, g, lane);
      const uint32_t yi = ybit(s, i);
      const int di = s.d(i), dn1 = s.d(i + 1);
      int lo = i, hi = i + 1;
      bool extends = false;
      if (di <= dn1) extends = scan_up(s, i, yi, di, live, lane, lo);
      if (!extends && di >= dn1) extends = scan_down(s, i, yi, dn1, M, live, lane, hi);
      if (lane == 0) {
        resS[g - g0] = make_int4(i, lo, hi, extends ? 0 : (i - lo) + (hi - i - 1));
        // -stats (pbwtMatch.c:130-131): one count per reporter, at the length of the longer of its two blocks
        if (sink.hist && !extends) atomicAdd(sink.hist + (k - min(di, dn1)), 1u);
      }
    }
    __syncthreads();  // the round's results are published (first round: the scatter is complete too)
    long long roundTotal = 0;
    for (int c0 = 0; c0 < nR; c0 += 32) {
      const int cntL = c0 + lane < nR ? resS[c0 + lane].w : 0;  // record count of entry c0 + lane
      const int tot = __reduce_add_sync(kFull, cntL);
      if (sink.pool && tot) {
        const long long base = recBase + roundTotal;
        ensure(base + tot);
        // entry c0 + warp is this warp's own (candidate g = g0 + c0 + warp); big blocks are shared by all warps
        const int mcnt = __shfl_sync(kFull, cntL, warp);
        if (mcnt > 0 && mcnt < kBigBlock) {
          const int4 en = resS[c0 + warp];
          const int mex = __reduce_add_sync(kFull, lane < warp ? cntL : 0);
          emit_block(sink, base + mex, s, en.x, en.y, en.z, s.d(en.x), s.d(en.x + 1), k, lane, 32);
        }
        unsigned big = __ballot_sync(kFull, cntL >= kBigBlock);
        while (big) {
          const int j = __ffs(big) - 1;
          big &= big - 1;
          const int4 en = resS[c0 + j];
          const int bex = __reduce_add_sync(kFull, lane < j ? cntL : 0);
          emit_block(sink, base + bex, s, en.x, en.y, en.z, s.d(en.x), s.d(en.x + 1), k, tid, 1024);
        }
      }
      roundTotal += tot;
    }
    recBase += roundTotal;
    if (g0 + kResMax < T) __syncthreads();  // resS is reused by the next round
  }
}

// ---- matchLongWithin2's per-site scan (-longWithin L, pbwtMatch.c:85-113) ----
// State k is cut into blocks at every position with d > thr = k - L.  A block that holds both alleles and does not
// run to M is reported: all pairs ia < ib with different alleles, start = max d(ia, ib], in (ia, ib) order.
// A block holds both alleles iff it contains an "interior head": a position h with y[h] != y[h-1] and d[h] <= thr.
// Interior heads are the candidates; the first one of each block owns the block, one warp emits it.

// interior heads of one word.  cw: the word's alleles, pb: allele before the word.
template <class S>
__device__ __forceinline__ uint32_t long_candidate_mask(const S &s, uint32_t cw, uint32_t pb, int base, uint32_t vm, int thr,
                                                        int lane) {
  const uint32_t H = (cw ^ ((cw << 1) | pb)) & vm;
  bool mine = false;
  if ((H >> lane) & 1u) mine = s.d(base + lane) <= thr;
  return __ballot_sync(kFull, mine);
}

// zero alleles among positions [lo, hi)
template <class S>
__device__ __forceinline__ int count_zeros(const S &s, int lo, int hi, int lane) {
  if (lo >= hi) return 0;
  const int w0 = lo >> 5, w1 = (hi - 1) >> 5;
  int n = 0;
  for (int w = w0 + lane; w <= w1; w += 32) {
    uint32_t bits = ~s.yw(w);
    if (w == w0) bits &= kFull << (lo & 31);
    if (w == w1 && (hi & 31)) bits &= (1u << (hi & 31)) - 1u;
    n += __popc(bits);
  }
  return __reduce_add_sync(kFull, n);
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int dl = 16; dl > 0; dl >>= 1) v += __shfl_xor_sync(kFull, v, dl);
  return v;
}

// all records of block [q, r), by one warp: rows ia ascending, 32 positions ib per step with a running max of d
template <class S>
__device__ __forceinline__ void emit_long_block(const RecSink &out, long long o, const S &s, int q, int r, int k, int lane) {
  const uint32_t ltMask = lanemask_lt(lane);
  for (int ia = q; ia + 1 < r; ++ia) {
    const uint32_t yia = ybit(s, ia);
    const int aia = s.a(ia);
    int running = 0;
    for (int base = ia + 1; base < r; base += 32) {
      const int ib = base + lane;
      const bool valid = ib < r;
      int v = valid ? s.d(ib) : 0;
#pragma unroll
      for (int dl = 1; dl < 32; dl <<= 1) {
        const int t = __shfl_up_sync(kFull, v, dl);
        if (lane >= dl) v = max(v, t);
      }
      v = max(v, running);
      const bool diff = valid && ybit(s, ib) != yia;
      const unsigned bal = __ballot_sync(kFull, diff);
      if (diff) out.put(o + __popc(bal & ltMask), make_int4(aia, s.a(ib), v, k));
      o += __popc(bal);
      running = __shfl_sync(kFull, v, 31);
    }
  }
}

// Same calling convention as match_rounds; cmaskS[] holds the interior heads.
template <class S, class Ensure>
__device__ __forceinline__ void long_rounds(const S &s, int k, int thr, int M, int E, const int32_t *wsumC,
                                            const uint32_t *cmaskS, int4 *resS, RecSink &sink, long long &recBase,
                                            Ensure &&ensure, int tid, int lane, int warp, int posOffset = 0) {
  const int myCnt = wsumC[lane];
  const int T = __reduce_add_sync(kFull, myCnt);
  if (T == 0) {
    __syncthreads();
    return;
  }
  int candBase = 0;
  if (warp < T) {
    int x = myCnt;
#pragma unroll
    for (int dl = 1; dl < 32; dl <<= 1) {
      const int t = __shfl_up_sync(kFull, x, dl);
      if (lane >= dl) x += t;
    }
    candBase = x - myCnt;
  }
  for (int g0 = 0; g0 < T; g0 += kResMax) {
    const int nR = min(kResMax, T - g0);
    for (int g = g0 + warp; g < g0 + nR; g += 32) {
      const int h = posOffset + locate_candidate(cmaskS, E, candBase, g, lane);
      const uint32_t yh = ybit(s, h);
      const int q = block_start(s, h, thr, lane);  // d[h] <= thr, so q < h
      long long cnt = 0;
      int r = 0;
      if (!any_same_allele(s, q, h, yh, lane)) {  // the block's first interior head owns it
        r = block_end(s, h + 1, thr, M, lane);
        if (r < M) {  // the reference never reports the block that runs to M (pbwtMatch.c:93-104)
          const long long z = count_zeros(s, q, r, lane);
          cnt = z * (long long)(r - q - z);
        }
      }
      if (lane == 0) resS[g - g0] = make_int4(q, r, (int)(cnt & 0xffffffffll), (int)(cnt >> 32));
    }
    __syncthreads();
    long long roundTotal = 0;
    for (int c0 = 0; c0 < nR; c0 += 32) {
      const int4 en0 = c0 + lane < nR ? resS[c0 + lane] : make_int4(0, 0, 0, 0);
      const long long cntL = ((long long)en0.w << 32) | (unsigned)en0.z;
      const long long tot = warp_sum_ll(cntL);
      if (sink.pool && tot) {
        const long long base = recBase + roundTotal;
        ensure(base + tot);
        const long long mex = warp_sum_ll(lane < warp ? cntL : 0);
        const long long mcnt = __shfl_sync(kFull, cntL, warp);
        if (mcnt > 0) {
          const int4 en = resS[c0 + warp];
          emit_long_block(sink, base + mex, s, en.x, en.y, k, lane);
        }
      }
      roundTotal += tot;
    }
    recBase += roundTotal;
    // unlike matchMaximalWithin's emission, this one reads the alleles: no warp may start the next site (and
    // overwrite the sorted column in shared memory) or the next round before every warp is done here
    __syncthreads();
  }
}

}  // namespace pbwtgpu
