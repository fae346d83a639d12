// sweep_global.cu -- per-site forward sweep for panels too wide for one CTA's shared memory
// (M > 13,312 with d / 25,600 without; BASELINE configs 3-5: M = 1e5 .. 1e6).
//
// Same algorithm and outputs as sweep_smem.cu (pbwtCursorForwardsA/AD, the gather build, sorted-column
// sweeps: pbwtCore.c:458-557, pbwtIO.c:572-574), different residency: a[] and d[] live in global memory,
// double buffered, and stay L2-resident (16 B per haplotype: 16 MB at M = 1e6, L2 is 126 MB).
// A window of sites is swept by a GROUP of G co-resident CTAs (cooperative launch); CTA g of a group owns
// the contiguous sorted positions [g*chunk, (g+1)*chunk).  Per site:
//   phase A  each CTA: coalesced ld.cg of a[]/d[] (L1 must be bypassed: other CTAs rewrite them every
//            site), gather of the allele bit from the site's column (read-only, L1-cached), ballots,
//            warp-local zero counts and the run-max scan for d (as in sweep_smem.cu); d' staged in smem.
//   exchange each CTA publishes (zeros, run-max summary) with an epoch flag; every CTA reads all G partials
//            (flag spin, no barrier) and derives its incoming zero count / carry and the column's c.
//   phase C  stable-partition scatter of a (and d') into the other buffer.
//   barrier  one group-wide barrier per site (monotonic arrival counter in global memory).
// Windows are independent groups, so the window-parallel scheme of windows.cu applies unchanged.
#include "sweep.cuh"

namespace pbwtgpu {
namespace {

constexpr int kThreads = 1024;
constexpr int kMaxChunk = 32768;  // sorted positions per CTA (d' staging: 128 KiB of shared memory)
constexpr int kBatch = 8;         // words per warp whose global loads are issued back to back

struct GroupSync {  // per group, in global memory
  unsigned int arrive;         // monotonic barrier counter
  unsigned int pad[31];
};

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int &total) {
  int x = v;
#pragma unroll
  for (int dl = 1; dl < 32; dl <<= 1) {
    const int t = __shfl_up_sync(kFull, x, dl);
    if (lane >= dl) x += t;
  }
  total = __shfl_sync(kFull, x, 31);
  return x - v;
}

__device__ __forceinline__ unsigned ld_volatile(const unsigned *p) {
  unsigned v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// run-max summary: bit 31 = "contains a run head", low 31 bits = max d since the last head (or overall)
__device__ __forceinline__ uint32_t rm_combine(uint32_t a, uint32_t b) {  // a earlier, b later
  return (b >> 31) ? b : ((a & 0x80000000u) | max(a & 0x7fffffffu, b));
}

constexpr int kQCap = 1024;  // queries handled by the query sweep (one thread each, CTA 0 of the group)

template <int MODE, bool WITH_D, bool QUERY = false>
__global__ void __launch_bounds__(kThreads, 1) sweep_global_kernel(const SweepArgs p, int G, int chunk, int32_t *aBuf,
                                                                   int32_t *dBuf, uint32_t *partials, GroupSync *gsync,
                                                                   uint32_t *ySorted, int32_t *zpre) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int chunkWords = chunk / 32;
  uint32_t *ycolS = smem;                                               // [chunkWords]
  int32_t *dnS = reinterpret_cast<int32_t *>(smem + chunkWords);        // [chunk] (WITH_D)
  // query sweep only: warp-local zero prefix per word, per-query state, chunk-id mirror, per-warp sums
  int32_t *zlocS = dnS + (WITH_D ? chunk : 0);                          // [chunkWords]
  int32_t *qfS = zlocS + chunkWords, *qeS = qfS + kQCap, *qaS = qeS + kQCap;  // qaS: [2][kQCap]
  int32_t *ctabS = qaS + 2 * kQCap;                                     // [kTabMirror]
  int32_t *wsq = ctabS + kTabMirror;                                    // [96]
  __shared__ int wexS[32];      // zeros in this CTA's warps below each warp
  __shared__ int wsumZ[32];
  __shared__ uint32_t wsumR[32];
  __shared__ int xZ, xC;        // zeros before this CTA, zeros in the whole column
  __shared__ uint32_t xR;       // run-max carry entering this CTA

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int win = blockIdx.x / G, g = blockIdx.x % G;
  const int M = p.M, wpc = p.wpc;
  const int nTotal = p.nTotal ? p.nTotal : p.N;
  const size_t MS = (size_t)G * chunk;  // padded state stride per buffer
  int32_t *aW = aBuf + (size_t)win * 2 * MS;
  int32_t *dW = WITH_D ? dBuf + (size_t)win * 2 * (MS + 4) : nullptr;
  uint32_t *part = partials + (size_t)win * G * 4;  // per CTA: zeros, rm, flag, unused
  GroupSync *gs = gsync + win;

  int k0, k1, colShift;
  const uint32_t *src;
  if (p.perWindowSites) {
    k0 = p.kLabelBase; k1 = k0 + p.perWindowSites; colShift = -k0;
    src = p.src + (size_t)win * p.srcWindowStride;
  } else {
    k0 = p.siteOffset + win * p.sitesPerWindow; k1 = min(p.siteOffset + p.N, k0 + p.sitesPerWindow); colShift = -p.siteOffset;
    src = p.src;
  }
  const int p0 = g * chunk;                 // first sorted position of this CTA
  const int EW = chunkWords / 32;           // words per warp
  const uint32_t ltMask = lanemask_lt(lane), laneBit = 1u << lane;
  unsigned int epoch = 0;                   // barriers passed so far

  auto group_barrier = [&]() {
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(&gs->arrive, 1u);
      const unsigned target = (epoch + 1) * (unsigned)G;
      while (ld_volatile(&gs->arrive) < target) {}
      __threadfence();
    }
    ++epoch;
    __syncthreads();
  };

  RecSink sink;
  sink.pool = QUERY ? reinterpret_cast<int4 *>(p.recPool) : nullptr;
  sink.tab = QUERY ? p.chunkTab : nullptr;
  sink.mirror = ctabS;
  sink.nChunks = 0;
  long long recBase = 0, nmBase = 0;
  uint32_t qWord = 0;
  auto ensure_chunks = [&](long long upTo) {  // CTA 0 only
    const int need = (int)((upTo + kRecChunk - 1) >> kRecChunkLog);
    if (need > sink.nChunks) {
      if (tid == 0)
        for (int c2 = sink.nChunks; c2 < need; ++c2) {
          int id = atomicAdd(p.chunkCounter, 1);
          if (id >= p.poolChunks) id = -1;
          p.chunkTab[c2] = id;
          ctabS[c2 & (kTabMirror - 1)] = id;
        }
      sink.nChunks = need;
      __syncthreads();
    }
  };
  if (QUERY && g == 0) {
    for (int i = tid; i < kQCap; i += kThreads) {
      qfS[i] = 0;
      qeS[i] = 0;
      qaS[i] = i < p.Q ? (p.qaStart ? p.qaStart[i] : i) : 0;
    }
    if (tid < p.Q && k0 < k1) qWord = p.qcols[(size_t)(k0 + colShift) * p.qwpc + warp];
  }

  // ---- window start state ----
  for (int i = p0 + tid; i < p0 + chunk; i += kThreads) {
    if (i < M) {
      aW[i] = p.aStart ? p.aStart[(p.aStartShared ? 0 : (size_t)win * M) + i] : i;
      if (WITH_D) dW[i] = (i > 0 && p.dStart) ? p.dStart[(p.dStartShared ? 0 : (size_t)win * (M + 1)) + i] : 0;
    }
  }
  group_barrier();

  int cur = 0;
  for (int k = k0; k < k1; ++k) {
    const uint32_t *col = src + (size_t)(k + colShift) * wpc;
    const int32_t *aC = aW + (size_t)cur * MS;
    int32_t *aN = aW + (size_t)(cur ^ 1) * MS;
    const int32_t *dC = WITH_D ? dW + (size_t)cur * (MS + 4) : nullptr;
    int32_t *dN = WITH_D ? dW + (size_t)(cur ^ 1) * (MS + 4) : nullptr;

    // ---------------- phase A ----------------
    const int wordBase = warp * EW;          // CTA-local word index of this warp's first word
    const int posBase = p0 + wordBase * 32;  // its first sorted position
    uint32_t pb;                             // allele just before this warp's first position
    if (posBase == 0 || posBase > M) pb = 2u;  // resolved at the first word: position 0 is a run head
    else if (MODE == kGather) {
      const int hp = __ldcg(aC + posBase - 1);
      pb = (__ldg(col + (hp >> 5)) >> (hp & 31)) & 1u;
    } else pb = __ldg(col + ((posBase - 1) >> 5)) >> 31;
    int zw = 0, carry = 0, lmax = 0;
    bool anyHead = false;
    int patchIdx = -1;  // CTA-local index (word*32+lane) of the warp's first run head, held by that lane only
    for (int e0 = 0; e0 < EW; e0 += kBatch) {  // kBatch words in flight per warp: the loads are L2 round trips
      int hB[kBatch], dB[kBatch];
      uint32_t cB[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int i = posBase + (e0 + u) * 32 + lane;
        const bool valid = e0 + u < EW && i < M;
        hB[u] = (MODE == kGather && valid) ? __ldcg(aC + i) : 0;
        dB[u] = (WITH_D && valid) ? __ldcg(dC + i) : 0;
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int base = posBase + (e0 + u) * 32;
        const bool valid = e0 + u < EW && base + lane < M;
        if (MODE == kGather) cB[u] = valid ? __ldg(col + (hB[u] >> 5)) : kFull;
        else cB[u] = (e0 + u < EW && base < M) ? __ldg(col + (base >> 5)) : 0u;
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int e = e0 + u;
        if (e >= EW) break;
        const int i = posBase + e * 32 + lane;
        uint32_t Bw;
        if (MODE == kGather) {
          Bw = __ballot_sync(kFull, (cB[u] >> (hB[u] & 31)) & 1u);
        } else {
          const int base = posBase + e * 32;
          const uint32_t vm = (M - base >= 32) ? kFull : (M > base ? ((1u << (M - base)) - 1u) : 0u);
          Bw = cB[u] | ~vm;
        }
        if (pb == 2u) pb = (~Bw) & 1u;
        const uint32_t H = Bw ^ ((Bw << 1) | pb);
        pb = Bw >> 31;
        if (lane == 0) ycolS[wordBase + e] = Bw;
        if (QUERY && lane == 0) zlocS[wordBase + e] = zw;  // zeros of this warp before the word
        if (WITH_D) {
          int dv = dB[u];
          if (i == 0) dv = k + 1;  // sentinel d[0] (pbwtCore.c:414,507)
          int dnv = dv;
          if (H == 0u) {
            lmax = max(lmax, dv);
          } else {
            carry = max(carry, __reduce_max_sync(kFull, lmax));
            lmax = 0;
            int v = dv;
#pragma unroll
            for (int dl = 1; dl < 32; dl <<= 1) {
              const int t = __shfl_up_sync(kFull, v, dl);
              if (lane >= dl && ((H >> (lane - dl + 1)) & ((1u << dl) - 1u)) == 0u) v = max(v, t);
            }
            const int prev = __shfl_up_sync(kFull, v, 1);
            const bool headBelow = (H & ltMask) != 0u;
            const int prevMax = lane == 0 ? carry : (headBelow ? prev : max(carry, prev));
            if (H & laneBit) dnv = max(dv, prevMax);
            if (!anyHead) {
              anyHead = true;
              if (lane == __ffs(H) - 1) patchIdx = (wordBase + e) * 32 + lane;
            }
            carry = __shfl_sync(kFull, v, 31);
          }
          dnS[(wordBase + e) * 32 + lane] = dnv;
        }
        zw += __popc(~Bw);
      }
    }
    if (WITH_D) carry = max(carry, __reduce_max_sync(kFull, lmax));
    if (lane == 0) {
      wsumZ[warp] = zw;
      if (WITH_D) wsumR[warp] = (anyHead ? 0x80000000u : 0u) | (uint32_t)carry;
    }
    __syncthreads();

    // ---------------- exchange: CTA totals out, all partials of the group in ----------------
    int zExLane, zCta;
    zExLane = warp_excl_scan(wsumZ[lane], lane, zCta);
    uint32_t rInc = WITH_D ? wsumR[lane] : 0u;  // inclusive run-max scan over the CTA's warps
    if (WITH_D) {
#pragma unroll
      for (int dl = 1; dl < 32; dl <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, rInc, dl);
        if (lane >= dl) rInc = rm_combine(t, rInc);
      }
    }
    if (warp == 0) {
      const unsigned tag = (unsigned)(k - k0) + 1u;
      const uint32_t rCta = __shfl_sync(kFull, rInc, 31);
      if (lane == 0) {
        part[g * 4 + 0] = (uint32_t)zCta;
        part[g * 4 + 1] = rCta;
        __threadfence();
        atomicExch(&part[g * 4 + 2], tag);
      }
      // gather the group's partials (G <= 148 -> up to 5 per lane)
      int zBefore = 0, zAll = 0;
      uint32_t rBefore = 0u;  // combined summary of CTAs below g, in order
      for (int base = 0; base < G; base += 32) {
        const int gg = base + lane;
        uint32_t z = 0, r = 0;
        if (gg < G) {
          while (ld_volatile(&part[gg * 4 + 2]) != tag) {}
          __threadfence();
          z = ld_volatile(&part[gg * 4 + 0]);
          r = ld_volatile(&part[gg * 4 + 1]);
        }
        zAll += __reduce_add_sync(kFull, (int)z);
        zBefore += __reduce_add_sync(kFull, gg < g ? (int)z : 0);
        if (WITH_D) {  // ordered combine of the CTAs below g within this batch of 32
          uint32_t x = gg < g ? r : 0u;  // identity = (no head, 0)
#pragma unroll
          for (int dl = 1; dl < 32; dl <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, x, dl);
            if (lane >= dl) x = rm_combine(t, x);
          }
          rBefore = rm_combine(rBefore, __shfl_sync(kFull, x, 31));
        }
      }
      if (lane == 0) { xZ = zBefore; xC = zAll; xR = rBefore; }
      wexS[lane] = zExLane;
    }
    __syncthreads();
    const int c = xC;
    const int zIn = xZ + __shfl_sync(kFull, zExLane, warp);
    int cIn = 0;
    if (WITH_D) {  // carry entering this warp: CTAs below, then this CTA's warps below
      const uint32_t below = __shfl_sync(kFull, rInc, warp > 0 ? warp - 1 : 0);
      const uint32_t comb = warp > 0 ? rm_combine(xR, below) : xR;
      cIn = (int)(comb & 0x7fffffffu);
    }

    const int par = (k - k0) & 1;
    const int WG = G * chunkWords;  // words of the whole (padded) sorted column
    if (QUERY) {  // publish this CTA's part of the sorted column and of the zero prefix, then meet the group
      for (int w = tid; w < chunkWords; w += kThreads) {
        ySorted[(size_t)par * WG + p0 / 32 + w] = ycolS[w];
        zpre[(size_t)par * WG + p0 / 32 + w] = xZ + wexS[w / EW] + zlocS[w];
      }
      group_barrier();
    }

    // ---------------- phase C: scatter ----------------
    {
      int zacc = zIn;
      for (int e0 = 0; e0 < EW; e0 += kBatch) {
        int hB[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int i = p0 + (wordBase + e0 + u) * 32 + lane;
          hB[u] = (e0 + u < EW && i < M) ? __ldcg(aC + i) : 0;
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int e = e0 + u;
          if (e >= EW) break;
          const int li = (wordBase + e) * 32 + lane;
          const int i = p0 + li;
          const uint32_t Bw = ycolS[wordBase + e];
          if (i < M) {
            const int z = zacc + __popc(~Bw & ltMask);
            const int pos = (Bw & laneBit) ? c + i - z : z;
            aN[pos] = hB[u];
            if (WITH_D) {
              int dv = dnS[li];
              if (li == patchIdx) dv = max(dv, cIn);
              dN[pos] = dv;
            }
          }
          zacc += __popc(~Bw);
        }
      }
    }
    if (p.yOut) {  // sorted column k, pad bits cleared
      for (int w = tid; w < chunkWords; w += kThreads) {
        const int gw = p0 / 32 + w;
        if (gw < wpc) {
          const int base = gw * 32;
          const uint32_t vm = (M - base >= 32) ? kFull : (M > base ? ((1u << (M - base)) - 1u) : 0u);
          p.yOut[(size_t)(k + colShift) * wpc + gw] = ycolS[w] & vm;
        }
      }
    }
    if (p.cOut && g == 0 && tid == 0) p.cOut[k + colShift] = c;
    if (QUERY && g == 0) {
      // ---------------- matchSequencesSweep on state k (pbwtMatch.c:376-426), one thread per query ----------------
      const uint32_t *yk = ySorted + (size_t)par * WG;
      const int32_t *zk = zpre + (size_t)par * WG;
      auto dAt = [&](int idx) { return (idx == 0 || idx == M) ? k + 1 : __ldcg(dC + idx); };
      auto yAt = [&](int j) { return (__ldcg(yk + (j >> 5)) >> (j & 31)) & 1u; };
      const bool active = tid < p.Q;
      const uint32_t x = active ? (qWord >> lane) & 1u : 1u;
      if (active && k + 1 < k1) qWord = p.qcols[(size_t)(k + 1 + colShift) * p.qwpc + warp];
      const int qC = par * kQCap, qN = (par ^ 1) * kQCap;
      const int jj = active ? qaS[qC + tid] : 0;
      int f = qfS[jj], e = qeS[jj];
      int repLo = 0, repHi = 0, repStart = 0;
      bool nm = false;
      if (active && yAt(f) != x) {
        int hi = f + 1;
        bool moved = false;
        while (hi < M && dAt(hi) <= e) {  // an alternate of the same length that extends? (pbwtMatch.c:381-383)
          if (yAt(hi) == x) { f = hi; moved = true; break; }
          ++hi;
        }
        if (!moved) {
          repLo = f; repHi = hi; repStart = e;  // report [f, hi) with start e (pbwtMatch.c:385)
          int lo = f;
          int dDown = hi < M ? dAt(hi) : k;
          int dUp = dAt(lo);
          for (;;) {  // reseed by divergence tiers, upper side first on ties (pbwtMatch.c:389-411)
            if (dUp <= dDown) {
              int top = -1;
              while (lo > 0 && dAt(lo) <= dUp) {
                --lo;
                if (yAt(lo) == x) top = lo;
              }
              if (top >= 0) { f = top; e = dUp; break; }
              dUp = dAt(lo);
            } else {
              bool hit = false;
              while (hi < M && dAt(hi) <= dDown) {
                if (yAt(hi) == x) { f = hi; e = dDown; hit = true; break; }
                ++hi;
              }
              if (hit) break;
              dDown = hi == M ? k : dAt(hi);
              if (lo == 0 && hi == M) { nm = true; e = k + 1; break; }
            }
          }
        }
      }
      {  // FM step (pbwtCursorCalculateU + pbwtCursorMap, pbwtCore.c:510-519, pbwt.h:130; pbwtMatch.c:417-423)
        const int fw = f >> 5;
        const int Z = __ldcg(zk + fw) + __popc(~__ldcg(yk + fw) & lanemask_lt(f & 31));
        int nf = x ? c + f - Z : Z;
        if (nf == M) nf = 0;
        if (active) { qfS[jj] = nf; qeS[jj] = e; }
      }
      const int cnt = repHi - repLo;
      int cntTot;
      const int cntEx = warp_excl_scan(cnt, lane, cntTot);
      const unsigned nmBal = __ballot_sync(kFull, nm), qzBal = __ballot_sync(kFull, active && x == 0u);
      if (lane == 0) { wsq[warp] = cntTot; wsq[32 + warp] = __popc(nmBal); wsq[64 + warp] = __popc(qzBal); }
      __syncthreads();
      int recTot, nmTot, cq;
      const int recEx = __shfl_sync(kFull, warp_excl_scan(wsq[lane], lane, recTot), warp);
      const int nmEx = __shfl_sync(kFull, warp_excl_scan(wsq[32 + lane], lane, nmTot), warp);
      const int qzEx = __shfl_sync(kFull, warp_excl_scan(wsq[64 + lane], lane, cq), warp);
      if (sink.pool && recTot) {
        ensure_chunks(recBase + recTot);
        long long o = recBase + recEx + cntEx;
        for (int i = repLo; i < repHi; ++i, ++o) sink.put(o, make_int4(jj, __ldcg(aC + i), repStart, k));
      }
      if (nm && p.nomatch) {
        const long long o = nmBase + nmEx + __popc(nmBal & ltMask);
        if (o < p.nomatchCap) p.nomatch[o] = pbwtgpu_nomatch{jj, (int)x, k};
      }
      recBase += recTot;
      nmBase += nmTot;
      if (active) {  // query cursor step (pbwtCursorForwardsRead on uq, pbwtMatch.c:426)
        const int zb = qzEx + __popc(qzBal & ltMask);
        qaS[qN + (x ? cq + tid - zb : zb)] = jj;
      }
      __syncthreads();  // wsq is reused at the next site
    }
    group_barrier();
    cur ^= 1;
  }

  if (QUERY && g == 0) {  // matches still open at N (pbwtMatch.c:430-436): a[f], then every following i with d[i] <= e
    const int32_t *aE = aW + (size_t)cur * MS;
    const int32_t *dE = dW + (size_t)cur * (MS + 4);
    const bool active = tid < p.Q;
    const int jj = active ? qaS[((k1 - k0) & 1) * kQCap + tid] : 0;
    const int f = qfS[jj], e = qeS[jj];
    int hi = f + 1;
    if (active) while (hi < M && __ldcg(dE + hi) <= e) ++hi;
    const int cnt = active ? hi - f : 0;
    int cntTot;
    const int cntEx = warp_excl_scan(cnt, lane, cntTot);
    if (lane == 0) wsq[warp] = cntTot;
    __syncthreads();
    int recTot;
    const int recEx = __shfl_sync(kFull, warp_excl_scan(wsq[lane], lane, recTot), warp);
    if (sink.pool && recTot) {
      ensure_chunks(recBase + recTot);
      long long o = recBase + recEx + cntEx;
      for (int i = f; i < f + cnt; ++i, ++o) sink.put(o, make_int4(jj, __ldcg(aE + i), e, nTotal));
    }
    recBase += recTot;
    if (tid == 0 && p.qCounts) { p.qCounts[0] = (unsigned long long)recBase; p.qCounts[1] = (unsigned long long)nmBase; }
  }

  // ---- window end ----
  const bool writeEnd = !p.endOnlyLast || k1 == nTotal;
  const size_t endSlot = p.endOnlyLast ? 0 : (size_t)win;
  const int32_t *aC = aW + (size_t)cur * MS;
  if (p.aEnd && writeEnd)
    for (int i = p0 + tid; i < min(M, p0 + chunk); i += kThreads) p.aEnd[endSlot * M + i] = __ldcg(aC + i);
  if (WITH_D && p.dEnd && writeEnd) {
    const int32_t *dC = dW + (size_t)cur * (MS + 4);
    for (int i = p0 + tid; i < min(M, p0 + chunk); i += kThreads)
      p.dEnd[endSlot * (M + 1) + i] = i == 0 ? k1 + 1 : __ldcg(dC + i);
    if (g == 0 && tid == 0) p.dEnd[endSlot * (M + 1) + M] = k1 + 1;
  }
}

}  // namespace

// Plan: G CTAs per window (positions per CTA <= kMaxChunk, multiple of 1024), W windows with W*G <= maxCtas.
void sweep_global_plan(int M, int maxCtas, int wantWindows, int *G, int *chunk, int *W) {
  const int gmin = (M + kMaxChunk - 1) / kMaxChunk;
  int w = wantWindows > 0 ? wantWindows : 1;
  if (w * gmin > maxCtas) w = maxCtas / gmin;
  if (w < 1) w = 1;
  int g = maxCtas / w;  // use every SM: more CTAs per window shorten the per-site critical path
  if (g < gmin) g = gmin;
  int ch = ((M + g - 1) / g + 1023) / 1024 * 1024;
  if (ch > kMaxChunk) ch = kMaxChunk;
  g = (M + ch - 1) / ch;
  *G = g; *chunk = ch; *W = w;
}

size_t sweep_global_state_ints(int G, int chunk, int W, bool with_d) {
  const size_t MS = (size_t)G * chunk;
  return (size_t)W * 2 * MS + (with_d ? (size_t)W * 2 * (MS + 4) : 0);
}

int launch_sweep_global(const SweepArgs &a, SweepMode mode, bool with_d, int G, int chunk, int32_t *aBuf, int32_t *dBuf,
                        uint32_t *partials, void *gsync, cudaStream_t stream) {
  const size_t smemBytes = (size_t)(chunk / 32) * 4 + (with_d ? (size_t)chunk * 4 : 0);
  void *kern;
  if (mode == kGather) kern = with_d ? (void *)sweep_global_kernel<kGather, true> : (void *)sweep_global_kernel<kGather, false>;
  else kern = with_d ? (void *)sweep_global_kernel<kSorted, true> : (void *)sweep_global_kernel<kSorted, false>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
  if (err != cudaSuccess) return (int)err;
  SweepArgs args = a;
  GroupSync *gs = static_cast<GroupSync *>(gsync);
  uint32_t *ys = nullptr;
  int32_t *zp = nullptr;
  void *params[] = {&args, &G, &chunk, &aBuf, &dBuf, &partials, &gs, &ys, &zp};
  err = cudaLaunchCooperativeKernel(kern, dim3(a.nWindows * G), dim3(kThreads), params, smemBytes, stream);
  return (int)err;
}

int query_sweep_global_max_q() { return kQCap; }

int launch_query_sweep_global(const SweepArgs &a, int G, int chunk, int32_t *aBuf, int32_t *dBuf, uint32_t *partials,
                              void *gsync, uint32_t *ySorted, int32_t *zpre, cudaStream_t stream) {
  if (a.Q < 1 || a.Q > kQCap) return (int)cudaErrorInvalidValue;
  const size_t smemBytes = (size_t)(chunk / 32) * 4 * 2 + (size_t)chunk * 4 + (size_t)(4 * kQCap + kTabMirror + 96) * 4;
  void *kern = (void *)sweep_global_kernel<kSorted, true, true>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
  if (err != cudaSuccess) return (int)err;
  SweepArgs args = a;
  args.nWindows = 1;
  GroupSync *gs = static_cast<GroupSync *>(gsync);
  void *params[] = {&args, &G, &chunk, &aBuf, &dBuf, &partials, &gs, &ySorted, &zpre};
  err = cudaLaunchCooperativeKernel(kern, dim3(G), dim3(kThreads), params, smemBytes, stream);
  return (int)err;
}

size_t sweep_global_sync_bytes(int W) { return sizeof(GroupSync) * (size_t)W; }

}  // namespace pbwtgpu
