//  K5: run-length count of the sorted k-mers, and the count-value histogram.
//
//  Replaces the counting half of merylCountArray::countSingleKmers
//  (src/meryl/merylCountArray.C:334-360: one pass over the sorted suffixes, a new (kmer, count)
//  entry at every change) and merylHistogram::addValue (src/utility/src/kmers-v1/kmers-histogram.H:41-55).
//
//  count_kernel: single pass, chained-scan (decoupled look-back) stream compaction.  A tile of
//  256 x ITEMS keys marks run heads with warp ballots (key != predecessor), scans them, obtains
//  its exclusive distinct-count from the look-back chain, then writes each head's key to
//  keys[out] IN PLACE and its index to runStart[out].  In place is safe: compaction only moves
//  data towards lower indexes, a tile publishes its aggregate only after its own keys are in
//  registers, and a tile writes only after it has read every earlier aggregate.
//  count_j = runStart[j+1] - runStart[j]  (truncated to 32 bits like the reference's kmvalu).
//
//  HBM bytes per key instance: W read + (W + 8) * D/N written.  Bound: HBM bandwidth.

#include <stdlib.h>

#include "kernels.cuh"

namespace {

const int CNT_THREADS = 256;
const int CNT_WARPS   = CNT_THREADS / 32;

template <int KW> struct CountShape;
template <> struct CountShape<1> { static const int ITEMS = 16; };
template <> struct CountShape<2> { static const int ITEMS = 8;  };

template <int KW>
__global__ void __launch_bounds__(CNT_THREADS)
count_kernel(typename KeyT<KW>::type *keys, u64 nKeys, u64 *__restrict__ runStart, u64 *__restrict__ numDistinct,
             u64 *__restrict__ status, u64 *__restrict__ tileCounter) {
  typedef typename KeyT<KW>::type Key;
  const int ITEMS = CountShape<KW>::ITEMS;
  const int TILE  = CNT_THREADS * ITEMS;

  __shared__ u32 warpTot[CNT_WARPS];
  __shared__ u64 sTile, sExcl;

  const u32 t = threadIdx.x, lane = t & 31, warp = t >> 5;

  if (t == 0) sTile = atomicAdd((unsigned long long *)tileCounter, 1ull);
  __syncthreads();
  const u64 tile  = sTile;
  const u64 begin = tile * TILE;
  const u64 left  = nKeys - begin;
  const u32 n     = left > (u64)TILE ? (u32)TILE : (u32)left;

  //  warp-striped load; head flags by comparing with the predecessor in index order
  Key key[ITEMS];
  u32 heads = 0;                                             //  bit i: my item i starts a run
  const u32 wbase = warp * 32 * ITEMS;
  Key carry;                                                 //  last key of the previous row (lane 31's)
  {
    const u64 g = begin + wbase;
    if (lane == 0 && g > 0 && wbase < n) carry = keys[g - 1];
  }
  u32 before[ITEMS];                                         //  heads in this warp before my item i
  u32 wrun = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const u32 j = wbase + i * 32 + lane;
    const bool in = j < n;
    if (in) key[i] = keys[begin + j];
    //  predecessor: lane-1's key of this row, or the carry for lane 0
    Key prev;
    if (KW == 1) {
      u64 up = __shfl_up_sync(0xffffffffu, key_lo(key[i]), 1);
      *(u64 *)&prev = up;
    } else {
      u64 upl = __shfl_up_sync(0xffffffffu, key_lo(key[i]), 1);
      u64 uph = __shfl_up_sync(0xffffffffu, key_hi(key[i]), 1);
      ((u64 *)&prev)[0] = upl;  ((u64 *)&prev)[1] = uph;
    }
    if (lane == 0) prev = carry;
    const bool first = (begin + j == 0);
    const bool head  = in && (first || !key_eq(key[i], prev));
    const u32  b     = __ballot_sync(0xffffffffu, head);
    before[i] = wrun + __popc(b & ((1u << lane) - 1));
    wrun += __popc(b);
    if (head) heads |= 1u << i;
    //  carry for the next row = this row's lane 31
    if (KW == 1) {
      u64 c = __shfl_sync(0xffffffffu, key_lo(key[i]), 31);
      *(u64 *)&carry = c;
    } else {
      u64 cl = __shfl_sync(0xffffffffu, key_lo(key[i]), 31);
      u64 ch = __shfl_sync(0xffffffffu, key_hi(key[i]), 31);
      ((u64 *)&carry)[0] = cl;  ((u64 *)&carry)[1] = ch;
    }
  }

  if (lane == 0) warpTot[warp] = wrun;
  __syncthreads();                                           //  every key of the tile is in registers now

  u32 wOff = 0, tileTot = 0;
#pragma unroll
  for (int w = 0; w < CNT_WARPS; w++) { const u32 c = warpTot[w]; if (w < (int)warp) wOff += c; tileTot += c; }

  //  look back (warp 0)
  if (warp == 0) {
    if (lane == 0) st_volatile_u64(status + tile, (tile == 0 ? MB_ST_INC : MB_ST_AGG) | tileTot);
    const u64 excl = lookback_warp(status, tile, lane);
    if (lane == 0) {
      if (tile != 0) st_volatile_u64(status + tile, MB_ST_INC | (excl + tileTot));
      sExcl = excl;
      if (begin + n == nKeys) {                              //  last tile: totals and the sentinel
        *numDistinct = excl + tileTot;
        runStart[excl + tileTot] = nKeys;
      }
    }
  }
  __syncthreads();

  const u64 outBase = sExcl + wOff;
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    if (heads & (1u << i)) {
      const u64 o = outBase + before[i];
      keys[o]     = key[i];
      runStart[o] = begin + wbase + i * 32 + lane;
    }
  }
}

//  ---- u32 suffix keys: segmented run-length count with expansion to full keys ----------------------
//
//  The sorted pass array holds only the low suffixBits of every k-mer; the prefix bin is the
//  segment.  Tiles are the sort's tiles (TileInfo: begin, n, segment, first-of-segment), so a tile
//  never straddles two bins and the first key of a segment's first tile always starts a run.  The
//  kernel walks the tiles [tileLo, tileHi) (whole segments), writes every distinct key expanded to
//  u64 = (binLo + segment) << suffixBits | suffix, and the pass-array index of its first instance.
//  Not in place (u64 out of u32 in); the look-back chain starts at tileLo.

struct __align__(16) CountTileInfo { u64 begin; u32 n; u32 seg; };       //  == sort.cu's TileInfo

const int C32_ITEMS  = 16;                                               //  a CTA counts HALF a sort tile (SortShape<0>::ITEMS == 32):
const int C32_HALF   = CNT_THREADS * C32_ITEMS;                          //  16 keys per thread keep it at 64 registers, 4 CTAs per SM

__global__ void __launch_bounds__(CNT_THREADS, 4)
count_suffix32_kernel(const u32 *__restrict__ keys, const CountTileInfo *__restrict__ info, u64 tileLo, u64 nHalves,
                      u64 endIndex, u32 binLo, u32 suffixBits, u64 *__restrict__ uniq, u64 *__restrict__ runStart,
                      u64 *__restrict__ numDistinct, u64 *__restrict__ status, u64 *__restrict__ tileCounter, u32 pfDist) {
  __shared__ u32 warpTot[CNT_WARPS];
  __shared__ u64 sTile, sExcl;

  const u32 t = threadIdx.x, lane = t & 31, warp = t >> 5;

  if (t == 0) sTile = atomicAdd((unsigned long long *)tileCounter, 1ull);
  __syncthreads();
  const u64 rel  = sTile;                                    //  half-tile index inside the group
  const u32 half = (u32)(rel & 1);
  const CountTileInfo ti = info[tileLo + (rel >> 1)];
  const u64  begin   = ti.begin + (u64)half * C32_HALF;
  const u32  n       = ti.n > half * C32_HALF ? min(ti.n - half * C32_HALF, (u32)C32_HALF) : 0;
  const u32  seg     = ti.seg & 0x7fffffffu;
  const bool segHead = (ti.seg >> 31) != 0 && half == 0;     //  the segment's very first key always starts a run

  //  the keys of the half-tile claimed pfDist claims from now: HBM -> L2 (the kernel is a chain of dependent round trips
  //  per 16 KB of keys -- claim, descriptor, keys, look-back -- so the DRAM part of the key load is taken out of it)
  if (pfDist && rel + pfDist < nHalves && t < 128) {
    const u64 rn = rel + pfDist;
    const CountTileInfo tn = info[tileLo + (rn >> 1)];
    const u32 hn = (u32)(rn & 1);
    if (tn.n > hn * C32_HALF) {
      const u32 nn = min(tn.n - hn * C32_HALF, (u32)C32_HALF);
      if (t * 32 < nn) prefetch_l2(keys + tn.begin + (u64)hn * C32_HALF + t * 32);
    }
  }

  u32 key[C32_ITEMS];
  u32 heads = 0;
  const u32 wbase = warp * 32 * C32_ITEMS;
  u32 carry = 0;
  if (lane == 0 && wbase < n && !(segHead && wbase == 0)) carry = keys[begin + wbase - 1];
  unsigned short before[C32_ITEMS];
  u32 wrun = 0;
#pragma unroll
  for (int i = 0; i < C32_ITEMS; i++) {
    const u32 j = wbase + i * 32 + lane;
    const bool in = j < n;
    key[i] = in ? keys[begin + j] : 0;
    u32 prev = __shfl_up_sync(0xffffffffu, key[i], 1);
    if (lane == 0) prev = carry;
    const bool first = segHead && (j == 0);
    const bool head  = in && (first || key[i] != prev);
    const u32  b     = __ballot_sync(0xffffffffu, head);
    before[i] = (unsigned short)(wrun + __popc(b & ((1u << lane) - 1)));
    wrun += __popc(b);
    if (head) heads |= 1u << i;
    carry = __shfl_sync(0xffffffffu, key[i], 31);
  }

  if (lane == 0) warpTot[warp] = wrun;
  __syncthreads();

  u32 wOff = 0, tileTot = 0;
#pragma unroll
  for (int w = 0; w < CNT_WARPS; w++) { const u32 c = warpTot[w]; if (w < (int)warp) wOff += c; tileTot += c; }

  if (warp == 0) {
    if (lane == 0) st_volatile_u64(status + rel, (rel == 0 ? MB_ST_INC : MB_ST_AGG) | tileTot);
    const u64 excl = lookback_warp(status, rel, lane);
    if (lane == 0) {
      if (rel != 0) st_volatile_u64(status + rel, MB_ST_INC | (excl + tileTot));
      sExcl = excl;
      if (rel + 1 == nHalves) {                              //  last half-tile of the group: total and sentinel
        *numDistinct = excl + tileTot;
        runStart[excl + tileTot] = endIndex;
      }
    }
  }
  __syncthreads();

  const u64 outBase = sExcl + wOff;
  const u64 hiBits  = (u64)(binLo + seg) << suffixBits;
#pragma unroll
  for (int i = 0; i < C32_ITEMS; i++) {
    if (heads & (1u << i)) {
      const u64 o = outBase + before[i];
      uniq[o]     = hiBits | key[i];
      runStart[o] = begin + wbase + i * 32 + lane;
    }
  }
}

//  Grouped form (default; MB_COUNT_GROUP=1 keeps the kernel above for A/B).  ncu of the kernel above: 7.7 % of the issue slots
//  busy, 53 % of the stall samples at the barrier behind the look-back and 26 % on the key loads -- a CTA is a chain of
//  dependent round trips (claim, descriptor, keys, look-back, stores) for 16 KB of keys.  Here a CTA takes G consecutive
//  half-tiles (64 KB of keys for G = 4) with ONE claim and ONE look-back: a first sweep over its keys counts the run heads
//  per (half-tile, warp), the look-back resolves the CTA's exclusive distinct count, a second sweep -- the keys come from
//  L2 now, every resident CTA's 64 KB together are far below its 126 MB -- writes the heads.  Twice the instructions per
//  key (still far below what the memory system can feed), a quarter of the chains.
template <int G, bool WRITE>
__device__ __forceinline__ u32
count32_half(const u32 *__restrict__ keys, const CountTileInfo &ti, u32 half, u32 lane, u32 warp,
             u64 outBase, u32 binLo, u32 suffixBits, u64 *__restrict__ uniq, u64 *__restrict__ runStart) {
  const u64  begin   = ti.begin + (u64)half * C32_HALF;
  const u32  n       = ti.n > half * C32_HALF ? min(ti.n - half * C32_HALF, (u32)C32_HALF) : 0;
  const bool segHead = (ti.seg >> 31) != 0 && half == 0;     //  the segment's very first key always starts a run
  const u32  wbase   = warp * 32 * C32_ITEMS;
  if (wbase >= n) return 0;                                   //  (warp-uniform)
  u32 key[C32_ITEMS];
#pragma unroll
  for (int i = 0; i < C32_ITEMS; i++) {
    const u32 j = wbase + i * 32 + lane;
    key[i] = j < n ? keys[begin + j] : 0;
  }
  u32 carry = 0;
  if (lane == 0 && !(segHead && wbase == 0)) carry = keys[begin + wbase - 1];
  const u64 hiBits = (u64)(binLo + (ti.seg & 0x7fffffffu)) << suffixBits;     //  (binLo may have wrapped below zero: 32-bit sum)
  u32 wrun = 0;
#pragma unroll
  for (int i = 0; i < C32_ITEMS; i++) {
    const u32 j = wbase + i * 32 + lane;
    u32 prev = __shfl_up_sync(0xffffffffu, key[i], 1);
    if (lane == 0) prev = carry;
    const bool head = j < n && ((segHead && j == 0) || key[i] != prev);
    const u32  b    = __ballot_sync(0xffffffffu, head);
    if (WRITE && head) {
      const u64 o = outBase + wrun + __popc(b & ((1u << lane) - 1));
      uniq[o]     = hiBits | key[i];
      runStart[o] = begin + j;
    }
    wrun += __popc(b);
    carry = __shfl_sync(0xffffffffu, key[i], 31);
  }
  return wrun;
}

template <int G>
__global__ void __launch_bounds__(CNT_THREADS, 4)
count_suffix32g_kernel(const u32 *__restrict__ keys, const CountTileInfo *__restrict__ info, u64 tileLo, u64 nHalves,
                       u64 endIndex, u32 binLo, u32 suffixBits, u64 *__restrict__ uniq, u64 *__restrict__ runStart,
                       u64 *__restrict__ numDistinct, u64 *__restrict__ status, u64 *__restrict__ tileCounter) {
  __shared__ u32 warpTot[G][CNT_WARPS];
  __shared__ u64 sTile, sExcl;
  const u32 t = threadIdx.x, lane = t & 31, warp = t >> 5;

  if (t == 0) sTile = atomicAdd((unsigned long long *)tileCounter, 1ull);
  __syncthreads();
  const u64 rel = sTile;                                     //  group of G half-tiles, in claim order
  const u64 nGroups = (nHalves + G - 1) / G;
  const u64 h0 = rel * G;
  const u32 nH = (u32)min((u64)G, nHalves - h0);

  CountTileInfo ti[G];
#pragma unroll
  for (int g = 0; g < G; g++) if ((u32)g < nH) ti[g] = info[tileLo + ((h0 + g) >> 1)];

  //  first sweep: run heads per (half-tile, warp)
#pragma unroll
  for (int g = 0; g < G; g++) {
    u32 c = 0;
    if ((u32)g < nH) c = count32_half<G, false>(keys, ti[g], (u32)((h0 + g) & 1), lane, warp, 0, 0, suffixBits, uniq, runStart);
    if (lane == 0) warpTot[g][warp] = c;
  }
  __syncthreads();

  u32 myOff[G], tot = 0;                                      //  heads before my warp's share of half-tile g, inside this CTA
#pragma unroll
  for (int g = 0; g < G; g++) {
#pragma unroll
    for (int w = 0; w < CNT_WARPS; w++) { if (w == (int)warp) myOff[g] = tot; tot += warpTot[g][w]; }
  }

  if (warp == 0) {
    if (lane == 0) st_volatile_u64(status + rel, (rel == 0 ? MB_ST_INC : MB_ST_AGG) | tot);
    const u64 excl = lookback_warp(status, rel, lane);
    if (lane == 0) {
      if (rel != 0) st_volatile_u64(status + rel, MB_ST_INC | (excl + tot));
      sExcl = excl;
      if (rel + 1 == nGroups) {                              //  last group: total and sentinel
        *numDistinct = excl + tot;
        runStart[excl + tot] = endIndex;
      }
    }
  }
  __syncthreads();

  //  second sweep: the same walk, writing
#pragma unroll
  for (int g = 0; g < G; g++)
    if ((u32)g < nH) count32_half<G, true>(keys, ti[g], (u32)((h0 + g) & 1), lane, warp, sExcl + myOff[g], binLo, suffixBits, uniq, runStart);
}

//  ---- value histogram --------------------------------------------------------------------------

const int VH_THREADS = 256;
const int VH_SMEM    = 1024;

__global__ void __launch_bounds__(VH_THREADS)
value_hist_kernel(const u64 *__restrict__ runStart, const u32 *__restrict__ counts, const u64 *__restrict__ numDistinct,
                  u64 *__restrict__ histSmall, u32 *__restrict__ histBig, u64 *__restrict__ histBigN, u64 histBigCap,
                  u64 *__restrict__ totals) {
  __shared__ u32 sh[VH_SMEM];
  __shared__ u64 sTot[3];
  const u32 t = threadIdx.x;
  for (u32 i = t; i < VH_SMEM; i += VH_THREADS) sh[i] = 0;
  if (t < 3) sTot[t] = 0;
  __syncthreads();

  const u64 D = *numDistinct;
  u64 uniq = 0, dist = 0, tot = 0;
  for (u64 j = (u64)blockIdx.x * VH_THREADS + t; j < D; j += (u64)gridDim.x * VH_THREADS) {
    const u32 c = counts ? counts[j] : (u32)(runStart[j + 1] - runStart[j]);      //  kmvalu is 32 bits and wraps (merylCountArray.C:357)
    if (c == 0) continue;                                    //  addValue ignores zero
    uniq += (c == 1);
    dist += 1;
    tot  += c;
    if (c < VH_SMEM)             atomicAdd(&sh[c], 1u);
    else if (c < MB_HIST_SMALL)  atomicAdd((unsigned long long *)&histSmall[c], 1ull);
    else {
      const u64 slot = atomicAdd((unsigned long long *)histBigN, 1ull);
      if (slot < histBigCap) histBig[slot] = c;
    }
  }

  //  block reduction of the three totals
  for (int o = 16; o > 0; o >>= 1) {
    uniq += __shfl_down_sync(0xffffffffu, uniq, o);
    dist += __shfl_down_sync(0xffffffffu, dist, o);
    tot  += __shfl_down_sync(0xffffffffu, tot,  o);
  }
  if ((t & 31) == 0) {
    atomicAdd((unsigned long long *)&sTot[0], (unsigned long long)uniq);
    atomicAdd((unsigned long long *)&sTot[1], (unsigned long long)dist);
    atomicAdd((unsigned long long *)&sTot[2], (unsigned long long)tot);
  }
  __syncthreads();
  if (t < 3 && sTot[t]) atomicAdd((unsigned long long *)&totals[t], (unsigned long long)sTot[t]);
  for (u32 i = t; i < VH_SMEM; i += VH_THREADS)
    if (sh[i]) atomicAdd((unsigned long long *)&histSmall[i], (unsigned long long)sh[i]);
}

//  test hook: one warp walks a prepared status array
__global__ void __launch_bounds__(32)
lookback_test_kernel(const u64 *__restrict__ status, u64 tile, u64 *__restrict__ out) {
  const u64 e = lookback_warp(status, tile, threadIdx.x);
  if (threadIdx.x == 0) *out = e;
}

}  //  namespace


cudaError_t mb_launch_lookback_test(const u64 *dStatus, u64 tile, u64 *dOut, cudaStream_t s) {
  lookback_test_kernel<<<1, 32, 0, s>>>(dStatus, tile, dOut);
  return cudaGetLastError();
}

static u64 count_tiles(u32 keyWords, u64 nKeys) {
  const u64 tile = CNT_THREADS * (keyWords == 1 ? CountShape<1>::ITEMS : CountShape<2>::ITEMS);
  return (nKeys + tile - 1) / tile;
}

//  workspace: [tile counter: 1 u64 (padded to 2)][status: nTiles u64]
size_t mb_count_workspace_bytes(u64 nKeys) {
  return (2 + count_tiles(1, nKeys) + count_tiles(2, nKeys)) * sizeof(u64);
}

cudaError_t mb_launch_count(u32 keyWords, void *dKeys, u64 nKeys, u64 *dRunStart, u64 *dNumDistinct,
                            void *dWorkspace, cudaStream_t s) {
  cudaError_t e;
  if (nKeys == 0) {
    if ((e = cudaMemsetAsync(dNumDistinct, 0, sizeof(u64), s)) != cudaSuccess) return e;
    return cudaMemsetAsync(dRunStart, 0, sizeof(u64), s);
  }
  const u64 nTiles = count_tiles(keyWords, nKeys);
  if ((e = cudaMemsetAsync(dWorkspace, 0, (2 + nTiles) * sizeof(u64), s)) != cudaSuccess) return e;
  u64 *counter = (u64 *)dWorkspace, *status = counter + 2;
  if (keyWords == 1) count_kernel<1><<<(unsigned)nTiles, CNT_THREADS, 0, s>>>((u64 *)dKeys, nKeys, dRunStart, dNumDistinct, status, counter);
  else               count_kernel<2><<<(unsigned)nTiles, CNT_THREADS, 0, s>>>((ulonglong2 *)dKeys, nKeys, dRunStart, dNumDistinct, status, counter);
  return cudaGetLastError();
}

cudaError_t mb_launch_count_suffix32(const u32 *dKeys, const void *dTileInfo, u64 tileLo, u64 tileHi, u64 endIndex,
                                     u32 binLo, u32 suffixBits, u64 *dUniq, u64 *dRunStart, u64 *dNumDistinct,
                                     void *dWorkspace, cudaStream_t s) {
  cudaError_t e;
  if (tileHi <= tileLo) {                                    //  empty group: D = 0, the sentinel is never read
    if ((e = cudaMemsetAsync(dNumDistinct, 0, sizeof(u64), s)) != cudaSuccess) return e;
    return cudaMemsetAsync(dRunStart, 0, sizeof(u64), s);
  }
  const u64 nHalves = 2 * (tileHi - tileLo);                 //  workspace: (2 + nHalves) u64
  if ((e = cudaMemsetAsync(dWorkspace, 0, (2 + nHalves) * sizeof(u64), s)) != cudaSuccess) return e;
  u64 *counter = (u64 *)dWorkspace, *status = counter + 2;
  //  (A blocked-key form of this kernel, transposing through shared memory, measured 40 % slower in round 1 and is gone.)
  //  MB_COUNT_GROUP=1: the one-half-tile-per-CTA kernel, for A/B (measured on B200, 6.9 G keys: 35.4 ms; two half-tiles per
  //  CTA 28.1 ms; four 24.5 ms -- the default).
  const bool grouped = !(getenv("MB_COUNT_GROUP") && getenv("MB_COUNT_GROUP")[0] == '1');
  const CountTileInfo *ti = (const CountTileInfo *)dTileInfo;
  if (grouped) count_suffix32g_kernel<4><<<(unsigned)((nHalves + 3) / 4), CNT_THREADS, 0, s>>>(dKeys, ti, tileLo, nHalves, endIndex, binLo, suffixBits, dUniq, dRunStart, dNumDistinct, status, counter);
  else
    count_suffix32_kernel<<<(unsigned)nHalves, CNT_THREADS, 0, s>>>(dKeys, ti, tileLo, nHalves, endIndex, binLo, suffixBits, dUniq, dRunStart, dNumDistinct, status, counter,
                                                                  getenv("MB_COUNT_PREFETCH") ? (u32)strtoul(getenv("MB_COUNT_PREFETCH"), nullptr, 10) : 0u);
  return cudaGetLastError();
}

cudaError_t mb_launch_value_histogram(const u64 *dRunStart, const u64 *dNumDistinct, u64 nMaxDistinct,
                                      u64 *dHistSmall, u32 *dHistBig, u64 *dHistBigN, u64 histBigCap,
                                      u64 *dTotals, cudaStream_t s) {
  if (nMaxDistinct == 0) return cudaSuccess;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  u64 grid = (nMaxDistinct + VH_THREADS - 1) / VH_THREADS;
  const u64 cap = (u64)sms * 8;
  if (grid > cap) grid = cap;
  value_hist_kernel<<<(unsigned)grid, VH_THREADS, 0, s>>>(dRunStart, nullptr, dNumDistinct, dHistSmall, dHistBig, dHistBigN, histBigCap, dTotals);
  return cudaGetLastError();
}

cudaError_t mb_launch_value_histogram_counts(const u32 *dCounts, const u64 *dNumDistinct, u64 nMaxDistinct,
                                             u64 *dHistSmall, u32 *dHistBig, u64 *dHistBigN, u64 histBigCap,
                                             u64 *dTotals, cudaStream_t s) {
  if (nMaxDistinct == 0) return cudaSuccess;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  u64 grid = (nMaxDistinct + VH_THREADS - 1) / VH_THREADS;
  const u64 cap = (u64)sms * 8;
  if (grid > cap) grid = cap;
  value_hist_kernel<<<(unsigned)grid, VH_THREADS, 0, s>>>(nullptr, dCounts, dNumDistinct, dHistSmall, dHistBig, dHistBigN, histBigCap, dTotals);
  return cudaGetLastError();
}
