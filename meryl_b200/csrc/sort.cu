//  K4: segmented LSD radix sort of packed k-mers ("onesweep": one read + one write of every key
//  per digit, decoupled look-back between tiles).
//
//  Replaces std::sort of each prefix bucket (merylCountArray::countSingleKmers,
//  src/meryl/merylCountArray.C:324-330).  The reference sorts 2^wPrefix small buckets on the CPU;
//  here one launch per digit sorts ALL prefix-bin segments of a pass at once: a segment is the
//  run of keys sharing their top bits (put there by the extraction scatter), and only the
//  remaining low bits are sorted, MB_RADIX_BITS (6) or fewer bits per pass.
//
//   sort_tiles_kernel   one 16-byte descriptor per tile (first key, count, segment, first-of-segment)
//   sort_hist_kernel    one read of the keys: per (segment, pass, digit) counts
//   sort_scan_kernel    exclusive scan over digits -> first output index of every (segment, pass, digit)
//   sort_sweep_kernel   per pass: tile = 256 threads x ITEMS keys
//       1. coalesced (warp-striped) load, transposed through shared memory so that every thread
//          owns ITEMS CONSECUTIVE keys -- rank order == memory order, which LSD needs (stable);
//       2. rank by PRIVATE counters: thread t bumps its own 16-bit counter of the key's digit
//          (two digits per 32-bit word, no atomics, no warp votes), a raking scan over the
//          32 x 256 packed counters turns them into tile-wide positions -- ~25 instructions per 32
//          keys, against ~95 for a match/ballot multisplit, which is what lets the kernel approach
//          the HBM roofline on B200 (0.17 warp-instructions per byte at peak bandwidth);
//       3. the tile publishes its 64 digit counts, and resolves its exclusive prefix by a batched
//          decoupled look-back over the earlier tiles OF ITS SEGMENT (status words carry a pass
//          epoch, so the table is cleared once per sort);
//       4. keys are placed digit-sorted in shared memory and leave as coalesced digit runs.
//
//  Algorithmic HBM bytes per key: hist 1*W, each sweep pass 2*W (W = 4, 8 or 16 bytes), plus the
//  tile status (512 B per tile written twice, read >= once).  Bound: HBM bandwidth.
//
//  Two sweep kernels share the tile table, the histogram and the scan (mb_sort_version()):
//   v1  sort_sweep_kernel    256 threads x 32/16/8 keys, 6-bit digits  (32 suffix bits = 6 passes)
//   v2-4 sortx_sweep_kernel  second generation, shaped by (threads, digit bits): the first one is issue-bound, not
//       DRAM-bound (ncu: 63 % of the issue slots busy at 21 % DRAM), so the lever is instructions per key: 16-bit
//       counter accesses, one PRMT per whole-byte digit, 32-bit output offsets -- about half the instructions per
//       key and pass.  Measured on B200 (profiles/): the 8-bit shape (4 passes for 32 bits, but 64 KB of counters
//       = 8 warps per SM) is latency-bound at 17 % issue utilisation and no faster than v1 over the whole sort.

#include <stdlib.h>
#include <string.h>
#include <algorithm>

#include "kernels.cuh"

namespace {

const int SORT_THREADS = 256;
const int SORT_WARPS   = SORT_THREADS / 32;
const int SORT_MAX_PASSES = 24;
#ifndef MB_SWEEP_PREFETCH_DEFAULT
#define MB_SWEEP_PREFETCH_DEFAULT 0      //  tiles ahead; measured A/B decides (MB_SWEEP_PREFETCH)
#endif

template <int KW> struct SortShape;
//  MIN_CTAS: resident CTAs per SM the register budget is compiled for; ~68 KB of shared memory per CTA
//  (counters + key buffer) allow three.
template <> struct SortShape<0> { static const int ITEMS = 32; static const int MIN_CTAS = 3; };   //  u32 suffix keys: 8192 per tile
template <> struct SortShape<1> { static const int ITEMS = 16; static const int MIN_CTAS = 3; };   //  u64: 4096 per tile
template <> struct SortShape<2> { static const int ITEMS = 8;  static const int MIN_CTAS = 3; };   //  128 bit: 2048 per tile

//  status word: [63:62] flag (0 none, 1 aggregate, 2 inclusive)  [61:57] epoch (pass tag, 1..24)  [56:0] value
#define ST_FLAG(v)    ((u32)((v) >> 62))
#define ST_EPOCH(v)   ((u32)(((v) >> 57) & 31))
#define ST_VALUE(v)   ((v) & ((1ull << 57) - 1))
#define ST_MAKE(f, e, x)  (((u64)(f) << 62) | ((u64)(e) << 57) | (u64)(x))

//  Largest s with tileStart[s] <= tile  (tileStart has nSegs+1 entries, non-decreasing).
__device__ __forceinline__ u32 find_segment(const u64 *__restrict__ tileStart, u32 nSegs, u64 tile) {
  u32 lo = 0, hi = nSegs;                    //  answer in [lo, hi)
  while (hi - lo > 1) {
    const u32 mid = (lo + hi) >> 1;
    if (tileStart[mid] <= tile) lo = mid; else hi = mid;
  }
  return lo;
}

struct SortPassTable { u32 nPasses; u32 shift[SORT_MAX_PASSES]; u32 bits[SORT_MAX_PASSES]; };

//  ---- tile descriptors: one 16-byte record per tile, so a sweep CTA starts with ONE dependent load --

struct __align__(16) TileInfo {
  u64 begin;                                 //  index of the tile's first key
  u32 n;                                     //  keys in the tile
  u32 seg;                                   //  segment; bit 31: first tile of its segment
};

__global__ void __launch_bounds__(256)
sort_tiles_kernel(const u64 *__restrict__ segStart, const u64 *__restrict__ segTile, u32 nSegs, u64 nTiles,
                  u32 tileKeys, TileInfo *__restrict__ info) {
  const u64 tile = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (tile >= nTiles) return;
  const u32 seg   = find_segment(segTile, nSegs, tile);
  const u64 first = segTile[seg];
  const u64 begin = segStart[seg] + (tile - first) * tileKeys;
  u64 n = segStart[seg + 1] - begin;
  if (n > tileKeys) n = tileKeys;
  TileInfo t;
  t.begin = begin;  t.n = (u32)n;  t.seg = seg | (tile == first ? 0x80000000u : 0u);
  info[tile] = t;
}

//  ---- histogram ------------------------------------------------------------------------------

const int HIST_TILES_PER_CTA = 16;
const int HIST_WORDS = 4096;                          //  24 passes x 64 digits, or 16 passes x 256 digits

template <int KW>
__global__ void __launch_bounds__(SORT_THREADS)
sort_hist_kernel(const typename KeyT<KW>::type *__restrict__ keys,
                 const u64 *__restrict__ segStart, const u64 *__restrict__ segTile, u32 nSegs, u64 nTiles,
                 SortPassTable pt, u32 radix, u64 *__restrict__ ghist, u32 unrolled) {
  typedef typename KeyT<KW>::type Key;
  const int TILE = SORT_THREADS * SortShape<KW>::ITEMS;
  const int HIST_UNROLL = 8;
  __shared__ u32 sh[HIST_WORDS];                      //  nPasses * radix <= HIST_WORDS (checked by the host)

  const u32 t = threadIdx.x;
  for (u32 i = t; i < pt.nPasses * radix; i += SORT_THREADS) sh[i] = 0;
  __syncthreads();

  const u64 tile0 = (u64)blockIdx.x * HIST_TILES_PER_CTA;
  u64 tileEnd = tile0 + HIST_TILES_PER_CTA;
  if (tileEnd > nTiles) tileEnd = nTiles;

  u32 seg = find_segment(segTile, nSegs, tile0);

  for (u64 tile = tile0; tile < tileEnd; tile++) {
    //  advance to the segment of this tile, flushing when it changes
    u32 s = seg;
    while (segTile[s + 1] <= tile) s++;
    if (s != seg) {
      __syncthreads();
      for (u32 i = t; i < pt.nPasses * radix; i += SORT_THREADS) {
        const u32 c = sh[i];
        if (c) { atomicAdd((unsigned long long *)&ghist[(u64)seg * pt.nPasses * radix + i], (unsigned long long)c); sh[i] = 0; }
      }
      __syncthreads();
      seg = s;
    }

    const u64 begin = segStart[seg] + (tile - segTile[seg]) * TILE;
    u64 n = segStart[seg + 1] - begin;
    if (n > (u64)TILE) n = TILE;

    //  HIST_UNROLL independent loads per thread are in flight before the first shared-memory atomic (one load at a time
    //  left the kernel waiting on HBM latency: 1.3 TB/s measured).  MB_HIST_SIMPLE=1 keeps the one-load loop for A/B.
    if (unrolled) {
      const Key *kp = keys + begin;
      u32 j = t;
      for (; j + (HIST_UNROLL - 1) * SORT_THREADS < (u32)n; j += HIST_UNROLL * SORT_THREADS) {
        Key k[HIST_UNROLL];
#pragma unroll
        for (int u = 0; u < HIST_UNROLL; u++) k[u] = kp[j + u * SORT_THREADS];
#pragma unroll
        for (int u = 0; u < HIST_UNROLL; u++)
          for (u32 p = 0; p < pt.nPasses; p++)
            atomicAdd(&sh[p * radix + key_field(k[u], pt.shift[p], pt.bits[p])], 1u);
      }
      for (; j < (u32)n; j += SORT_THREADS) {
        const Key key = kp[j];
        for (u32 p = 0; p < pt.nPasses; p++)
          atomicAdd(&sh[p * radix + key_field(key, pt.shift[p], pt.bits[p])], 1u);
      }
    } else
    for (u32 j = t; j < (u32)n; j += SORT_THREADS) {
      const Key key = keys[begin + j];
      for (u32 p = 0; p < pt.nPasses; p++)
        atomicAdd(&sh[p * radix + key_field(key, pt.shift[p], pt.bits[p])], 1u);
    }
  }

  __syncthreads();
  for (u32 i = t; i < pt.nPasses * radix; i += SORT_THREADS) {
    const u32 c = sh[i];
    if (c) atomicAdd((unsigned long long *)&ghist[(u64)seg * pt.nPasses * radix + i], (unsigned long long)c);
  }
}

//  Joint form (MB_HIST_JOINT=1; not the default: measured equal): adjacent passes are histogrammed together (one
//  shared-memory atomic per PAIR of digits into a 2^(b1+b2)-entry table) and split into the two marginal histograms
//  only when a CTA flushes.
struct SortHistGroups {
  u32 nGroups;
  u32 shift[SORT_MAX_PASSES];          //  of the group's first pass
  u32 bitsLo[SORT_MAX_PASSES];         //  digit width of the first pass
  u32 bitsAll[SORT_MAX_PASSES];        //  of the whole group (== bitsLo for a single pass)
  u32 firstPass[SORT_MAX_PASSES];
  u32 offset[SORT_MAX_PASSES];         //  of the group's table in shared memory (entries)
  u32 total;                           //  entries of all tables
};

template <int KW, int NG>
__global__ void __launch_bounds__(SORT_THREADS)
sort_hist2_kernel(const typename KeyT<KW>::type *__restrict__ keys,
                  const u64 *__restrict__ segStart, const u64 *__restrict__ segTile, u32 nSegs, u64 nTiles,
                  SortHistGroups hg, u32 nPasses, u32 radix, u64 *__restrict__ ghist) {
  typedef typename KeyT<KW>::type Key;
  const int TILE = SORT_THREADS * SortShape<KW>::ITEMS;
  const int UNROLL = 8;
  extern __shared__ __align__(16) unsigned char histSmem[];
  u32 *jt = (u32 *)histSmem;                           //  hg.total joint counters
  __shared__ u32 marg[HIST_WORDS];                     //  nPasses * radix marginal counters

  const u32 t = threadIdx.x;
  for (u32 i = t; i < hg.total; i += SORT_THREADS) jt[i] = 0;
  __syncthreads();

  //  the groups' fields, once, in registers (NG is a compile-time constant: the loops below unroll and the hot loop is
  //  shift / mask / add / one shared-memory atomic per group -- ncu of the first form: 83 instructions per key, a
  //  seventh of them constant-bank loads of these very numbers, and the atomics' queue (MIO) the top stall)
  u32 gShift[NG], gBits[NG];  u32 *gTab[NG];
#pragma unroll
  for (int g = 0; g < NG; g++) { gShift[g] = hg.shift[g];  gBits[g] = hg.bitsAll[g];  gTab[g] = jt + hg.offset[g]; }

  const u64 tile0 = (u64)blockIdx.x * HIST_TILES_PER_CTA;
  u64 tileEnd = tile0 + HIST_TILES_PER_CTA;
  if (tileEnd > nTiles) tileEnd = nTiles;
  u32 seg = find_segment(segTile, nSegs, tile0);

  auto flush = [&](u32 sg) {                           //  called by all threads
    __syncthreads();
    for (u32 i = t; i < nPasses * radix; i += SORT_THREADS) marg[i] = 0;
    __syncthreads();
    for (u32 g = 0; g < hg.nGroups; g++) {
      const u32 n = 1u << hg.bitsAll[g], lowMask = (1u << hg.bitsLo[g]) - 1;
      const bool pair = hg.bitsAll[g] != hg.bitsLo[g];
      u32 *tab = jt + hg.offset[g];
      for (u32 e = t; e < n; e += SORT_THREADS) {
        const u32 c = tab[e];
        if (c) {
          atomicAdd(&marg[hg.firstPass[g] * radix + (e & lowMask)], c);
          if (pair) atomicAdd(&marg[(hg.firstPass[g] + 1) * radix + (e >> hg.bitsLo[g])], c);
          tab[e] = 0;
        }
      }
    }
    __syncthreads();
    for (u32 i = t; i < nPasses * radix; i += SORT_THREADS) {
      const u32 c = marg[i];
      if (c) atomicAdd((unsigned long long *)&ghist[(u64)sg * nPasses * radix + i], (unsigned long long)c);
    }
  };

  for (u64 tile = tile0; tile < tileEnd; tile++) {
    u32 sNow = seg;
    while (segTile[sNow + 1] <= tile) sNow++;
    if (sNow != seg) { flush(seg); __syncthreads(); seg = sNow; }

    const u64 begin = segStart[seg] + (tile - segTile[seg]) * TILE;
    u64 n = segStart[seg + 1] - begin;
    if (n > (u64)TILE) n = TILE;
    const Key *kp = keys + begin;
    u32 j = t;
    for (; j + (UNROLL - 1) * SORT_THREADS < (u32)n; j += UNROLL * SORT_THREADS) {     //  UNROLL independent loads in flight
      Key k[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; u++) k[u] = kp[j + u * SORT_THREADS];
#pragma unroll
      for (int u = 0; u < UNROLL; u++)
#pragma unroll
        for (int g = 0; g < NG; g++) atomicAdd(gTab[g] + key_field(k[u], gShift[g], gBits[g]), 1u);
    }
    for (; j < (u32)n; j += SORT_THREADS) {
      const Key key = kp[j];
#pragma unroll
      for (int g = 0; g < NG; g++) atomicAdd(gTab[g] + key_field(key, gShift[g], gBits[g]), 1u);
    }
  }
  flush(seg);
}

//  ---- scan: counts -> first output index of each (segment, pass, digit) -------------------------

__global__ void __launch_bounds__(256)
sort_scan_kernel(u64 *__restrict__ ghist, const u64 *__restrict__ segStart, u32 nPasses) {
  __shared__ u64 warpSum[8];
  const u32 seg = blockIdx.x, pass = blockIdx.y, d = threadIdx.x;       //  blockDim.x == radix (64 or 256)
  u64 *h = ghist + ((u64)seg * nPasses + pass) * blockDim.x;

  const u64 c = h[d];
  u64 incl = c;
  for (int o = 1; o < 32; o <<= 1) {
    const u64 v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((d & 31) >= (u32)o) incl += v;
  }
  if ((d & 31) == 31) warpSum[d >> 5] = incl;
  __syncthreads();
  u64 base = 0;
  for (u32 w = 0; w < (d >> 5); w++) base += warpSum[w];
  h[d] = segStart[seg] + base + incl - c;
}

//  ---- one digit pass ---------------------------------------------------------------------------

//  Dynamic shared memory: the private digit counters, then a key buffer used first as the
//  striped->blocked transpose area and later as the digit-sorted stage.  Sizes in bytes, for TILE keys of KB bytes:
//    transpose / stage : TILE*KB, plus one key of padding per 128 bytes (bank-conflict-free blocked reads)
//    counters          : (MB_RADIX/2) lanes x 256 threads x 4 B = 32 KB, plus one word of padding per 32
template <int KW> struct SweepSmem {
  typedef typename KeyT<KW>::type Key;
  static const int TILE     = SORT_THREADS * SortShape<KW>::ITEMS;
  static const int PAD_EVERY= 128 / sizeof(Key);                               //  keys between two padding keys
  static const int XPOSE    = (TILE + TILE / PAD_EVERY) * sizeof(Key);
  static const int LANES    = MB_RADIX / 2;                                     //  two 16-bit counters per word
  static const int CWORDS   = LANES * SORT_THREADS;
  static const int COUNTERS = (CWORDS + CWORDS / 32) * 4;
  static const int BYTES    = COUNTERS + XPOSE;                                 //  [counters][transpose area == stage]
};

__device__ __forceinline__ void key_zero(u32 &k)        { k = 0; }
__device__ __forceinline__ void key_zero(u64 &k)        { k = 0; }
__device__ __forceinline__ void key_zero(ulonglong2 &k) { k.x = 0; k.y = 0; }

//  The work of one tile.  FULL: the tile holds exactly TILE keys, so the per-key bounds checks vanish
//  (all tiles but the last one of each segment); the kernel picks the variant per tile, uniformly.
struct SweepShared {
  unsigned char *buf;  u32 *binOff;  u64 *gofs;  u32 *warpTot;
};

template <int KW, bool FULL>
__device__ __forceinline__ void
sweep_tile(const typename KeyT<KW>::type *__restrict__ in, typename KeyT<KW>::type *__restrict__ out,
           u64 tile, u64 begin, u32 nIn, u32 seg, bool isFirst,
           const u64 *__restrict__ gbase, u32 nPasses, u32 pass, u32 shift, u32 bits,
           u64 *__restrict__ status, u32 epoch, const SweepShared &sm,
           u64 *prof, long long &tk2, long long &tk3, long long &tk4, long long &tk5) {
  typedef typename KeyT<KW>::type Key;
  typedef SweepSmem<KW>           S;
  const int ITEMS = SortShape<KW>::ITEMS;
#define MB_TICK(i) do { if (prof && threadIdx.x == 0) tk##i = clock64(); } while (0)

  u32 *counters = (u32 *)sm.buf;
  Key *xpose    = (Key *)(sm.buf + S::COUNTERS);
  Key *stage    = xpose;
  u32 *binOff   = sm.binOff;
  u64 *gofs     = sm.gofs;
  u32 *warpTot  = sm.warpTot;

  const u32 t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const u32 n = FULL ? (u32)S::TILE : nIn;

  //  1. load striped (coalesced), transpose to blocked: thread t ends up with keys [t*ITEMS, (t+1)*ITEMS)
  Key key[ITEMS];
  {
    //  The warp's slice is private: only __syncwarp needed.  One padding key per 128 bytes; ITEMS keys are
    //  exactly 128 bytes, so the padded index of key j = t*ITEMS + i is t*(ITEMS+1) + i, and that of the
    //  striped key wbase + i*32 + lane is w0 + i*(32 + 32/PAD) -- constant strides the compiler can fold.
    const u32 wbase = warp * 32 * ITEMS;
    const u32 w0    = wbase + wbase / S::PAD_EVERY + lane + lane / S::PAD_EVERY;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      Key k;
      if (FULL || wbase + i * 32 + lane < n) k = in[begin + wbase + i * 32 + lane]; else key_zero(k);
      xpose[w0 + i * (32 + 32 / S::PAD_EVERY)] = k;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < ITEMS; i++)
      key[i] = xpose[t * (ITEMS + 1) + i];
  }
  //  (the transpose area is read only by the warp that wrote it; its next use, as the stage, comes after
  //  several block-wide barriers)

  //  2. private counters.  Digit d lives in word (d & 31) of my column, half (d >> 5).
  for (u32 i = t; i < (u32)(S::CWORDS + S::CWORDS / 32); i += SORT_THREADS) counters[i] = 0;
  __syncthreads();

  const u32 myFirst = t * ITEMS;                         //  my keys are tile positions [myFirst, myFirst + ITEMS)
  const u32 tcol    = t + (t >> 5);                      //  my column in the padded counter array
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    if (FULL || myFirst + i < n) {
      const u32 d = key_field(key[i], shift, bits);
      //  word (d & 31) of my column: linear index (d&31)*256 + t, padded by one word per 32;
      //  digits >= 32 use the high half-word
      u32 *cp = &counters[(d & 31) * (SORT_THREADS + SORT_THREADS / 32) + tcol];
      *cp += 1u << ((d & 32) >> 1);
    }
  }
  __syncthreads();

  //  raking scan over the CWORDS packed words: thread t owns the 32 consecutive words [32t, 32t+32)
  u32 tileLow = 0;                                       //  keys whose digit is < 32 (low halves)
  {
    u32 *seg32 = &counters[t * 33];
    u32 sum = 0;
#pragma unroll
    for (int j = 0; j < 32; j++) sum += seg32[j];
    u32 incl = sum;
    for (int o = 1; o < 32; o <<= 1) {
      const u32 v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (u32)o) incl += v;
    }
    if (lane == 31) warpTot[warp] = incl;
    __syncthreads();
    u32 run = incl - sum, tot = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) { const u32 c = warpTot[w]; if (w < (int)warp) run += c; tot += c; }
    tileLow = tot & 0xffff;
#pragma unroll
    for (int j = 0; j < 32; j++) { const u32 c = seg32[j]; seg32[j] = run; run += c; }
  }
  __syncthreads();
  MB_TICK(2);                                            //  2: keys loaded, counted, scanned

  //  start of every digit's run (thread 0's slot of digit d holds it), then my keys go straight to their
  //  place in the stage buffer: position = my (digit, thread) prefix, bumped as I walk my keys in order.
  if (t < MB_RADIX) {
    const u32 wv = counters[(t & 31) * (SORT_THREADS + SORT_THREADS / 32)];
    const u32 hi = t & 32;
    binOff[t] = ((wv >> (hi >> 1)) & 0xffff) + (hi ? tileLow : 0);
  }
  if (t == 0) binOff[MB_RADIX] = n;
  __syncthreads();                                       //  thread 0's column is read before its owner bumps it

  //  3a. publish this tile's digit counts as early as possible ...
  const u32 myCount = (t < MB_RADIX) ? binOff[t + 1] - binOff[t] : 0;
  if (t < MB_RADIX)
    st_volatile_u64(status + tile * MB_RADIX + t, ST_MAKE(isFirst ? 2 : 1, epoch, myCount));

  //  4a. ... put the keys in digit-sorted order in shared memory ...
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    if (FULL || myFirst + i < n) {
      const u32 d    = key_field(key[i], shift, bits);
      u32 *cp        = &counters[(d & 31) * (SORT_THREADS + SORT_THREADS / 32) + tcol];
      const u32 hi   = d & 32;
      const u32 wv   = *cp;
      *cp = wv + (1u << (hi >> 1));
      stage[((wv >> (hi >> 1)) & 0xffff) + (hi ? tileLow : 0)] = key[i];
    }
  }
  MB_TICK(3);                                            //  3: reorder issued

  //  3b. ... and only then wait for the earlier tiles of this segment (batched walk: LB status words
  //  per L2 round trip; the segment's first tile ends it).
  if (t < MB_RADIX) {
    const int LB = 8;
    const u32 d = t;
    u64 excl = 0;
    if (!isFirst) {
      const u64 *p = status + (tile - 1) * MB_RADIX + d;
      bool done = false;
      while (!done) {
        u64 v[LB];
#pragma unroll
        for (int b = 0; b < LB; b++) v[b] = ld_volatile_u64(p - (size_t)b * MB_RADIX);   //  may read before the segment: ignored below
#pragma unroll
        for (int b = 0; b < LB; b++) {
          if (done) break;
          if (ST_FLAG(v[b]) == 0 || ST_EPOCH(v[b]) != epoch) break;                      //  not published yet: poll again from here
          excl += ST_VALUE(v[b]);
          p -= MB_RADIX;
          if (ST_FLAG(v[b]) == 2) done = true;
        }
      }
      st_volatile_u64(status + tile * MB_RADIX + d, ST_MAKE(2, epoch, excl + myCount));
    }
    gofs[d] = gbase[((u64)seg * nPasses + pass) * MB_RADIX + d] + excl - binOff[d];
  }
  MB_TICK(4);                                            //  4: my own look-back done
  __syncthreads();
  MB_TICK(5);                                            //  5: everybody's look-back done

  //  4b. digit runs to HBM
  for (u32 j = t; j < n; j += SORT_THREADS) {
    const Key k = stage[j];
    out[gofs[key_field(k, shift, bits)] + j] = k;
  }
#undef MB_TICK
}

template <int KW>
__global__ void __launch_bounds__(SORT_THREADS, SortShape<KW>::MIN_CTAS)
sort_sweep_kernel(const typename KeyT<KW>::type *__restrict__ in, typename KeyT<KW>::type *__restrict__ out,
                  const TileInfo *__restrict__ info,
                  const u64 *__restrict__ gbase, u32 nPasses, u32 pass, u32 shift, u32 bits,
                  u64 *__restrict__ status, u64 *__restrict__ tileCounter, u32 epoch, u64 *__restrict__ prof) {
  typedef SweepSmem<KW> S;

  //  optional phase timing (debug): thread 0 accumulates clock deltas into prof[1..6]
  long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0, tk4 = 0, tk5 = 0, tk6 = 0;
#define MB_TICK(i) do { if (prof && threadIdx.x == 0) tk##i = clock64(); } while (0)
  MB_TICK(0);

  extern __shared__ __align__(16) unsigned char buf[];   //  S::BYTES
  __shared__ u32  binOff[MB_RADIX + 1];
  __shared__ u64  gofs[MB_RADIX];
  __shared__ u32  warpTot[SORT_WARPS];
  __shared__ u64  sTile;

  if (threadIdx.x == 0) sTile = atomicAdd((unsigned long long *)tileCounter, 1ull);
  __syncthreads();
  MB_TICK(1);                                            //  1: tile id

  const u64 tile = sTile;
  const TileInfo ti = info[tile];
  const u32  seg     = ti.seg & 0x7fffffffu;
  const bool isFirst = (ti.seg >> 31) != 0;
  SweepShared sm;
  sm.buf = buf;  sm.binOff = binOff;  sm.gofs = gofs;  sm.warpTot = warpTot;

#ifdef MB_SWEEP_ONLY_FULL
  if (true)
#else
  if (ti.n == (u32)S::TILE)
#endif
    sweep_tile<KW, true >(in, out, tile, ti.begin, ti.n, seg, isFirst, gbase, nPasses, pass, shift, bits, status, epoch, sm, prof, tk2, tk3, tk4, tk5);
#ifndef MB_SWEEP_ONLY_FULL
  else
    sweep_tile<KW, false>(in, out, tile, ti.begin, ti.n, seg, isFirst, gbase, nPasses, pass, shift, bits, status, epoch, sm, prof, tk2, tk3, tk4, tk5);
#endif

  MB_TICK(6);                                            //  6: my stores issued
  if (prof && threadIdx.x == 0) {
    atomicAdd((unsigned long long *)&prof[1], (unsigned long long)(tk1 - tk0));
    atomicAdd((unsigned long long *)&prof[2], (unsigned long long)(tk2 - tk1));
    atomicAdd((unsigned long long *)&prof[3], (unsigned long long)(tk3 - tk2));
    atomicAdd((unsigned long long *)&prof[4], (unsigned long long)(tk4 - tk3));
    atomicAdd((unsigned long long *)&prof[5], (unsigned long long)(tk5 - tk4));
    atomicAdd((unsigned long long *)&prof[6], (unsigned long long)(tk6 - tk5));
    atomicAdd((unsigned long long *)&prof[0], 1ull);
  }
#undef MB_TICK
}

//  ---- sweep kernel, second generation: one digit pass, shaped by (threads, digit bits) ------------------------
//
//  Same algorithm as sort_sweep_kernel (blocked keys, private 16-bit counters, raking scan, publish,
//  reorder in shared memory, look-back inside the segment, coalesced digit runs) on a smaller instruction
//  budget, and for any digit width:
//    * 16-bit counter accesses: counter word (d mod ROWS, t) holds digit d in its low and digit d + ROWS in
//      its high half (ROWS = half the radix); the scan leaves every counter holding the stage position of the
//      thread's next key of that digit (the high halves start at the number of low-half keys), so placing a
//      key is load / +1 / store / store;
//    * thread t rakes the ROWS consecutive words [ROWS t, ROWS t + ROWS) of the row-major counter array with
//      128-bit accesses; 4 words of padding behind every such segment keep them bank-conflict free;
//    * whole-byte digits are extracted with one PRMT; output offsets are 32 bits from the segment start when
//      the segments are shorter than 2^32 keys.
//  Shapes (mb_sort_version):  2: 128 threads x 8 bits, 2 CTAs/SM   3: 256 threads x 6 bits, 3 CTAs/SM (default)
//                             4: 256 threads x 7 bits, 2 CTAs/SM   5: 512 x 6, 2 CTAs/SM   6: 256 x 6, 2 CTAs/SM

template <int VER> struct SweepShape;
template <> struct SweepShape<2> { static const int THREADS = 128, RBITS = 8, MIN_CTAS = 2; };
template <> struct SweepShape<3> { static const int THREADS = 256, RBITS = 6, MIN_CTAS = 3; };
template <> struct SweepShape<4> { static const int THREADS = 256, RBITS = 7, MIN_CTAS = 2; };
//  not yet measured (written without a GPU at hand; tools/ab_bench.py ... sortonly compares all shapes):
template <> struct SweepShape<5> { static const int THREADS = 512, RBITS = 6, MIN_CTAS = 2; };   //  32 warps/SM, 16 keys per thread
template <> struct SweepShape<6> { static const int THREADS = 256, RBITS = 6, MIN_CTAS = 2; };   //  16 warps/SM, 128 registers: no address recomputation
template <> struct SweepShape<7> { static const int THREADS = 256, RBITS = 6, MIN_CTAS = 3; };   //  shape 3 as a persistent kernel that prefetches the next tile (sortp_sweep_kernel)

template <int KW, int VER> struct SweepX {
  typedef typename KeyT<KW>::type Key;
  static const int THREADS  = SweepShape<VER>::THREADS;
  static const int WARPS    = THREADS / 32;
  static const int RBITS    = SweepShape<VER>::RBITS;
  static const int RADIX    = 1 << RBITS;
  static const int ROWS     = RADIX / 2;                                        //  counter rows (digit pairs d, d + ROWS)
  static const int SEG      = ROWS + 4;                                         //  a thread's rake segment + padding, in words
  static const int PER_ROW  = THREADS / ROWS;                                   //  rake segments per counter row
  static const int ROWW     = THREADS + 4 * PER_ROW;                            //  padded row stride in words
  static const int TILE     = (KW == 0 ? 8192 : KW == 1 ? 4096 : 2048);         //  == mb_sort_tile_keys()
  static const int ITEMS    = TILE / THREADS;
  static const int XPOSE    = (TILE + THREADS) * sizeof(Key);                   //  one padding key per thread (ITEMS + 1 stride)
  static const int COUNTERS = THREADS * SEG * 4;
  static const int BYTES    = COUNTERS + XPOSE;                                 //  [counters][transpose area == stage]
  static const int DPT      = (RADIX + THREADS - 1) / THREADS;                  //  digits a thread publishes / resolves
  static_assert(THREADS % ROWS == 0 && ROWS % 32 == 0, "shape");
};

struct SweepXShared {
  unsigned char *buf;  u32 *binOff;  u64 *gofs;  u32 *gofs32;  u32 *warpTot;
};

//  Digit of a pass.  BYTE: the digit is one whole byte of the key (8 bits at a multiple of 8) -- a single
//  PRMT instead of shift + mask; the pass plan of the 8-bit shape makes every pass but possibly the top one so.
template <bool BYTE> __device__ __forceinline__ u32 digit_of(u32 k, u32 shift, u32 bits) {
  return BYTE ? __byte_perm(k, 0, 0x4440 | (shift >> 3)) : key_field(k, shift, bits);
}
template <bool BYTE> __device__ __forceinline__ u32 digit_of(u64 k, u32 shift, u32 bits) {
  if (!BYTE) return key_field(k, shift, bits);
  const u32 w = (shift & 32) ? (u32)(k >> 32) : (u32)k;
  return __byte_perm(w, 0, 0x4440 | ((shift >> 3) & 3));
}
template <bool BYTE> __device__ __forceinline__ u32 digit_of(const ulonglong2 &k, u32 shift, u32 bits) {
  if (!BYTE) return key_field(k, shift, bits);
  const u64 v = (shift & 64) ? k.y : k.x;
  const u32 w = (shift & 32) ? (u32)(v >> 32) : (u32)v;
  return __byte_perm(w, 0, 0x4440 | ((shift >> 3) & 3));
}

//  16-bit counter of (digit d, thread t): row d mod ROWS, low half for d < ROWS.  cbOff = 4 * (t + 4 * (t / ROWS)); the
//  result is buf + (d mod ROWS) * ROWBYTES + (d / ROWS) * 2 + cbOff, written so that it compiles to a shift and two multiply-adds.
template <int ROWS, int ROWW>
__device__ __forceinline__ unsigned short *counter_of(unsigned char *buf, u32 cbOff, u32 d) {
  u32 off = d * (ROWW * 4) + cbOff;
  off -= (d / ROWS) * (ROWS * ROWW * 4 - 2);
  return (unsigned short *)(buf + off);
}

//  Persistent form (sortp_sweep_kernel): a CTA keeps claiming tiles.  While it works on a tile it claims the next one (the
//  atomic's round trip hides behind the counting), fetches that tile's descriptor after the scan, and -- as soon as this
//  tile's keys sit in the stage buffer and their registers are free -- loads the next tile's keys, so the loads fly during
//  the look-back and the copy-out instead of in front of an idle CTA.
template <int KW, int VER> struct SweepNext {
  typename KeyT<KW>::type key[SweepX<KW, VER>::ITEMS];   //  in: this tile's keys, striped (thread t holds keys wbase + 32 i + lane); out: the next tile's
  const TileInfo *info;
  u64   nTiles;
  u64  *tileCounter;
  u64  *sNext;                                            //  shared memory word: the tile claimed for the next round
  u64   nextTile;
  TileInfo nextInfo;
};

template <int KW, int VER, bool FULL, bool BYTE, bool PRE = false>
__device__ __forceinline__ void
sweepx_tile(const typename KeyT<KW>::type *__restrict__ in, typename KeyT<KW>::type *__restrict__ out,
            u64 tile, u64 begin, u32 nIn, u32 seg, bool isFirst,
            const u64 *__restrict__ gbase, u32 nPasses, u32 pass, u32 shift, u32 bits, bool wide,
            u64 *__restrict__ status, u32 epoch, const SweepXShared &sm, SweepNext<KW, VER> *nx = nullptr) {
  typedef typename KeyT<KW>::type Key;
  typedef SweepX<KW, VER>         S;
  const int ITEMS = S::ITEMS, THREADS = S::THREADS, ROWS = S::ROWS, RADIX = S::RADIX;

  u32 *counters = (u32 *)sm.buf;
  Key *xpose    = (Key *)(sm.buf + S::COUNTERS);
  Key *stage    = xpose;
  u32 *binOff   = sm.binOff;
  u64 *gofs     = sm.gofs;
  u32 *gofs32   = sm.gofs32;
  u32 *warpTot  = sm.warpTot;

  const u32 t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const u32 n = FULL ? (u32)S::TILE : nIn;

  //  1. load striped (coalesced), transpose to blocked: thread t ends up with keys [t*ITEMS, (t+1)*ITEMS).
  //     Key j of the tile sits at xpose[j + j / ITEMS] (one padding key per thread: conflict-free blocked
  //     reads); a warp only touches the slice of its own 32 threads, so __syncwarp is enough.
  //     j = wbase + 32 i + lane: the padding term splits into constants the compiler folds into the addresses.
  Key key[ITEMS];
  {
    const u32 wbase = warp * 32 * ITEMS;
    const u32 w0    = wbase + wbase / ITEMS + lane + (ITEMS >= 32 ? 0 : lane / ITEMS);
    //  one base pointer per thread, the rows at constant offsets from it (ncu of the first form: two LEA per load for the
    //  64-bit index arithmetic -- 178 of the load phase's 312 instructions per warp were address and predicate math)
    const Key *pin = in + (begin + wbase + lane);
    Key *px = xpose + w0;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const u32 j = wbase + i * 32 + lane;
      Key k;
      if (PRE) k = nx->key[i];
      else if (FULL || j < n) k = pin[i * 32]; else key_zero(k);
      px[i * 32 + (i * 32) / ITEMS] = k;
    }
    if (PRE && t == 0) *nx->sNext = atomicAdd((unsigned long long *)nx->tileCounter, 1ull);     //  claim the next tile now, read it after the scan
    __syncwarp();
#pragma unroll
    for (int i = 0; i < ITEMS; i++)
      key[i] = xpose[t * (ITEMS + 1) + i];
  }

  //  2. private counters: thread t zeroes its rake segment, then bumps its 16-bit counter of each key's digit
  uint4 *segp = (uint4 *)&counters[t * S::SEG];
#pragma unroll
  for (int j = 0; j < S::SEG / 4; j++) segp[j] = make_uint4(0, 0, 0, 0);
  __syncthreads();

  const u32 myFirst = t * ITEMS;
  const u32 cbOff   = 4 * (t + 4 * (t / ROWS));          //  my column of the counter rows
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    if (FULL || myFirst + i < n) {
      unsigned short *cp = counter_of<ROWS, S::ROWW>(sm.buf, cbOff, digit_of<BYTE>(key[i], shift, bits));
      *cp = (unsigned short)(*cp + 1);
    }
  }
  __syncthreads();

  //  raking scan, digit-major, over packed words.  Digits >= ROWS (high halves) follow all lower digits in the
  //  stage: their prefixes start at the number of low-half keys.
  {
    u32 sum = 0;
#pragma unroll
    for (int j = 0; j < ROWS / 4; j++) { const uint4 v = segp[j]; sum += v.x + v.y + v.z + v.w; }
    u32 incl = sum;
    for (int o = 1; o < 32; o <<= 1) {
      const u32 v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (u32)o) incl += v;
    }
    if (lane == 31) warpTot[warp] = incl;
    __syncthreads();
    u32 run = incl - sum, tot = 0;
#pragma unroll
    for (int w = 0; w < S::WARPS; w++) { const u32 c = warpTot[w]; if (w < (int)warp) run += c; tot += c; }
    run += tot << 16;                                    //  + low-half total in the high half (both halves stay <= 2 * TILE)
    if (t % S::PER_ROW == 0) {                           //  my segment starts a counter row: thread 0's slot of digits dl, dl + ROWS
      const u32 dl = t / S::PER_ROW;
      binOff[dl]        = run & 0xffff;                  //  start of the two digits' runs in the tile
      binOff[dl + ROWS] = run >> 16;
    }
#pragma unroll
    for (int j = 0; j < ROWS / 4; j++) {
      const uint4 v = segp[j];
      uint4 o;
      o.x = run;  run += v.x;
      o.y = run;  run += v.y;
      o.z = run;  run += v.z;
      o.w = run;  run += v.w;
      segp[j] = o;
    }
  }
  if (t == 0) binOff[RADIX] = n;
  __syncthreads();
  if (PRE) {
    nx->nextTile = *nx->sNext;
    if (nx->nextTile < nx->nTiles) nx->nextInfo = nx->info[nx->nextTile];
  }

  //  3a. publish this tile's digit counts as early as possible ...
  //  (digit ROWS-1's run ends where digit ROWS's begins, so binOff[d + 1] - binOff[d] holds for every d)
  u32 myCount[S::DPT];
#pragma unroll
  for (int h = 0; h < S::DPT; h++) {
    const u32 d = t + THREADS * h;
    if (d < (u32)RADIX) {
      myCount[h] = binOff[d + 1] - binOff[d];
      st_volatile_u64(status + tile * RADIX + d, ST_MAKE(isFirst ? 2 : 1, epoch, myCount[h]));
    }
  }

  //  4a. ... put the keys in digit-sorted order in shared memory: my counter of the key's digit now holds the
  //  place of my next key with that digit ...
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    if (FULL || myFirst + i < n) {
      unsigned short *cp = counter_of<ROWS, S::ROWW>(sm.buf, cbOff, digit_of<BYTE>(key[i], shift, bits));
      const u32 pos = *cp;
      *cp = (unsigned short)(pos + 1);
      stage[pos] = key[i];
    }
  }

  //  (persistent form: my key registers are free now -- the next tile's keys start their way from HBM)
  if (PRE && nx->nextTile < nx->nTiles) {
    const u32 wbase = warp * 32 * ITEMS;
    const u64 nb = nx->nextInfo.begin;
    const u32 nn = nx->nextInfo.n;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const u32 j = wbase + i * 32 + lane;
      if (j < nn) nx->key[i] = in[nb + j]; else key_zero(nx->key[i]);
    }
  }

  //  3b. ... and only then wait for the earlier tiles of this segment (batched walk, LB status words per
  //  L2 round trip; the segment's first tile ends it).
  const u64 segBase = gbase[((u64)seg * nPasses + pass) * RADIX];              //  first key of the segment (digit 0 starts there)
#pragma unroll
  for (int h = 0; h < S::DPT; h++) {
    const int LB = (S::DPT > 1) ? 4 : 8;
    const u32 d = t + THREADS * h;
    if (d < (u32)RADIX) {
      u64 excl = 0;
      if (!isFirst) {
        const u64 *p = status + (tile - 1) * RADIX + d;
        bool done = false;
        while (!done) {
          u64 v[LB];
#pragma unroll
          for (int b = 0; b < LB; b++) v[b] = ld_volatile_u64(p - (size_t)b * RADIX);     //  may read before the segment: ignored below
#pragma unroll
          for (int b = 0; b < LB; b++) {
            if (done) break;
            if (ST_FLAG(v[b]) == 0 || ST_EPOCH(v[b]) != epoch) break;                     //  not published yet: poll again from here
            excl += ST_VALUE(v[b]);
            p -= RADIX;
            if (ST_FLAG(v[b]) == 2) done = true;
          }
        }
        st_volatile_u64(status + tile * RADIX + d, ST_MAKE(2, epoch, excl + myCount[h]));
      }
      const u64 g = gbase[((u64)seg * nPasses + pass) * RADIX + d] + excl - binOff[d];   //  + stage index = global index
      gofs[d]   = g;
      gofs32[d] = (u32)(g - segBase);                    //  same, relative to the segment (mod 2^32)
    }
  }
  __syncthreads();

  //  4b. digit runs to HBM.  Segments shorter than 2^32 keys: 32-bit offsets from the segment's first key.
  if (!wide) {
    Key *outSeg = out + segBase;
    MB_OPAQUE_PTR(outSeg);                               //  keep it a base pointer: one IMAD.WIDE per store instead of 64-bit index math
    if (FULL) {
#pragma unroll 16
      for (u32 j = t; j < (u32)S::TILE; j += THREADS) {
        const Key k = stage[j];
        outSeg[(u32)(gofs32[digit_of<BYTE>(k, shift, bits)] + j)] = k;
      }
    } else {
      for (u32 j = t; j < n; j += THREADS) {
        const Key k = stage[j];
        outSeg[(u32)(gofs32[digit_of<BYTE>(k, shift, bits)] + j)] = k;
      }
    }
  } else {
    for (u32 j = t; j < n; j += THREADS) {
      const Key k = stage[j];
      out[gofs[digit_of<BYTE>(k, shift, bits)] + j] = k;
    }
  }
}

template <int KW, int VER>
__global__ void __launch_bounds__(SweepShape<VER>::THREADS, SweepShape<VER>::MIN_CTAS)
sortx_sweep_kernel(const typename KeyT<KW>::type *__restrict__ in, typename KeyT<KW>::type *__restrict__ out,
                   const TileInfo *__restrict__ info,
                   const u64 *__restrict__ gbase, u32 nPasses, u32 pass, u32 shift, u32 bits, u32 wide,
                   u64 *__restrict__ status, u64 *__restrict__ tileCounter, u32 epoch, u64 nTiles, u32 pfDist) {
  typedef SweepX<KW, VER> S;
  extern __shared__ __align__(16) unsigned char buf[];   //  S::BYTES
  __shared__ u32  binOff[S::RADIX + 1];
  __shared__ u64  gofs[S::RADIX];
  __shared__ u32  gofs32[S::RADIX];
  __shared__ u32  warpTot[S::WARPS];
  __shared__ u64  sTile;

  if (threadIdx.x == 0) sTile = atomicAdd((unsigned long long *)tileCounter, 1ull);
  __syncthreads();

  const u64 tile = sTile;
  const TileInfo ti = info[tile];
  //  The keys of the tile that is claimed pfDist tiles from now start their way from HBM to L2 (one 128-byte line per
  //  thread and round): the CTA that gets that tile then waits for L2, not for DRAM, in front of an otherwise idle tile.
  if (pfDist && tile + pfDist < nTiles) {
    const TileInfo tn = info[tile + pfDist];
    const unsigned char *p0 = (const unsigned char *)(in + tn.begin);
    for (u32 o = threadIdx.x * 128; o < tn.n * (u32)sizeof(typename KeyT<KW>::type); o += S::THREADS * 128) prefetch_l2(p0 + o);
  }
  const u32  seg     = ti.seg & 0x7fffffffu;
  const bool isFirst = (ti.seg >> 31) != 0;
  SweepXShared sm;
  sm.buf = buf;  sm.binOff = binOff;  sm.gofs = gofs;  sm.gofs32 = gofs32;  sm.warpTot = warpTot;

  const bool full = (ti.n == (u32)S::TILE);
  const bool byte = (S::RBITS == 8) && (bits == 8) && ((shift & 7) == 0);
#define MB_SX_CALL(F, B) sweepx_tile<KW, VER, F, B>(in, out, tile, ti.begin, ti.n, seg, isFirst, gbase, nPasses, pass, shift, bits, wide != 0, status, epoch, sm)
  if (S::RBITS == 8 && byte) { if (full) MB_SX_CALL(true, true);  else MB_SX_CALL(false, true);  }
  else                       { if (full) MB_SX_CALL(true, false); else MB_SX_CALL(false, false); }
#undef MB_SX_CALL
}

//  Persistent form of the sweep (MB_SORT=7): grid = the CTAs that fit the device at once; tiles are still claimed from the
//  pass's counter in increasing order, so a tile's predecessors are always claimed by a resident CTA that depends on lower
//  tiles only -- the look-back cannot dead-lock.
template <int KW, int VER>
__global__ void __launch_bounds__(SweepShape<VER>::THREADS, SweepShape<VER>::MIN_CTAS)
sortp_sweep_kernel(const typename KeyT<KW>::type *__restrict__ in, typename KeyT<KW>::type *__restrict__ out,
                   const TileInfo *__restrict__ info, u64 nTiles,
                   const u64 *__restrict__ gbase, u32 nPasses, u32 pass, u32 shift, u32 bits, u32 wide,
                   u64 *__restrict__ status, u64 *__restrict__ tileCounter, u32 epoch) {
  typedef SweepX<KW, VER> S;
  extern __shared__ __align__(16) unsigned char buf[];   //  S::BYTES
  __shared__ u32  binOff[S::RADIX + 1];
  __shared__ u64  gofs[S::RADIX];
  __shared__ u32  gofs32[S::RADIX];
  __shared__ u32  warpTot[S::WARPS];
  __shared__ u64  sTile, sNext;

  if (threadIdx.x == 0) sTile = atomicAdd((unsigned long long *)tileCounter, 1ull);
  __syncthreads();
  u64 tile = sTile;
  if (tile >= nTiles) return;
  TileInfo ti = info[tile];

  SweepXShared sm;
  sm.buf = buf;  sm.binOff = binOff;  sm.gofs = gofs;  sm.gofs32 = gofs32;  sm.warpTot = warpTot;
  SweepNext<KW, VER> nx;
  nx.info = info;  nx.nTiles = nTiles;  nx.tileCounter = tileCounter;  nx.sNext = &sNext;  nx.nextTile = ~0ull;
  {                                                       //  the first tile's keys: the only exposed load of this CTA
    const u32 lane = threadIdx.x & 31, wbase = (threadIdx.x >> 5) * 32 * S::ITEMS;
#pragma unroll
    for (int i = 0; i < S::ITEMS; i++) {
      const u32 j = wbase + i * 32 + lane;
      if (j < ti.n) nx.key[i] = in[ti.begin + j]; else key_zero(nx.key[i]);
    }
  }
  const bool byte = (S::RBITS == 8) && (bits == 8) && ((shift & 7) == 0);
  for (;;) {
    const u32  seg     = ti.seg & 0x7fffffffu;
    const bool isFirst = (ti.seg >> 31) != 0;
    const bool full    = (ti.n == (u32)S::TILE);
#define MB_SP_CALL(F, B) sweepx_tile<KW, VER, F, B, true>(in, out, tile, ti.begin, ti.n, seg, isFirst, gbase, nPasses, pass, shift, bits, wide != 0, status, epoch, sm, &nx)
    if (S::RBITS == 8 && byte) { if (full) MB_SP_CALL(true, true);  else MB_SP_CALL(false, true);  }
    else                       { if (full) MB_SP_CALL(true, false); else MB_SP_CALL(false, false); }
#undef MB_SP_CALL
    if (nx.nextTile >= nTiles) break;
    tile = nx.nextTile;  ti = nx.nextInfo;
    __syncthreads();                                      //  everybody has copied its keys out of the stage buffer
  }
}

}  //  namespace


//  ---- host side ---------------------------------------------------------------------------------

//  Which sweep kernel: 1 (first generation, 6-bit digits) or a shape of the second generation (2: 128 threads x
//  8 bits, 3: 256 x 6, 4: 256 x 7).  MB_SORT=1..4 in the environment, or mb_sort_set_version(); fixed for the
//  lifetime of a plan / workspace.
static int g_sortVersion = 0;
int mb_sort_version(void) {
  if (g_sortVersion == 0) {
    const char *e = getenv("MB_SORT");
    g_sortVersion = (e && e[0] >= '1' && e[0] <= '7' && e[1] == 0) ? e[0] - '0' : MB_SORT_DEFAULT_VERSION;
  }
  return g_sortVersion;
}
void mb_sort_set_version(int v) { g_sortVersion = (v >= 1 && v <= 7) ? v : MB_SORT_DEFAULT_VERSION; }
u32  mb_sort_radix_bits(void) {
  switch (mb_sort_version()) {
    case 2:  return SweepShape<2>::RBITS;
    case 3:  return SweepShape<3>::RBITS;
    case 4:  return SweepShape<4>::RBITS;
    case 5:  return SweepShape<5>::RBITS;
    case 6:  return SweepShape<6>::RBITS;
    case 7:  return SweepShape<7>::RBITS;
    default: return MB_RADIX_BITS;
  }
}
u32  mb_sort_radix(void)        { return 1u << mb_sort_radix_bits(); }

void mb_sort_plan(u32 sortBits, MbSortPlan *plan) {
  const u32 RB = mb_sort_radix_bits();
  const u32 P = (sortBits + RB - 1) / RB;
  plan->nPasses = P;
  plan->maxSegKeys = 0;
  u32 shift = 0;
  for (u32 p = 0; p < P; p++) {
    //  v1: digits of (almost) equal width.  v2: whole bytes from the bottom (one PRMT per digit), the top
    //  pass takes what is left.
    u32 b = sortBits / P + (p < sortBits % P ? 1 : 0);
    if (RB == 8) b = (sortBits - shift < 8) ? sortBits - shift : 8;
    plan->shift[p] = shift;
    plan->bits[p]  = b;
    shift += b;
  }
}

u32 mb_sort_tile_keys(u32 keyBytes) {                  //  the same for both sweep kernels: 8192 / 4096 / 2048
  static_assert(SORT_THREADS * SortShape<0>::ITEMS == SweepX<0, 2>::TILE && SweepX<0, 2>::TILE == SweepX<0, 3>::TILE, "tile size");
  static_assert(SORT_THREADS * SortShape<1>::ITEMS == SweepX<1, 2>::TILE && SweepX<1, 4>::TILE == SweepX<1, 3>::TILE, "tile size");
  static_assert(SORT_THREADS * SortShape<2>::ITEMS == SweepX<2, 2>::TILE && SweepX<2, 4>::TILE == SweepX<2, 3>::TILE, "tile size");
  return SORT_THREADS * (keyBytes == 4 ? SortShape<0>::ITEMS : keyBytes == 8 ? SortShape<1>::ITEMS : SortShape<2>::ITEMS);
}

//  Debug: device buffer of 8 u64 that receives per-phase clock sums of the sweep kernel (NULL = off).
static u64 *g_sortProf = nullptr;
void mb_sort_set_profile_buffer(u64 *dProf) { g_sortProf = dProf; }

//  workspace: [ghist: nSegs*nPasses*R u64][tile counters: 32 u64][guard: 8*R u64][status: nTiles*R u64][tile info: nTiles*16 B]
//  (R = mb_sort_radix(); the guard lets the batched look-back of the first tiles read a few status rows "before" row 0)
static const size_t kCounters    = 32;
static size_t statusGuard(void) { return 8 * (size_t)mb_sort_radix(); }

size_t mb_sort_workspace_bytes(u64 nTiles, u32 nSegs, u32 nPasses) {
  const size_t R = mb_sort_radix();
  return ((size_t)nSegs * nPasses * R + kCounters + statusGuard() + (size_t)nTiles * R) * sizeof(u64) + (size_t)nTiles * sizeof(TileInfo) + 64;
}

const void *mb_sort_tile_info(void *dWorkspace, u64 nTiles, u32 nSegs, u32 nPasses) {
  const size_t R = mb_sort_radix();
  u64 *ghist   = (u64 *)dWorkspace;
  u64 *counter = ghist + (size_t)nSegs * nPasses * R;
  u64 *status  = counter + kCounters + statusGuard();
  return (const void *)(((uintptr_t)(status + (size_t)nTiles * R) + 15) & ~(uintptr_t)15);
}

template <int KW>
static cudaError_t sort_impl(void *dKeysA, void *dKeysB, u64 nKeys,
                             const u64 *dSegStart, const u64 *dSegTile, u32 nSegs, u64 nTiles,
                             const MbSortPlan &plan, void *dWorkspace, size_t workspaceBytes,
                             int *resultInB, u32 *nLaunches, cudaStream_t s, cudaEvent_t evAfterHist) {
  typedef typename KeyT<KW>::type Key;
  *resultInB = 0;
  if (nLaunches) *nLaunches = 0;
  if (nKeys == 0 || nTiles == 0) {
    if (evAfterHist) cudaEventRecord(evAfterHist, s);
    return cudaSuccess;
  }
  if (plan.nPasses > SORT_MAX_PASSES) return cudaErrorInvalidValue;
  if (workspaceBytes < mb_sort_workspace_bytes(nTiles, nSegs, plan.nPasses)) return cudaErrorInvalidValue;

  const int    version = mb_sort_version();
  const size_t R = mb_sort_radix();
  //  v2 writes with 32-bit offsets from the segment start when every segment is known to be shorter than 2^32 keys
  const u32    wide = (plan.maxSegKeys == 0 || plan.maxSegKeys >= ((u64)1 << 32)) ? 1 : 0;
  if ((size_t)plan.nPasses * R > (size_t)HIST_WORDS) return cudaErrorInvalidValue;
  u64 *ghist   = (u64 *)dWorkspace;
  u64 *counter = ghist + (size_t)nSegs * plan.nPasses * R;
  u64 *status  = counter + kCounters + statusGuard();
  TileInfo *info = (TileInfo *)mb_sort_tile_info(dWorkspace, nTiles, nSegs, plan.nPasses);
  const u32 tileKeys = SORT_THREADS * SortShape<KW>::ITEMS;

  sort_tiles_kernel<<<(unsigned)((nTiles + 255) / 256), 256, 0, s>>>(dSegStart, dSegTile, nSegs, nTiles, tileKeys, info);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (nLaunches) *nLaunches += 1;
  if (plan.nPasses == 0) {                               //  nothing to sort, but the tile table is part of the contract
    if (evAfterHist) cudaEventRecord(evAfterHist, s);
    return cudaSuccess;
  }

  const size_t zeroBytes = (size_t)((uint8_t *)(status + (size_t)nTiles * R) - (uint8_t *)dWorkspace);
  if ((e = cudaMemsetAsync(dWorkspace, 0, zeroBytes, s)) != cudaSuccess) return e;

  SortPassTable pt;
  pt.nPasses = plan.nPasses;
  for (u32 p = 0; p < SORT_MAX_PASSES; p++) { pt.shift[p] = p < plan.nPasses ? plan.shift[p] : 0; pt.bits[p] = p < plan.nPasses ? plan.bits[p] : 0; }

  const u64 histGrid = (nTiles + HIST_TILES_PER_CTA - 1) / HIST_TILES_PER_CTA;
  //  Measured on B200 (profiles/r1_ab_8gbp_countopts.jsonl): the joint form is no faster (22.2 ms either way per 6.9 G keys:
  //  the kernel is not bound by the number of shared-memory atomics) -- the single form stays the default.
  //  Digit histograms: adjacent passes are counted jointly (one shared-memory atomic per PAIR of digits) while the joint
  //  tables stay small; MB_HIST_SINGLE=1 keeps the one-atomic-per-digit kernel for A/B.
  const bool histSingle = getenv("MB_HIST_SINGLE") != nullptr;
  SortHistGroups hg;
  memset(&hg, 0, sizeof(hg));
  for (u32 p = 0; p < plan.nPasses; ) {                 //  pair adjacent passes while the joint table stays small
    const u32 g = hg.nGroups++;
    hg.shift[g] = plan.shift[p];  hg.bitsLo[g] = plan.bits[p];  hg.bitsAll[g] = plan.bits[p];  hg.firstPass[g] = p;  hg.offset[g] = hg.total;
    if (!histSingle && p + 1 < plan.nPasses && plan.bits[p] + plan.bits[p + 1] <= 12 && hg.total + (1u << (plan.bits[p] + plan.bits[p + 1])) <= 16384) {
      hg.bitsAll[g] += plan.bits[p + 1];
      p += 2;
    } else p += 1;
    hg.total += 1u << hg.bitsAll[g];
  }
  if (histSingle || hg.nGroups > 8 || hg.total * 4 > 96 * 1024) {
    sort_hist_kernel<KW><<<(unsigned)histGrid, SORT_THREADS, 0, s>>>((const Key *)dKeysA, dSegStart, dSegTile, nSegs, nTiles, pt, (u32)R, ghist,
                                                                     getenv("MB_HIST_SIMPLE") ? 0u : 1u);
  } else {
#define MB_HIST2(NG) do {                                                                                                                   \
      if ((e = cudaFuncSetAttribute(sort_hist2_kernel<KW, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(hg.total * 4))) != cudaSuccess) return e;  \
      sort_hist2_kernel<KW, NG><<<(unsigned)histGrid, SORT_THREADS, hg.total * 4, s>>>((const Key *)dKeysA, dSegStart, dSegTile, nSegs, nTiles, hg, plan.nPasses, (u32)R, ghist); } while (0)
    switch (hg.nGroups) {
      case 1: MB_HIST2(1); break;  case 2: MB_HIST2(2); break;  case 3: MB_HIST2(3); break;  case 4: MB_HIST2(4); break;
      case 5: MB_HIST2(5); break;  case 6: MB_HIST2(6); break;  case 7: MB_HIST2(7); break;  default: MB_HIST2(8); break;
    }
#undef MB_HIST2
  }
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  sort_scan_kernel<<<dim3(nSegs, plan.nPasses), (unsigned)R, 0, s>>>(ghist, dSegStart, plan.nPasses);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (nLaunches) *nLaunches += 2;
  if (evAfterHist) cudaEventRecord(evAfterHist, s);

  Key *src = (Key *)dKeysA, *dst = (Key *)dKeysB;
  if ((e = cudaFuncSetAttribute(sort_sweep_kernel<KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, SweepSmem<KW>::BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(sortx_sweep_kernel<KW, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SweepX<KW, 2>::BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(sortx_sweep_kernel<KW, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SweepX<KW, 3>::BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(sortx_sweep_kernel<KW, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SweepX<KW, 4>::BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(sortx_sweep_kernel<KW, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, SweepX<KW, 5>::BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(sortx_sweep_kernel<KW, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, SweepX<KW, 6>::BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(sortp_sweep_kernel<KW, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, SweepX<KW, 7>::BYTES)) != cudaSuccess) return e;
  //  MB_SWEEP_PREFETCH=<tiles>: L2 prefetch distance of the sweep kernel (0 = off)
  const u32 pfDist = getenv("MB_SWEEP_PREFETCH") ? (u32)strtoul(getenv("MB_SWEEP_PREFETCH"), nullptr, 10) : MB_SWEEP_PREFETCH_DEFAULT;
  for (u32 p = 0; p < plan.nPasses; p++) {
#define MB_SX_LAUNCH(V) do { typedef SweepX<KW, V> SX;                                                                                         \
      sortx_sweep_kernel<KW, V><<<(unsigned)nTiles, SX::THREADS, SX::BYTES, s>>>(src, dst, info, ghist, plan.nPasses, p,                       \
                                                                     plan.shift[p], plan.bits[p], wide, status, counter + p, p + 1, nTiles, pfDist); } while (0)
    if      (version == 2) MB_SX_LAUNCH(2);
    else if (version == 3) MB_SX_LAUNCH(3);
    else if (version == 4) MB_SX_LAUNCH(4);
    else if (version == 5) MB_SX_LAUNCH(5);
    else if (version == 6) MB_SX_LAUNCH(6);
    else if (version == 7) {
      typedef SweepX<KW, 7> SX;
      static int perSm = 0, sms = 0;                         //  co-resident CTAs: the grid of the persistent kernel
      if (!perSm) {
        int dev = 0;  cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, sortp_sweep_kernel<KW, 7>, SX::THREADS, SX::BYTES)) != cudaSuccess) return e;
        if (perSm < 1) perSm = 1;
      }
      const u64 grid = std::min<u64>(nTiles, (u64)sms * perSm);
      sortp_sweep_kernel<KW, 7><<<(unsigned)grid, SX::THREADS, SX::BYTES, s>>>(src, dst, info, nTiles, ghist, plan.nPasses, p,
                                                                     plan.shift[p], plan.bits[p], wide, status, counter + p, p + 1);
    }
    else
      sort_sweep_kernel<KW><<<(unsigned)nTiles, SORT_THREADS, SweepSmem<KW>::BYTES, s>>>(src, dst, info, ghist, plan.nPasses, p,
                                                                     plan.shift[p], plan.bits[p], status, counter + p, p + 1, g_sortProf);
#undef MB_SX_LAUNCH
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (nLaunches) *nLaunches += 1;
    Key *tmp = src; src = dst; dst = tmp;
  }
  *resultInB = (plan.nPasses & 1);
  return cudaSuccess;
}

cudaError_t mb_launch_sort(u32 keyBytes, void *dKeysA, void *dKeysB, u64 nKeys,
                           const u64 *dSegStart, const u64 *dSegTile, u32 nSegs, u64 nTiles,
                           const MbSortPlan &plan, void *dWorkspace, size_t workspaceBytes,
                           int *resultInB, u32 *nLaunches, cudaStream_t s, cudaEvent_t evAfterHist) {
  if (keyBytes == 4)
    return sort_impl<0>(dKeysA, dKeysB, nKeys, dSegStart, dSegTile, nSegs, nTiles, plan, dWorkspace, workspaceBytes, resultInB, nLaunches, s, evAfterHist);
  return keyBytes == 8
    ? sort_impl<1>(dKeysA, dKeysB, nKeys, dSegStart, dSegTile, nSegs, nTiles, plan, dWorkspace, workspaceBytes, resultInB, nLaunches, s, evAfterHist)
    : sort_impl<2>(dKeysA, dKeysB, nKeys, dSegStart, dSegTile, nSegs, nTiles, plan, dWorkspace, workspaceBytes, resultInB, nLaunches, s, evAfterHist);
}
