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
//   sortx_sweep_kernel  per pass: tile = THREADS x ITEMS keys (8192 u32 / 4096 u64 / 2048 u128 keys)
//       1. coalesced (warp-striped) load, transposed through shared memory so that every thread
//          owns ITEMS CONSECUTIVE keys -- rank order == memory order, which LSD needs (stable);
//       2. rank by PRIVATE counters: thread t bumps its own 16-bit counter of the key's digit
//          (two digits per 32-bit word, no atomics, no warp votes), a raking scan over the
//          packed counters turns them into tile-wide positions;
//       3. the tile publishes its digit counts, and resolves its exclusive prefix by a batched
//          decoupled look-back over the earlier tiles OF ITS SEGMENT (status words carry a pass
//          epoch, so the table is cleared once per sort);
//       4. keys are placed digit-sorted in shared memory and leave as coalesced digit runs.
//
//  Algorithmic HBM bytes per key: hist 1*W, each sweep pass 2*W (W = 4, 8 or 16 bytes), plus the
//  tile status (8 B per digit and tile, written twice, read >= once).  Measured bound (ncu, profiles/): not DRAM -- the
//  shared-memory pipe (about 510 wavefronts per 1024 keys) together with the issue slots (41 instructions per key and pass).

#include <stdlib.h>
#include <string.h>
#include <algorithm>

#include "kernels.cuh"

namespace {

const int SORT_THREADS = 256;
const int SORT_MAX_PASSES = 24;
#ifndef MB_SWEEP_PREFETCH_DEFAULT
#define MB_SWEEP_PREFETCH_DEFAULT 444    //  tiles ahead = one wave of resident CTAs on 148 SMs (measured at 8 Gbp: 0 -> 131.1 ms, 444 -> 124.2 ms, 1332 -> 136.0 ms)
#endif

template <int KW> struct SortShape;
//  keys per thread of the 256-thread tile that the tile table, the histogram and the count kernels share
template <> struct SortShape<0> { static const int ITEMS = 32; };   //  u32 suffix keys: 8192 per tile
template <> struct SortShape<1> { static const int ITEMS = 16; };   //  u64: 4096 per tile
template <> struct SortShape<2> { static const int ITEMS = 8;  };   //  128 bit: 2048 per tile

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
  //  segTile[] may be a slice of a longer table (a group of a pass's segments sorted on its own): tiles are numbered
  //  from the slice's first tile
  const u64 ta    = tile + segTile[0];
  const u32 seg   = find_segment(segTile, nSegs, ta);
  const u64 first = segTile[seg];
  const u64 begin = segStart[seg] + (ta - first) * tileKeys;
  u64 n = segStart[seg + 1] - begin;
  if (n > tileKeys) n = tileKeys;
  TileInfo t;
  t.begin = begin;  t.n = (u32)n;  t.seg = seg | (ta == first ? 0x80000000u : 0u);
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

  u64 tile0 = (u64)blockIdx.x * HIST_TILES_PER_CTA;
  u64 tileEnd = tile0 + HIST_TILES_PER_CTA;
  if (tileEnd > nTiles) tileEnd = nTiles;
  tile0 += segTile[0];  tileEnd += segTile[0];             //  (segTile[] may be a slice of a longer table: absolute tile numbers from here on)

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
    //  left the kernel waiting on HBM latency: 1.3 TB/s measured).  (unrolled == 0: the one-load loop, kept for reference.)
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

//  Joint form (the default since round 2; MB_HIST_SINGLE=1 for the kernel above): adjacent passes are histogrammed together (one
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

  u64 tile0 = (u64)blockIdx.x * HIST_TILES_PER_CTA;
  u64 tileEnd = tile0 + HIST_TILES_PER_CTA;
  if (tileEnd > nTiles) tileEnd = nTiles;
  tile0 += segTile[0];  tileEnd += segTile[0];             //  (segTile[] may be a slice of a longer table: absolute tile numbers from here on)
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

__device__ __forceinline__ void key_zero(u32 &k)        { k = 0; }
__device__ __forceinline__ void key_zero(u64 &k)        { k = 0; }
__device__ __forceinline__ void key_zero(ulonglong2 &k) { k.x = 0; k.y = 0; }


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

template <int KW, int VER, bool FULL, bool BYTE>
__device__ __forceinline__ void
sweepx_tile(const typename KeyT<KW>::type *__restrict__ in, typename KeyT<KW>::type *__restrict__ out,
            u64 tile, u64 begin, u32 nIn, u32 seg, bool isFirst,
            const u64 *__restrict__ gbase, u32 nPasses, u32 pass, u32 shift, u32 bits, bool wide,
            u64 *__restrict__ status, u32 epoch, const SweepXShared &sm) {
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
      if (FULL || j < n) k = pin[i * 32]; else key_zero(k);
      px[i * 32 + (i * 32) / ITEMS] = k;
    }
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

}  //  namespace


//  ---- host side ---------------------------------------------------------------------------------

//  Which sweep shape: 3 = 256 threads x 6-bit digits (default), 2 = 128 threads x 8-bit digits.  MB_SORT=2|3 in the environment,
//  or mb_sort_set_version(); fixed for the lifetime of a plan / workspace.  (Round 1 and 2 measured five more on B200 -- the first
//  generation with warp votes, 256 x 7, 512 x 6, 256 x 6 at 128 registers, a persistent form with next-tile prefetch in
//  registers: all slower than 256 x 6, see profiles/README.md -- and they are gone.)
static int g_sortVersion = 0;
int mb_sort_version(void) {
  if (g_sortVersion == 0) {
    const char *e = getenv("MB_SORT");
    g_sortVersion = (e && (e[0] == '2' || e[0] == '3') && e[1] == 0) ? e[0] - '0' : MB_SORT_DEFAULT_VERSION;
  }
  return g_sortVersion;
}
void mb_sort_set_version(int v) { g_sortVersion = (v == 2 || v == 3) ? v : MB_SORT_DEFAULT_VERSION; }
u32  mb_sort_radix_bits(void) { return mb_sort_version() == 2 ? SweepShape<2>::RBITS : SweepShape<3>::RBITS; }
u32  mb_sort_radix(void)        { return 1u << mb_sort_radix_bits(); }

void mb_sort_plan(u32 sortBits, MbSortPlan *plan) {
  const u32 RB = mb_sort_radix_bits();
  const u32 P = (sortBits + RB - 1) / RB;
  plan->nPasses = P;
  plan->maxSegKeys = 0;
  u32 shift = 0;
  for (u32 p = 0; p < P; p++) {
    //  6-bit shape: digits of (almost) equal width.  8-bit shape: whole bytes from the bottom (one PRMT per digit), the
    //  top pass takes what is left.
    u32 b = sortBits / P + (p < sortBits % P ? 1 : 0);
    if (RB == 8) b = (sortBits - shift < 8) ? sortBits - shift : 8;
    plan->shift[p] = shift;
    plan->bits[p]  = b;
    shift += b;
  }
}

u32 mb_sort_tile_keys(u32 keyBytes) {                  //  the same for both sweep kernels: 8192 / 4096 / 2048
  static_assert(SORT_THREADS * SortShape<0>::ITEMS == SweepX<0, 2>::TILE && SweepX<0, 2>::TILE == SweepX<0, 3>::TILE, "tile size");
  static_assert(SORT_THREADS * SortShape<1>::ITEMS == SweepX<1, 2>::TILE && SweepX<1, 2>::TILE == SweepX<1, 3>::TILE, "tile size");
  static_assert(SORT_THREADS * SortShape<2>::ITEMS == SweepX<2, 2>::TILE && SweepX<2, 2>::TILE == SweepX<2, 3>::TILE, "tile size");
  return SORT_THREADS * (keyBytes == 4 ? SortShape<0>::ITEMS : keyBytes == 8 ? SortShape<1>::ITEMS : SortShape<2>::ITEMS);
}

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
                                                                     1u);
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
  if ((e = cudaFuncSetAttribute(sortx_sweep_kernel<KW, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SweepX<KW, 2>::BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(sortx_sweep_kernel<KW, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SweepX<KW, 3>::BYTES)) != cudaSuccess) return e;
  //  MB_SWEEP_PREFETCH=<tiles>: L2 prefetch distance of the sweep kernel (0 = off).  (A copy-out by digit runs -- a warp per
  //  chunk of one digit's run, no per-key digit extraction and offset lookup -- measured 14 % SLOWER, 149.0 vs 131.1 ms, and is gone.)
  const u32 pfDist = getenv("MB_SWEEP_PREFETCH") ? (u32)strtoul(getenv("MB_SWEEP_PREFETCH"), nullptr, 10) : MB_SWEEP_PREFETCH_DEFAULT;
  for (u32 p = 0; p < plan.nPasses; p++) {
#define MB_SX_LAUNCH(V) do { typedef SweepX<KW, V> SX;                                                                                         \
      sortx_sweep_kernel<KW, V><<<(unsigned)nTiles, SX::THREADS, SX::BYTES, s>>>(src, dst, info, ghist, plan.nPasses, p,                       \
                                                                     plan.shift[p], plan.bits[p], wide, status, counter + p, p + 1, nTiles, pfDist); } while (0)
    if (version == 2) MB_SX_LAUNCH(2);
    else              MB_SX_LAUNCH(3);
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
