//  Device-side helpers and the launch entry points of the sm_100a kernels.
//
//  HBM data layout (all little-endian words, all arrays dense):
//
//   packed read store   one 12-byte record per 32 consecutive positions of the base stream:
//                         u32 basesHi, u32 basesLo   2 bits per base, position j of the record in
//                                                    bits [63-2j, 62-2j] of (basesHi:basesLo)
//                                                    (A=0 C=1 T=2 G=3, i.e. (ascii>>1)&3)
//                         u32 valid                  bit 31-j set iff position j holds ACGTacgt
//                       record breaks ('.'), N, IUPAC ... are simply invalid positions.
//   k-mer keys          k <= 32: one u64 per k-mer (2k low bits);  k > 32: ulonglong2 {lo, hi}.
//                       First base is the most significant 2 bits (kmerTiny, kmers-tiny.H:86).
//   pass buffers        keysA / keysB (ping-pong for the LSD sort), segmented by prefix bin.
//   counted k-mers      distinct keys compacted in place at the front of the sorted buffer,
//                       runStart[j] = index of the first instance of distinct key j (u64),
//                       runStart[D] = number of instances; count_j = runStart[j+1]-runStart[j].
//   encoded db          byte stream of serialised stuffedBits blocks in ascending block-prefix
//                       order (files are contiguous ranges of it) + 24-byte index entries.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long u64;
typedef unsigned int       u32;

#ifndef MB_MODE_CANONICAL                //  same values as include/meryl_b200.h
#define MB_MODE_CANONICAL 0
#define MB_MODE_FORWARD   1
#define MB_MODE_REVERSE   2
#endif

#define MB_STORE_PAD_RECORDS 1024       //  zero records that must follow the last record of the store
#define MB_STORE_REC_WORDS   3          //  u32 words per 32-position record
#define MB_MAX_BINS          1024       //  2^10 prefix bins
#define MB_RADIX_BITS        6          //  LSD digit width: 64 digits = 32 words of two 16-bit private counters per thread
#define MB_RADIX             64
#define MB_SORT_DEFAULT_VERSION 3       //  sweep kernel used unless MB_SORT / mb_sort_set_version says otherwise (sort.cu)

//  ---- key helpers -----------------------------------------------------------------------

//  KW = key words: 1 -> u64 (k <= 32), 2 -> 128 bits (k <= 64), 0 -> u32 SUFFIX keys: when the bits
//  below the prefix bin fit 32 bits (2k - binBits <= 32, i.e. k <= 21) a pass stores only those;
//  the bin is implied by the segment the key sits in.  Halves the sort traffic and doubles the
//  number of k-mers per pass for the common k = 21.
template <int KW> struct KeyT;
template <> struct KeyT<0> { typedef u32        type; };
template <> struct KeyT<1> { typedef u64        type; };
template <> struct KeyT<2> { typedef ulonglong2 type; };     //  x = low 64 bits, y = high bits

__host__ __device__ __forceinline__ u64 key_lo(u32 k)               { return k;   }
__host__ __device__ __forceinline__ u64 key_hi(u32)                 { return 0;   }
__host__ __device__ __forceinline__ bool key_eq(u32 a, u32 b)       { return a == b; }
__host__ __device__ __forceinline__ u32 key_field(u32 k, u32 shift, u32 bits) {      //  shift < 32 (a field of a 32-bit key)
  return (k >> shift) & ((1u << bits) - 1);
}
__host__ __device__ __forceinline__ u64 key_shr64(u32 k, u32 shift) { return shift >= 32 ? 0 : (u64)(k >> shift); }

__host__ __device__ __forceinline__ u64 key_lo(u64 k)               { return k;   }
__host__ __device__ __forceinline__ u64 key_hi(u64)                 { return 0;   }
__host__ __device__ __forceinline__ u64 key_lo(const ulonglong2 &k) { return k.x; }
__host__ __device__ __forceinline__ u64 key_hi(const ulonglong2 &k) { return k.y; }

__host__ __device__ __forceinline__ bool key_eq(u64 a, u64 b)                             { return a == b; }
__host__ __device__ __forceinline__ bool key_eq(const ulonglong2 &a, const ulonglong2 &b) { return a.x == b.x && a.y == b.y; }

//  (key >> shift) & ((1<<bits)-1), bits <= 32, the field may straddle the two words.
__host__ __device__ __forceinline__ u32 key_field(u64 k, u32 shift, u32 bits) {
  return (u32)(k >> shift) & ((1u << bits) - 1);
}
__host__ __device__ __forceinline__ u32 key_field(const ulonglong2 &k, u32 shift, u32 bits) {
  u64 v;
  if      (shift >= 64)  v = k.y >> (shift - 64);
  else if (shift == 0)   v = k.x;
  else                   v = (k.x >> shift) | (k.y << (64 - shift));
  return (u32)v & ((1u << bits) - 1);
}

//  key >> shift as a u64 (caller guarantees the result fits, i.e. 2k - shift <= 64).
__host__ __device__ __forceinline__ u64 key_shr64(u64 k, u32 shift) { return shift >= 64 ? 0 : k >> shift; }
__host__ __device__ __forceinline__ u64 key_shr64(const ulonglong2 &k, u32 shift) {
  if (shift >= 64) return k.y >> (shift - 64);
  if (shift == 0)  return k.x;
  return (k.x >> shift) | (k.y << (64 - shift));
}

//  ---- decoupled look-back status words --------------------------------------------------
//  bits 63..62: 0 = not ready, 1 = tile aggregate, 2 = inclusive prefix; low 62 bits: value.
#define MB_ST_AGG   (1ull << 62)
#define MB_ST_INC   (2ull << 62)
#define MB_ST_MASK  ((1ull << 62) - 1)

//  Hide a pointer's provenance from the optimiser (it otherwise re-associates base + index arithmetic).
#ifdef MB_CUSIM
#define MB_OPAQUE_PTR(p) do { } while (0)
#else
#define MB_OPAQUE_PTR(p) asm volatile("" : "+l"(p))
#endif

//  One 128-byte line on its way from HBM to L2 (no register, no wait): what a CTA does for the tile a LATER CTA will load.
#ifdef MB_CUSIM
__device__ __forceinline__ void prefetch_l2(const void *) { }
#else
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
#endif

#ifdef MB_CUSIM   //  host emulator used by the CPU tests (tests/cusim): a polling load lets the other fibers run
__device__ __forceinline__ u64  ld_volatile_u64(const u64 *p)  { cusim::yield(); return *(const volatile u64 *)p; }
__device__ __forceinline__ void st_volatile_u64(u64 *p, u64 v) { *(volatile u64 *)p = v; cusim::g_progress++; }
#else
__device__ __forceinline__ u64 ld_volatile_u64(const u64 *p) {
  u64 v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(u64 *p, u64 v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
#endif

//  Decoupled look-back by one warp: exclusive prefix of the aggregates of tiles [0, tile).  Every round trip to
//  L2 fetches the 32 status words below the current position; the walk consumes the published ones up to and
//  including the first inclusive prefix.  (A single thread walking one word per round trip is latency-bound:
//  with ~1000 tiles in flight the chain of aggregates to cross is hundreds of tiles long.)
//  Call with all 32 lanes of a warp; returns the same value in every lane.
__device__ __forceinline__ u64 lookback_warp(const u64 *__restrict__ status, u64 tile, u32 lane) {
  u64 excl = 0;
  u64 top  = tile;                                           //  tiles [0, top) are still to be accounted for
  while (top > 0) {
    const bool have = top > (u64)lane;                       //  my word: tile top - 1 - lane
    const u64  v    = have ? ld_volatile_u64(status + (top - 1 - lane)) : MB_ST_INC;   //  below tile 0: an inclusive prefix of 0
    const u32  flag = (u32)(v >> 62);
    const u32  ready = __ballot_sync(0xffffffffu, flag != 0);
    const u32  inc   = __ballot_sync(0xffffffffu, flag == 2);
    const u32  nReady = (ready == 0xffffffffu) ? 32u : (u32)(__ffs((int)~ready) - 1);   //  published words from lane 0 on
    if (nReady == 0) continue;                               //  the nearest predecessor has not published yet: poll
    u32 take = nReady;                                       //  stop behind the first inclusive prefix among them
    const u32 incIn = inc & (nReady == 32 ? 0xffffffffu : ((1u << nReady) - 1));
    if (incIn) take = (u32)__ffs((int)incIn);
    u64 c = (lane < take) ? (v & MB_ST_MASK) : 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    excl += c;
    if (incIn) break;
    top -= take;
  }
  return excl;
}

//  ---- launch entry points (each returns cudaGetLastError()) ----------------------------

//  K1  pack.cu: ASCII work-buffer bytes -> store records.  nBytes is padded by the kernel with
//  invalid positions up to a multiple of 32; firstRecord = record index of dAscii[0].
cudaError_t mb_launch_pack(const uint8_t *dAscii, u64 nBytes, u32 *dStore, u64 firstRecord, cudaStream_t s);

struct MbExtractParams {
  const u32 *store;          //  packed read store (records), padded with >= 4 zero records
  u64        nPositions;     //  positions in the store (multiple of 32)
  u32        k;
  u32        mode;           //  MB_MODE_*
  u32        binBits;        //  bins = top binBits bits of the k-mer
  u32        binLo, binHi;   //  this pass keeps bins in [binLo, binHi)
  u32        suffix32;       //  1: store only the low 2k-binBits (<= 32) bits of each k-mer as u32
  u32        noPrefilter;    //  1: never use the bit-sliced file-range filter (debug / A-B; env MB_NO_PREFILTER)
  u32        keptPerMille;   //  share of the store's k-mers this pass keeps, in 1/1000 (0 = unknown)
};

//  K2a extract.cu: bin histogram of every k-mer in the store (count-only pre-pass).  dBinCount[MB_MAX_BINS] += ...
cudaError_t mb_launch_bin_histogram(const MbExtractParams &p, u64 *dBinCount, cudaStream_t s);

//  K2/K3 extract.cu: canonical k-mers of the bins in [binLo,binHi) scattered into their bin
//  segments: key of bin b lands at dKeys[dBinStart[b] + (atomic cursor of b)].
//  dBinCursor[MB_MAX_BINS] must be zero on entry.
cudaError_t mb_launch_extract(const MbExtractParams &p, void *dKeys, const u64 *dBinStart, u64 *dBinCursor, cudaStream_t s);

//  K4  sort.cu: segmented LSD onesweep.  Segment s holds keys [dSegStart[s], dSegStart[s+1]); its
//  tiles are [dSegTile[s], dSegTile[s+1]).  Sorts every segment on key bits [0, sortBits).
//  Result is left in dKeysA if the number of digit passes is even, else in dKeysB; *resultInB says which.
struct MbSortPlan {
  u32  nPasses;
  u32  shift[24];
  u32  bits[24];
  u64  maxSegKeys;             //  longest segment, if the caller knows it (0 = unknown): < 2^32 enables 32-bit output offsets
};
int         mb_sort_version(void);                          //  1: sort_sweep_kernel (6-bit digits); 2..4: sortx_sweep_kernel shapes (8 / 6 / 7 bits); env MB_SORT
void        mb_sort_set_version(int v);
u32         mb_sort_radix_bits(void);
u32         mb_sort_radix(void);
void        mb_sort_plan(u32 sortBits, MbSortPlan *plan);
u32         mb_sort_tile_keys(u32 keyBytes);                //  keyBytes: 4 (u32 suffix keys), 8 or 16
size_t      mb_sort_workspace_bytes(u64 nTiles, u32 nSegs, u32 nPasses);
const void *mb_sort_tile_info(void *dWorkspace, u64 nTiles, u32 nSegs, u32 nPasses);   //  device TileInfo table left by mb_launch_sort
cudaError_t mb_launch_sort(u32 keyBytes, void *dKeysA, void *dKeysB, u64 nKeys,
                           const u64 *dSegStart, const u64 *dSegTile, u32 nSegs, u64 nTiles,
                           const MbSortPlan &plan, void *dWorkspace, size_t workspaceBytes,
                           int *resultInB, u32 *nLaunches, cudaStream_t s,
                           cudaEvent_t evAfterHist /* may be NULL */);

//  K5  count.cu: run-length count of a globally sorted key array, compacting in place.
//  dRunStart needs nKeys+1 entries; *dNumDistinct receives D.  Also accumulates the value histogram:
//  dHistSmall[MB_HIST_SMALL] (u64), values >= MB_HIST_SMALL appended to dHistBig (u32 values) through dHistBigN.
#define MB_HIST_SMALL  (1u << 20)
size_t      mb_count_workspace_bytes(u64 nKeys);
cudaError_t mb_launch_count(u32 keyWords, void *dKeys, u64 nKeys, u64 *dRunStart, u64 *dNumDistinct,
                            void *dWorkspace, cudaStream_t s);
//  u32 suffix keys: run-length count of the sort tiles [tileLo, tileHi) (whole segments), expanding
//  every distinct key to u64 = (binLo + segment) << suffixBits | suffix.  dUniq / dRunStart receive the
//  group's distinct keys and the pass-array index of each run's first instance (+ a sentinel).
cudaError_t mb_launch_count_suffix32(const u32 *dKeys, const void *dTileInfo, u64 tileLo, u64 tileHi, u64 endIndex,
                                     u32 binLo, u32 suffixBits, u64 *dUniq, u64 *dRunStart, u64 *dNumDistinct,
                                     void *dWorkspace, cudaStream_t s);
cudaError_t mb_launch_lookback_test(const u64 *dStatus, u64 tile, u64 *dOut, cudaStream_t s);     //  test hook (debug.cu)
cudaError_t mb_launch_value_histogram(const u64 *dRunStart, const u64 *dNumDistinct, u64 nMaxDistinct,
                                      u64 *dHistSmall, u32 *dHistBig, u64 *dHistBigN, u64 histBigCap,
                                      u64 *dTotals /* [3]: unique, distinct, total */, cudaStream_t s);
//  Same with explicit values (set operations): value j = dCounts[j], j < *dNumDistinct.
cudaError_t mb_launch_value_histogram_counts(const u32 *dCounts, const u64 *dNumDistinct, u64 nMaxDistinct,
                                             u64 *dHistSmall, u32 *dHistBig, u64 *dHistBigN, u64 histBigCap,
                                             u64 *dTotals, cudaStream_t s);

//  K6  encode.cu
struct MbBlockMeta {           //  one per block of the pass, written by the layout kernel
  u64 first;                   //  index of the block's first distinct k-mer
  u64 outByte;                 //  byte offset of the serialised block in the pass output (after the scan)
  u32 nSub;                    //  sub-blocks
  u32 subFirst;                //  index of its first entry in the sub-block word table
  u64 bytes;                   //  serialised size
};
struct MbSplitPart {           //  one of the consecutive blocks an over-full stream-writer prefix is written as
  u64 relByte;                 //  its place after the first part's, in bytes ...
  u32 relSub;                  //  ... and in the sub-block word table
  u32 nSub;
};
struct MbEncodeParams {
  const void *keys;            //  distinct keys of the pass (sorted)
  const u64  *runStart;
  u64         nDistinct;
  u32         keyWords;
  u32         k, wPrefix;
  u64         blockLo;         //  first block prefix of the pass
  u64         nBlocks;         //  number of consecutive block prefixes in the pass
  const u32  *counts;          //  NULL: value j = runStart[j+1] - runStart[j];  else value j = counts[j] (set operations)
  u32         streamMode;      //  1: merylStreamWriter rules (kmers-writer-stream.C:89-157): a prefix without k-mers writes
                               //     no block and leaves its index entry {0, UINT64_MAX, 0}; a prefix with more than
                               //     streamBlockMax k-mers is written as consecutive blocks of streamBlockMax k-mers (the last
                               //     one shorter) under ONE index entry {prefix, place of the first, total k-mers}
                               //     (merylStreamWriter::addMer :131-141, merylFileIndex::set, kmers-files.H:116-128)
  u64         streamBlockMax = 0;      //  16 Mi in the reference (_batchMaxKmers)
  MbSplitPart *parts = nullptr;        //  streamMode: 2 * (nDistinct / streamBlockMax) + 4 slots; part q of the block whose first
                                       //     k-mer is `first` lives in slot 2 * (first / streamBlockMax) + q  (two over-full
                                       //     blocks are more than streamBlockMax k-mers apart, so their slots never meet)
};
//  E0+E1 for the whole pass: per-block extent and layout size into dMeta[nBlocks]; an exclusive scan
//  then fills outByte/subFirst (relative to the pass).  dTotals[0] = total bytes, dTotals[1] = total
//  sub-blocks, dTotals[2..3] = [first, end) distinct k-mer covered by the pass's blocks.
cudaError_t mb_launch_encode_layout(const MbEncodeParams &p, MbBlockMeta *dMeta, u64 *dTotals, void *dScanTmp, cudaStream_t s);
size_t      mb_encode_scan_tmp_bytes(u64 nBlocks);
//  E2+E3 for the blocks [chunkFirst, chunkFirst+chunkCount) of the pass: container headers, sub-block
//  tables, block headers, index entries, then the per-k-mer bit scatter.  The chunk's bytes go to
//  dOut[outByte - chunkByte0] (dOut zeroed by the caller), its sub-block word table to
//  dSubWord[subFirst - chunkSub0], its index entries {prefix, byte offset within its data file, nKmers}
//  to dIndex[3*(blk-chunkFirst)]; the file offset is outByte - (outByte of the file's first block in
//  this pass) + dFileBase[file].
struct MbEncodeChunk {
  u64 chunkFirst, chunkCount;
  u64 chunkByte0, chunkSub0;
};
cudaError_t mb_launch_encode_write(const MbEncodeParams &p, const MbEncodeChunk &c, const MbBlockMeta *dMeta, const u64 *dTotals,
                                   u32 *dSubWord, uint8_t *dOut, u64 *dIndex, const u64 *dFileBase, cudaStream_t s);
