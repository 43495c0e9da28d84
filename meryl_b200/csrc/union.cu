//  `meryl union-sum A B ... output C` on the device (BASELINE.json configs[4]).
//
//  Replaces, for this one operator, the reference's streaming set-operation path:
//    merylFileReader / merylFileBlockReader   src/utility/src/kmers-v1/kmers-reader.C:456-531,
//                                             src/utility/src/kmers-v1/kmers-files.C:25-128     (read + decode blocks)
//    merylOperation::nextMer, findSumCount    src/meryl/merylOp-nextMer.C:419-683, :44-48       (k-way merge, uint32 sum)
//    merylStreamWriter                        src/utility/src/kmers-v1/kmers-writer-stream.C:89-157 (block per prefix, prefix 12)
//
//  The reference walks 64 files with 64 threads, one k-mer at a time.  Here the unit of work is a GROUP of
//  consecutive data files that fits the device:
//    host     reads the group's .merylData bytes of every input and walks the block containers (sizes only)
//    K7       decode_kernel      one thread per block: Elias-Fano fields -> full k-mers + values, at the block's
//                                exclusive k-mer offset (a block is a strictly sequential bit stream: every unary
//                                code must be decoded to find the next field; blocks are the parallelism)
//    K8       merge_op_kernel    merge-path union of two sorted k-mer lists with the counts of equal k-mers added
//                                (uint32, wraps; zero sums are dropped), compacted by a chained scan; more than two
//                                inputs fold left to right
//    K5'/K6   value histogram and the stuffedBits encoder of the count path, in stream-writer mode
//  Output bytes go straight to the 129 files.  HBM bytes per input k-mer: ~(b+34)/8 read, 2(W+4) written+read by the
//  merge, (W+4) read by the encoder.  Bound: HBM bandwidth (decode: latency of one dependent chain per thread).

#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/stat.h>
#include <algorithm>
#include <atomic>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "common.h"
#include "kernels.cuh"
#include "filesink.h"

namespace {

//  ---- K7: block decode ------------------------------------------------------------------------------

struct __align__(16) UnionBlock {       //  one serialised block of an input file (host-made table)
  u64 byteOff;                          //  of the container, in the group's byte buffer of that input (8-aligned)
  u64 outIdx;                           //  exclusive number of k-mers before it
};

//  Sequential MSB-first reader over the sub-blocks of one stuffedBits container (stuffedBits-v1.H:246-285:
//  a field lies wholly inside one sub-block; it is read from the next one iff it does not fit the rest of this one).
struct FetchGlobal {                    //  words straight from global memory (one reader per thread)
  __device__ __forceinline__ u64 fetch(const u64 *p) { return *p; }
};

//  A warp's sliding window over the container's words in shared memory: all lanes walk the same bit stream in lock step
//  (every call has the same p in every lane), the window is refilled by one coalesced load whenever the walk leaves it.
const int DW_WINDOW = 64;               //  words
struct FetchWindow {
  u64       *win;                       //  DW_WINDOW words of shared memory, this warp's
  const u64 *first;                     //  win[i] == first[i]
  const u64 *end;                       //  end of the input buffer
  u32        lane;
  __device__ __forceinline__ u64 fetch(const u64 *p) {
    if (p < first || p >= first + DW_WINDOW) {
      __syncwarp();
      first = p;
      for (int t = (int)lane; t < DW_WINDOW; t += 32) win[t] = (p + t < end) ? p[t] : 0;
      __syncwarp();
    }
    return win[p - first];
  }
};

template <class Fetch>
struct SubReaderT : Fetch {
  const u64 *len;                       //  sub-block lengths in bits
  const u64 *cur;                       //  first word of the current sub-block
  u32        blk, nSub;
  u64        pos;                       //  bit position inside the current sub-block
  u64        lenCur;                    //  len[blk], in a register
  u64        cw, cv;                    //  one-word cache: cv == cur[cw] (every word is fetched from memory once)
  bool       bad;

  //  bytesLeft: what the buffer holds from the container's first byte on; a container that claims more is bad
  //  (its place comes from the block index, nobody has walked the bytes before).
  __device__ __forceinline__ void open(const u64 *container, u64 bytesLeft) {
    nSub = bytesLeft >= 16 ? (u32)container[1] : 0;
    len  = container + 2 + nSub;
    cur  = len + nSub;
    blk  = 0;  pos = 0;  bad = (nSub == 0) || (16 + 16 * (u64)nSub > bytesLeft);
    if (!bad) {
      u64 words = 0;
      for (u32 i = 0; i < nSub; i++) {
        const u64 w = (len[i] + 63) >> 6;
        if (w > (bytesLeft >> 3)) { bad = true; break; }
        words += w;
      }
      if (!bad && 16 + 16 * (u64)nSub + 8 * words > bytesLeft) bad = true;
    }
    lenCur = bad ? 0 : len[0];
    cw = ~0ull;  cv = 0;
  }
  __device__ __forceinline__ u64 word(u64 wd) {
    if (wd != cw) { cv = this->fetch(cur + wd); cw = wd; }
    return cv;
  }
  __device__ __forceinline__ void move(u64 w) {
    if (pos + w <= lenCur) return;
    if (blk + 1 >= nSub) { bad = true; return; }
    cur += (lenCur + 63) >> 6;
    blk++;
    lenCur = len[blk];
    pos = 0;
    cw = ~0ull;
  }
  __device__ __forceinline__ void skip(u32 w) {             //  step over a w-bit field without reading it
    move(w);
    pos += w;
  }
  __device__ __forceinline__ u64 binary(u32 w) {           //  0 < w <= 64
    move(w);
    if (bad) return 0;
    const u64 wd = pos >> 6;
    const u32 fr = 64 - (u32)(pos & 63);
    u64 v;
    if (w <= fr) v = word(wd) >> (fr - w);
    else       { const u64 a = word(wd);  v = (a << (w - fr)) | (word(wd + 1) >> (64 - (w - fr))); }
    if (w < 64) v &= (((u64)1 << w) - 1);
    pos += w;
    return v;
  }
  __device__ __forceinline__ u64 unary(void) {             //  zeros before the next one
    move(1);
    if (bad) return 0;
    u64 v = 0;
    u64 wd = pos >> 6;
    u64 bits = word(wd) << (pos & 63);                      //  bits from pos on, at the top
    u32 have = 64 - (u32)(pos & 63);                        //  of which are valid
    while (bits == 0) {
      v += have;  pos += have;
      if (pos >= lenCur) { bad = true; return 0; }
      wd++;  bits = word(wd);  have = 64;
    }
    const u32 z = (u32)__clzll((long long)bits);
    v += z;  pos += z + 1;
    if (pos > lenCur) bad = true;
    return v;
  }
};
typedef SubReaderT<FetchGlobal> SubReader;

__device__ __forceinline__ void store_kmer(u64 *out, u64 i, u64 prefix, u32 wSuffix, u64 sufHi, u64 sufLo) {
  out[i] = (wSuffix >= 64 ? 0 : (prefix << wSuffix)) | sufLo;      //  k <= 32: 2k <= 64
  (void)sufHi;
}
__device__ __forceinline__ void store_kmer(ulonglong2 *out, u64 i, u64 prefix, u32 wSuffix, u64 sufHi, u64 sufLo) {
  ulonglong2 v;
  v.x = sufLo;  v.y = sufHi;
  if (wSuffix < 64)       { v.x |= prefix << wSuffix;  v.y |= wSuffix ? (prefix >> (64 - wSuffix)) : 0; }
  else if (wSuffix < 128) { v.y |= prefix << (wSuffix - 64); }
  out[i] = v;
}

template <int KW>
__global__ void __launch_bounds__(128)
decode_kernel(const uint8_t *__restrict__ data, u64 dataBytes, const UnionBlock *__restrict__ tab, u64 nBlocks, u32 wSuffix,
              typename KeyT<KW>::type *__restrict__ keys, u32 *__restrict__ vals, u64 capacity, u32 *__restrict__ err) {
  const u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nBlocks) return;
  const u64 at = tab[b].byteOff;
  if (at >= dataBytes || (at & 7)) { atomicOr(err, 1u); return; }
  const u64 *c = (const u64 *)(data + at);
  SubReader r;
  r.open(c, dataBytes - at);
  if (r.bad) { atomicOr(err, 1u); return; }
  const u64 expect = (b + 1 < nBlocks ? tab[b + 1].outIdx : capacity) - tab[b].outIdx;     //  what the block index says

  //  528 header bits (kmers-files.C:36-52)
  const u64 m1 = r.binary(64), m2 = r.binary(64);
  const u64 prefix = r.binary(64), n = r.binary(64);
  const u32 kCode = (u32)r.binary(8);
  r.binary(32);                                             //  unaryBits: implied by the stream
  const u32 bBits = (u32)r.binary(32);
  r.binary(64);
  const u32 cCode = (u32)r.binary(8);
  r.binary(64);  r.binary(64);
  if (r.bad || m1 != 0x7461446c7972656dull || m2 != 0x0a3030656c694661ull || kCode != 1 || cCode != 1 ||
      bBits > wSuffix || n != expect) { atomicOr(err, 2u); return; }

  const u32 ls = bBits <= 64 ? 0 : bBits - 64, rs = bBits <= 64 ? bBits : 64;
  const u64 o  = tab[b].outIdx;
  u64 hi = 0;                                               //  running high part (<= 2^unaryBits)
  for (u64 i = 0; i < n && !r.bad; i++) {
    hi += r.unary();
    const u64 bh = ls ? r.binary(ls) : 0;
    const u64 bl = rs ? r.binary(rs) : 0;
    //  suffix = hi << bBits | binary  (kmers-files.C:97-107), as (sufHi : sufLo)
    u64 sLo, sHi;
    if (bBits >= 64) { sLo = bl;  sHi = bh | (bBits - 64 < 64 ? (hi << (bBits - 64)) : 0); }
    else             { sLo = bl | (bBits ? (hi << bBits) : hi);  sHi = bBits ? (hi >> (64 - bBits)) : 0; }
    store_kmer(keys, o + i, prefix, wSuffix, sHi, sLo);
  }
  for (u64 i = 0; i < n && !r.bad; i++) vals[o + i] = (u32)r.binary(32);
  if (r.bad) atomicOr(err, 4u);
}

//  K7w: one WARP per block, for inputs with few, long blocks (a database written by the stream writer has 4096 blocks
//  of up to 16 Mi k-mers; one thread per block leaves the device empty and walks a dependent chain of global loads).
//  The boundaries of the k-mer fields can only be found one after the other (every unary code must be read to find the
//  next field), so all lanes walk that chain together, in lock step, over a shared-memory window of the stream that is
//  refilled with coalesced loads; at step j lane j keeps the reader's state.  Then the 32 lanes read the binary parts of
//  "their" k-mers (from global memory: the lines are hot) and store 32 consecutive k-mers with one coalesced store.  The
//  32-bit values have closed-form places inside a sub-block and are read lane-strided.
const int DW_WARPS = 4;

template <int KW>
__global__ void __launch_bounds__(DW_WARPS * 32)
decode_warp_kernel(const uint8_t *__restrict__ data, u64 dataBytes, const UnionBlock *__restrict__ tab, u64 nBlocks, u32 wSuffix,
                   typename KeyT<KW>::type *__restrict__ keys, u32 *__restrict__ vals, u64 capacity, u32 *__restrict__ err) {
  __shared__ u64 window[DW_WARPS][DW_WINDOW];
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u64 b = (u64)blockIdx.x * DW_WARPS + warp;
  if (b >= nBlocks) return;                                 //  (whole warps leave)
  const u64 at = tab[b].byteOff;
  if (at >= dataBytes || (at & 7)) { if (lane == 0) atomicOr(err, 1u); return; }
  const u64 *c = (const u64 *)(data + at);

  SubReaderT<FetchWindow> r;
  r.win = window[warp];  r.first = nullptr;  r.end = (const u64 *)(data + (dataBytes & ~7ull));  r.lane = lane;
  r.first = r.end;                                          //  empty window: the first fetch fills it
  r.open(c, dataBytes - at);
  if (r.bad) { if (lane == 0) atomicOr(err, 1u); return; }
  const u64 expect = (b + 1 < nBlocks ? tab[b + 1].outIdx : capacity) - tab[b].outIdx;

  const u64 m1 = r.binary(64), m2 = r.binary(64);
  const u64 prefix = r.binary(64), n = r.binary(64);
  const u32 kCode = (u32)r.binary(8);
  r.binary(32);
  const u32 bBits = (u32)r.binary(32);
  r.binary(64);
  const u32 cCode = (u32)r.binary(8);
  r.binary(64);  r.binary(64);
  if (r.bad || m1 != 0x7461446c7972656dull || m2 != 0x0a3030656c694661ull || kCode != 1 || cCode != 1 ||
      bBits > wSuffix || n != expect) { if (lane == 0) atomicOr(err, 2u); return; }

  const u32 ls = bBits <= 64 ? 0 : bBits - 64, rs = bBits <= 64 ? bBits : 64;
  const u64 o  = tab[b].outIdx;
  u64 hi = 0;
  bool laneBad = false;
  for (u64 i0 = 0; i0 < n && !r.bad; i0 += 32) {
    const u32 m = (u32)(n - i0 < 32 ? n - i0 : 32);
    SubReader mine;                                         //  the reader's state in front of my k-mer's binary fields
    mine.bad = true;  mine.len = r.len;  mine.nSub = r.nSub;  mine.cw = ~0ull;  mine.cv = 0;
    mine.cur = r.cur;  mine.blk = 0;  mine.pos = 0;  mine.lenCur = 0;
    u64 myHi = 0;
    for (u32 j = 0; j < m; j++) {                           //  the chain, all lanes together
      hi += r.unary();
      if (lane == j) { mine.cur = r.cur;  mine.blk = r.blk;  mine.pos = r.pos;  mine.lenCur = r.lenCur;  mine.bad = r.bad;  myHi = hi; }
      if (ls) r.skip(ls);
      if (rs) r.skip(rs);
    }
    if (lane < m) {
      const u64 bh = ls ? mine.binary(ls) : 0;
      const u64 bl = rs ? mine.binary(rs) : 0;
      u64 sLo, sHi;
      if (bBits >= 64) { sLo = bl;  sHi = bh | (bBits - 64 < 64 ? (myHi << (bBits - 64)) : 0); }
      else             { sLo = bl | (bBits ? (myHi << bBits) : myHi);  sHi = bBits ? (myHi >> (64 - bBits)) : 0; }
      if (!mine.bad) store_kmer(keys, o + i0 + lane, prefix, wSuffix, sHi, sLo);
      else           laneBad = true;
    }
  }

  //  values: inside a sub-block they follow each other every 32 bits; a value that does not fit starts the next one
  u64 i = 0;
  while (i < n && !r.bad) {
    const u64 fit = (r.pos + 32 <= r.lenCur) ? (r.lenCur - r.pos) / 32 : 0;
    if (fit == 0) { r.move(32); if (r.pos + 32 > r.lenCur) r.bad = true; continue; }
    const u64 m = fit < n - i ? fit : n - i;
    for (u64 t = lane; t < m; t += 32) {
      const u64 p  = r.pos + 32 * t;
      const u64 wd = p >> 6;
      const u32 fr = 64 - (u32)(p & 63);                    //  bits left in the word from p on
      const u64 a  = r.cur[wd];
      const u64 v  = (32 <= fr) ? (a >> (fr - 32)) : ((a << (32 - fr)) | (r.cur[wd + 1] >> (64 - (32 - fr))));
      vals[o + i + t] = (u32)v;
    }
    r.pos += 32 * m;  i += m;
  }
  if (__any_sync(0xffffffffu, r.bad || laneBad) && lane == 0) atomicOr(err, 4u);
}

//  ---- K8: merge-path union with summed values ------------------------------------------------------------

__device__ __forceinline__ bool key_less(u64 a, u64 b)                              { return a < b; }
__device__ __forceinline__ bool key_less(const ulonglong2 &a, const ulonglong2 &b)  { return a.y < b.y || (a.y == b.y && a.x < b.x); }

__device__ __forceinline__ void make_query(u64 &q, u32 &file, u64 lo, u64 hi, u32 k)        { (void)hi; q = lo; file = (u32)(lo >> (2 * k - 6)) & 63u; }
__device__ __forceinline__ void make_query(ulonglong2 &q, u32 &file, u64 lo, u64 hi, u32 k) {
  q.x = lo;  q.y = hi;
  const u32 sh = 2 * k - 6;                                 //  > 58 here (k > 32)
  file = (u32)(sh >= 64 ? (hi >> (sh - 64)) : ((hi << (64 - sh)) | (lo >> sh))) & 63u;
}

//  ---- K9: exact lookup ----------------------------------------------------------------------------------------
//  merylExactLookup::value / exists (src/utility/src/kmers-v1/kmers-lookup.H:139-146, kmers-exact.C): the value of a k-mer in
//  a database, 0 if it is not there (or its value lies outside [minValue, maxValue], the table's load-time filter).  The
//  reference builds a prefix-indexed table in host memory; here the database's blocks are decoded on the device (K7) into
//  their sorted (k-mer, value) lists, one list per data file, and every query is a binary search in the list of its file
//  (top 6 bits).  One thread per query; queries whose file is not in the resident group are left alone.
template <int KW>
__global__ void __launch_bounds__(256)
lookup_kernel(const typename KeyT<KW>::type *__restrict__ keys, const u32 *__restrict__ vals, const u64 *__restrict__ fileStart /* [65] */,
              u32 fileLo, u32 fileHi, u32 k, const u64 *__restrict__ qLo, const u64 *__restrict__ qHi, u64 nQueries,
              u32 minValue, u32 maxValue, u32 *__restrict__ out) {
  typedef typename KeyT<KW>::type Key;
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nQueries) return;
  Key q;
  u32 file;
  make_query(q, file, qLo[i], KW == 2 ? qHi[i] : 0, k);
  if (file < fileLo || file >= fileHi) return;
  u64 lo = fileStart[file], hi = fileStart[file + 1];      //  first key >= q in [lo, hi)
  while (lo < hi) {
    const u64 mid = (lo + hi) >> 1;
    if (key_less(keys[mid], q)) lo = mid + 1; else hi = mid;
  }
  u32 v = 0;
  if (lo < fileStart[file + 1] && key_eq(keys[lo], q)) { v = vals[lo]; if (v < minValue || v > maxValue) v = 0; }
  out[i] = v;
}

//  ---- K10: existence of a sequence's k-mers ---------------------------------------------------------------------
//  meryl-lookup -existence / -include / -exclude (src/meryl-lookup/existence.C:59-86, include-exclude.C:62-97): for every
//  position of every sequence the k-mer that starts there (kmerIterator: k consecutive ACGTacgt letters, code (c >> 1) & 3,
//  first base in the top bits; kmers-iterator.H:63-103) counts as found when the table holds its forward or its
//  reverse-complement form with a value in [minValue, maxValue].  One thread per position: the sequence it lies in is a
//  binary search in seqStart, both forms are built from the k letters, the smaller one is searched first (K9's search).
//  flag[p]: bit 0 = a k-mer starts at p, bit 1 = found (in this or an earlier group of files), bit 2 = outside the
//  positions that count for `found` (skipFirst: the 10x barcode rule, include-exclude.C:70).
template <int KW>
__device__ __forceinline__ u32 list_value(const typename KeyT<KW>::type *__restrict__ keys, const u32 *__restrict__ vals,
                                          const u64 *__restrict__ fileStart, u32 fileLo, u32 fileHi, u32 k,
                                          u64 qlo, u64 qhi, u32 minValue, u32 maxValue) {
  typename KeyT<KW>::type q;
  u32 file;
  make_query(q, file, qlo, qhi, k);
  if (file < fileLo || file >= fileHi) return 0;
  u64 lo = fileStart[file], hi = fileStart[file + 1];
  const u64 end = hi;
  while (lo < hi) {
    const u64 mid = (lo + hi) >> 1;
    if (key_less(keys[mid], q)) lo = mid + 1; else hi = mid;
  }
  if (lo >= end || !key_eq(keys[lo], q)) return 0;
  const u32 v = vals[lo];
  return (v >= minValue && v <= maxValue) ? v : 0;
}

template <int KW>
__global__ void __launch_bounds__(256)
existence_kernel(const typename KeyT<KW>::type *__restrict__ keys, const u32 *__restrict__ vals, const u64 *__restrict__ fileStart /* [65] */,
                 u32 fileLo, u32 fileHi, u32 k, const unsigned char *__restrict__ bases, u64 nBases,
                 const u64 *__restrict__ seqStart, u64 nSeqs, u32 skipFirst, u32 minValue, u32 maxValue, unsigned char *__restrict__ flag,
                 u32 *__restrict__ values /* NULL, or per position: value(fmer) + value(rmer), once for a palindrome (dump.C:139-157) */) {
  const u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nBases) return;
  unsigned char fl = flag[p];
  if ((fl & 2) && !values) return;                          //  found in an earlier group
  u64 a = 0, b = nSeqs;                                     //  last sequence that starts at or before p
  while (b - a > 1) { const u64 m = (a + b) >> 1; if (seqStart[m] <= p) a = m; else b = m; }
  if (p + k > seqStart[a + 1]) return;
  u64 flo = 0, fhi = 0, rlo = 0, rhi = 0;                   //  forward and reverse-complement k-mer, 2k bits each
  bool ok = true;
  for (u32 i = 0; i < k; i++) {
    const unsigned char c = bases[p + i];
    const unsigned char up = c & 0xdf;
    ok = ok && (up == 'A' || up == 'C' || up == 'G' || up == 'T');
    const u64 code = (c >> 1) & 3;
    fhi = (fhi << 2) | (flo >> 62);  flo = (flo << 2) | code;                //  kmers-tiny.H:86 (addR)
    rlo = (rlo >> 2) | (rhi << 62);  rhi >>= 2;                              //  kmers-tiny.H:87 (addL): complement = code ^ 2
    const u32 sh = 2 * k - 2;
    if (sh >= 64) rhi |= (code ^ 2) << (sh - 64); else rlo |= (code ^ 2) << sh;
  }
  if (!ok) return;
  if (KW == 1) { fhi = 0; rhi = 0; }
  else if (2 * k < 128) { fhi &= (~0ull) >> (128 - 2 * k); }
  if (KW == 1 && 2 * k < 64) flo &= (~0ull) >> (64 - 2 * k);
  fl |= 1;
  if (p - seqStart[a] < skipFirst) fl |= 4;
  else {
    const bool fFirst = fhi < rhi || (fhi == rhi && flo <= rlo);
    const u64 alo = fFirst ? flo : rlo, ahi = fFirst ? fhi : rhi, blo = fFirst ? rlo : flo, bhi = fFirst ? rhi : fhi;
    const bool palindrome = (alo == blo && ahi == bhi);
    const u32 va = list_value<KW>(keys, vals, fileStart, fileLo, fileHi, k, alo, ahi, minValue, maxValue);
    u32 vb = 0;
    if (!palindrome && (values || va == 0)) vb = list_value<KW>(keys, vals, fileStart, fileLo, fileHi, k, blo, bhi, minValue, maxValue);
    if (va | vb) fl |= 2;
    //  merylExactLookup::value hands back what the table stores: the value less (minValue - 1) (kmers-exact.C:57-62,
    //  kmers-lookup.H:405-449 -- value_value() would add it back, value() does not); -wig-count sums exactly that.
    //  (The two forms may lie in different groups of files: the sum builds up.)
    const u32 off = minValue > 1 ? minValue - 1 : 0;
    if (values && (va | vb)) values[p] += (va ? va - off : 0) + (vb ? vb - off : 0);
  }
  flag[p] = fl;
}

//  Per-sequence sums of the flags: a warp whose 32 positions lie in one sequence adds once (ballot + popc), else per lane.
__global__ void __launch_bounds__(256)
existence_sum_kernel(const unsigned char *__restrict__ flag, u64 nBases, const u64 *__restrict__ seqStart, u64 nSeqs,
                     u64 *__restrict__ nTotal, u64 *__restrict__ nFound) {
  const u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u32 lane = threadIdx.x & 31;
  const unsigned char fl = p < nBases ? flag[p] : 0;
  u64 a = 0, b = nSeqs;
  if (p < nBases) while (b - a > 1) { const u64 m = (a + b) >> 1; if (seqStart[m] <= p) a = m; else b = m; }
  const u64 a0 = __shfl_sync(0xffffffffu, a, 0);
  const bool same = __all_sync(0xffffffffu, p >= nBases || a == a0);
  const u32 tot = __ballot_sync(0xffffffffu, (fl & 1) != 0), fnd = __ballot_sync(0xffffffffu, (fl & 2) != 0);
  if (same) {
    if (lane == 0) {
      if (tot) atomicAdd((unsigned long long *)&nTotal[a0], (unsigned long long)__popc(tot));
      if (fnd) atomicAdd((unsigned long long *)&nFound[a0], (unsigned long long)__popc(fnd));
    }
  } else {
    if (fl & 1) atomicAdd((unsigned long long *)&nTotal[a], 1ull);
    if (fl & 2) atomicAdd((unsigned long long *)&nFound[a], 1ull);
  }
}

//  merylExactLookup::nKmers (kmers-lookup.H:136): the k-mers of the database whose value lies in [minValue, maxValue].
__global__ void __launch_bounds__(256)
count_in_range_kernel(const u32 *__restrict__ vals, u64 n, u32 minValue, u32 maxValue, u64 *__restrict__ out) {
  u64 mine = 0;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    const u32 v = vals[i];
    mine += (v >= minValue && v <= maxValue) ? 1 : 0;
  }
  for (int d = 16; d > 0; d >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, d);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd((unsigned long long *)out, (unsigned long long)mine);
}

const int MG_THREADS = 256;
const int MG_ITEMS   = 8;
const int MG_TILE    = MG_THREADS * MG_ITEMS;

//  Merge path: how many elements of A are among the first `diag` elements of merge(A, B), A first on ties.
template <typename Key>
__device__ __forceinline__ u64 merge_path(const Key *a, u64 nA, const Key *b, u64 nB, u64 diag) {
  u64 lo = diag > nB ? diag - nB : 0, hi = diag < nA ? diag : nA;
  while (lo < hi) {
    const u64 mid = (lo + hi) >> 1;                         //  take mid from A, diag - mid from B?
    if (!key_less(b[diag - 1 - mid], a[mid])) lo = mid + 1; //  a[mid] <= b[diag-1-mid]: a[mid] comes first
    else                                      hi = mid;
  }
  return lo;
}

//  One fold step of a set operation: list A (the result so far; input 0 in the first step) against list B (the next
//  input).  keep / value of a k-mer that is only in A, only in B, or in both (merylOp-nextMer.C:559-612, folded left to
//  right: every operator of the family is a left fold of a binary rule):
//    union                 count of inputs holding it: A's value (1 if A is raw input 0) + 1
//    union-min/-max/-sum   min / max / uint32 sum of the values present
//    intersect             in both: A's value (== input 0's)        intersect-min/-max/-sum   in both: min / max / sum
//    subtract              only A: A's value;  both: A - B if that stays positive, else dropped for good;  only B: dropped
//    difference            only A                                    symmetric-difference (two inputs)   only A or only B
//  dropZero: values that are 0 are not written (only sums can wrap to 0; a wrapped intermediate sum of intersect-sum
//  must survive to the next step, so the host sets dropZero in the last step only for that operator).
struct MergeRule { u32 op, firstRaw, dropZero; u64 param, param2; };

__device__ __forceinline__ double rint_half_away(double x) { return floor(x + 0.5); }      //  C round() for x >= 0

//  The one-input operators (merylOp-nextMer.C:497-557): value -> value, 0 = not written.
__device__ __forceinline__ u32 value_rule(u32 op, u64 param, u32 a) {
  switch (op) {
    case MB_OP_LESS_THAN:     return ((u64)a <  param) ? a : 0;
    case MB_OP_GREATER_THAN:  return ((u64)a >  param) ? a : 0;
    case MB_OP_AT_LEAST:      return ((u64)a >= param) ? a : 0;
    case MB_OP_AT_MOST:       return ((u64)a <= param) ? a : 0;
    case MB_OP_EQUAL_TO:      return ((u64)a == param) ? a : 0;
    case MB_OP_NOT_EQUAL_TO:  return ((u64)a != param) ? a : 0;
    case MB_OP_INCREASE:      return (~0ull - a < param) ? ~0u : (u32)(a + param);
    case MB_OP_DECREASE:      return ((u64)a < param) ? 0 : (u32)(a - param);
    case MB_OP_MULTIPLY:      return (a && ~0ull / a < param) ? ~0u : (u32)(a * param);
    case MB_OP_DIVIDE:        return param ? (u32)(a / param) : 0;
    case MB_OP_DIVIDE_ROUND:  return !param ? 0 : ((u64)a < param) ? 1u : (u32)(long long)rint_half_away((double)a / (double)param);
    case MB_OP_MODULO:        return param ? (u32)(a % param) : 0;
    default:                  return a;
  }
}

__device__ __forceinline__ u64 key_low(u64 k)                { return k; }
__device__ __forceinline__ u64 key_low(const ulonglong2 &k)  { return k.x; }

__device__ __forceinline__ bool rule_only_a(const MergeRule &r, u32 va, u64 keyLo, u32 &v) {
  switch (r.op) {
    case MB_OP_KEEP_SUFFIX:  v = va;  return (keyLo & r.param2) == r.param;      //  merylOp-countSimple.C:88-93
    case MB_OP_UNION:        v = r.firstRaw ? 1u : va;  return true;
    case MB_OP_UNION_MIN: case MB_OP_UNION_MAX: case MB_OP_UNION_SUM: case MB_OP_MERGE_BATCHES:
    case MB_OP_SUBTRACT:  case MB_OP_DIFFERENCE: case MB_OP_SYMMETRIC_DIFFERENCE:
                             v = va;  return true;
    case MB_OP_INTERSECT: case MB_OP_INTERSECT_MIN: case MB_OP_INTERSECT_MAX: case MB_OP_INTERSECT_SUM:
                             return false;
    default:                 v = value_rule(r.op, r.param, va);  return true;     //  one-input operators
  }
}
__device__ __forceinline__ bool rule_only_b(const MergeRule &r, u32 vb, u32 &v) {
  switch (r.op) {
    case MB_OP_UNION:        v = 1u;  return true;
    case MB_OP_UNION_MIN: case MB_OP_UNION_MAX: case MB_OP_UNION_SUM: case MB_OP_SYMMETRIC_DIFFERENCE: case MB_OP_MERGE_BATCHES:
                             v = vb;  return true;
    default:                 return false;
  }
}
__device__ __forceinline__ bool rule_both(const MergeRule &r, u32 va, u32 vb, u32 &v) {
  switch (r.op) {
    case MB_OP_UNION:          v = (r.firstRaw ? 1u : va) + 1u;  return true;
    case MB_OP_UNION_MIN: case MB_OP_INTERSECT_MIN:   v = va < vb ? va : vb;  return true;
    case MB_OP_UNION_MAX: case MB_OP_INTERSECT_MAX:   v = va < vb ? vb : va;  return true;
    case MB_OP_UNION_SUM: case MB_OP_INTERSECT_SUM:   v = va + vb;            return true;
    case MB_OP_INTERSECT:      v = va;  return true;
    case MB_OP_MERGE_BATCHES:  v = va + vb;  if (v < vb) v = ~0u;  return true;     //  kmers-writer-block.C:305-309: saturates
    case MB_OP_SUBTRACT:       v = va - vb;  return va > vb;
    default:                   return false;               //  difference, symmetric-difference
  }
}

template <int KW>
__global__ void __launch_bounds__(MG_THREADS)
merge_op_kernel(const typename KeyT<KW>::type *__restrict__ ka, const u32 *__restrict__ va, u64 nA,
                const typename KeyT<KW>::type *__restrict__ kb, const u32 *__restrict__ vb, u64 nB,
                typename KeyT<KW>::type *__restrict__ ko, u32 *__restrict__ vo, u64 *__restrict__ nOut,
                u64 *__restrict__ status, u64 *__restrict__ tileCounter, MergeRule rule) {
  typedef typename KeyT<KW>::type Key;
  __shared__ Key  sKey[MG_TILE + 2];
  __shared__ u32  sVal[MG_TILE + 2];
  __shared__ u64  sPart[2], sTile, sExcl;
  __shared__ u32  warpTot[MG_THREADS / 32];

  const u32 t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) sTile = atomicAdd((unsigned long long *)tileCounter, 1ull);
  __syncthreads();
  const u64 tile  = sTile;
  const u64 total = nA + nB;
  const u64 d0 = tile * MG_TILE, d1 = (d0 + MG_TILE < total) ? d0 + MG_TILE : total;
  if (t < 2) sPart[t] = merge_path(ka, nA, kb, nB, t == 0 ? d0 : d1);
  __syncthreads();
  const u64 aLo = sPart[0], aHi = sPart[1], bLo = d0 - aLo, bHi = d1 - aHi;
  const u32 na = (u32)(aHi - aLo), nb = (u32)(bHi - bLo);

  //  window: [A part | B part] plus, behind each, the element that follows it in its list (partner / guard)
  Key *sa = sKey;            u32 *sva = sVal;
  Key *sb = sKey + na + 1;   u32 *svb = sVal + na + 1;
  for (u32 i = t; i < na; i += MG_THREADS) { sa[i] = ka[aLo + i]; sva[i] = va[aLo + i]; }
  for (u32 i = t; i <= nb; i += MG_THREADS) if (bLo + i < nB) { sb[i] = kb[bLo + i]; svb[i] = vb[bLo + i]; }
  __syncthreads();

  //  my MG_ITEMS consecutive outputs of the tile
  const u32 n  = (u32)(d1 - d0);
  const u32 m0 = min(t * MG_ITEMS, n), m1 = min(m0 + MG_ITEMS, n);
  u32 ai = (u32)merge_path(sa, (u64)na, sb, (u64)nb, (u64)m0), bi = m0 - ai;

  Key oKey[MG_ITEMS];
  u32 oVal[MG_ITEMS];
  u32 keepMask = 0;
  for (u32 m = m0; m < m1; m++) {
    const bool takeA = (bi >= nb) || (ai < na && !key_less(sb[bi], sa[ai]));
    Key k;  u32 v = 0;  bool keep;
    if (takeA) {
      k = sa[ai];                                            //  k-mers are distinct within a database
      //  its partner in B, if any, is the next unconsumed element there (possibly the guard behind the window)
      if (bLo + bi < nB && key_eq(sb[bi], k)) keep = rule_both(rule, sva[ai], svb[bi], v);
      else                                    keep = rule_only_a(rule, sva[ai], key_low(k), v);
      ai++;
    } else {
      k = sb[bi];
      //  already handled with its partner iff A holds the same k-mer: it would be the last element of A before this point
      bool dup = false;
      if (ai > 0)        dup = key_eq(sa[ai - 1], k);
      else if (aLo > 0)  dup = key_eq(ka[aLo - 1], k);
      keep = !dup && rule_only_b(rule, svb[bi], v);
      bi++;
    }
    const u32 j = m - m0;
    oKey[j] = k;  oVal[j] = v;
    if (keep && !(rule.dropZero && v == 0)) keepMask |= 1u << j;   //  merylOp-nextMer.C:640-641: a zero value is not output
  }

  //  compaction: exclusive position of my kept elements in the tile, then in the output (chained scan)
  const u32 mine = __popc(keepMask);
  u32 incl = mine;
  for (int o = 1; o < 32; o <<= 1) {
    const u32 x = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (u32)o) incl += x;
  }
  if (lane == 31) warpTot[warp] = incl;
  __syncthreads();
  u32 wOff = 0, tileTot = 0;
  for (u32 w = 0; w < MG_THREADS / 32; w++) { const u32 c = warpTot[w]; if (w < warp) wOff += c; tileTot += c; }

  if (warp == 0) {
    if (lane == 0) st_volatile_u64(status + tile, (tile == 0 ? MB_ST_INC : MB_ST_AGG) | tileTot);
    const u64 excl = lookback_warp(status, tile, lane);
    if (lane == 0) {
      if (tile != 0) st_volatile_u64(status + tile, MB_ST_INC | (excl + tileTot));
      sExcl = excl;
      if (d1 == total) *nOut = excl + tileTot;
    }
  }
  __syncthreads();

  u64 o = sExcl + wOff + (incl - mine);
  for (u32 j = 0; j < MG_ITEMS; j++)
    if (keepMask & (1u << j)) { ko[o] = oKey[j]; vo[o] = oVal[j]; o++; }
}

//  Drop zero values of a single list (a lone input still goes through `_value == 0` -> skip).
template <int KW>
__global__ void __launch_bounds__(256)
count_nonzero_kernel(const u32 *__restrict__ v, u64 n, u64 *__restrict__ out) {
  u64 c = 0;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) c += (v[i] != 0);
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd((unsigned long long *)out, (unsigned long long)c);
}


//  ---- host: stuffedBits reader for the master index, file helpers ------------------------------------------

struct HostBits {                                           //  sequential reader over a serialised stuffedBits
  const u64 *len = nullptr, *cur = nullptr;
  u32 blk = 0, nSub = 0;
  u64 pos = 0;
  bool open(const uint8_t *d, size_t n) {
    if (n < 16) return false;
    memcpy(&nSub, d + 8, 4);
    if (nSub == 0 || n < 16 + 16 * (size_t)nSub) return false;
    len = (const u64 *)(d + 16) + nSub;
    cur = len + nSub;
    size_t words = 0;
    for (u32 i = 0; i < nSub; i++) words += (len[i] + 63) / 64;
    return n >= 16 + 16 * (size_t)nSub + 8 * words;
  }
  u64 get(u32 w) {
    if (pos + w > len[blk]) { cur += (len[blk] + 63) / 64; blk++; pos = 0; }
    const u64 wd = pos >> 6;  const u32 fr = 64 - (u32)(pos & 63);
    u64 v = (w <= fr) ? (cur[wd] >> (fr - w)) : ((cur[wd] << (w - fr)) | (cur[wd + 1] >> (64 - (w - fr))));
    if (w < 64) v &= (((u64)1 << w) - 1);
    pos += w;
    return v;
  }
};

int readWhole(const std::string &path, std::vector<uint8_t> &out) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return mb_fail(MB_EIO, "cannot open '%s': %s", path.c_str(), strerror(errno));
  struct stat st;
  if (fstat(fileno(f), &st) != 0) { fclose(f); return mb_fail(MB_EIO, "cannot stat '%s'", path.c_str()); }
  out.resize((size_t)st.st_size);
  if (st.st_size && fread(out.data(), 1, out.size(), f) != out.size()) { fclose(f); return mb_fail(MB_EIO, "short read from '%s'", path.c_str()); }
  fclose(f);
  return MB_OK;
}

int writeFile(const std::string &path, const void *p, size_t n) {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) return mb_fail(MB_EIO, "cannot create '%s': %s", path.c_str(), strerror(errno));
  if (n && fwrite(p, 1, n, f) != n) { fclose(f); return mb_fail(MB_EIO, "short write to '%s'", path.c_str()); }
  if (fclose(f) != 0) return mb_fail(MB_EIO, "cannot close '%s'", path.c_str());
  return MB_OK;
}

//  Run job(0..n-1) on a few host threads (file reads / writes only: no CUDA calls inside).
template <typename F>
void parallelFor(size_t n, F job) {
  unsigned nt = std::thread::hardware_concurrency();
  if (nt > 16) nt = 16;
  if (nt < 1)  nt = 1;
  if (nt > n)  nt = (unsigned)n;
  if (nt <= 1) { for (size_t i = 0; i < n; i++) job(i); return; }
  std::atomic<size_t> next(0);
  std::vector<std::thread> th;
  for (unsigned w = 0; w < nt; w++)
    th.emplace_back([&]() { for (size_t i = next++; i < n; i = next++) job(i); });
  for (auto &x : th) x.join();
}

struct HostBuf {                                            //  grow-only, never zero-filled
  uint8_t *p = nullptr;  size_t cap = 0;
  HostBuf() {}
  HostBuf(const HostBuf &) = delete;
  HostBuf &operator=(const HostBuf &) = delete;
  ~HostBuf() { free(p); }
  bool need(size_t n) { if (n > cap) { free(p); p = (uint8_t *)malloc(n ? n : 1); cap = p ? n : 0; } return p != nullptr; }
  uint8_t *data(void) { return p; }
};

//  File bytes travel through a few page-locked staging buffers: host threads fill (or drain) one buffer while the DMA
//  engine moves another, so neither a pageable copy nor a whole-group host buffer sits between the files and the device.
struct PinnedRing {
  static const int    N     = 4;
  size_t      BYTES      = (size_t)32 << 20;          //  (MB_RING_BYTES: testing knob, so that small inputs take many chunks)
  size_t      PIECE      = (size_t)4 << 20;           //  what one host thread reads or writes at a time
  uint8_t    *buf[N]     = { nullptr, nullptr, nullptr, nullptr };
  cudaEvent_t ev[N]      = { nullptr, nullptr, nullptr, nullptr };
  bool        pending[N] = { false, false, false, false };
  int init(void) {
    if (getenv("MB_RING_BYTES")) { BYTES = std::max<size_t>(4096, strtoull(getenv("MB_RING_BYTES"), nullptr, 10) / 4096 * 4096); PIECE = std::max<size_t>(1024, BYTES / 8); }
    for (int i = 0; i < N; i++) {
      if (cudaMallocHost((void **)&buf[i], BYTES) != cudaSuccess) { cudaGetLastError(); buf[i] = nullptr; return mb_fail(MB_ENOMEM, "cannot allocate page-locked staging buffers"); }
      if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) return mb_fail(MB_ENODEVICE, "cudaEventCreate failed");
    }
    return MB_OK;
  }
  void quiesce(void) { for (int i = 0; i < N; i++) if (pending[i]) { cudaEventSynchronize(ev[i]); pending[i] = false; } }
  ~PinnedRing() { for (int i = 0; i < N; i++) { if (ev[i]) cudaEventDestroy(ev[i]); if (buf[i]) cudaFreeHost(buf[i]); } }
};

//  fill(off, len, dst) -> MB_*: produce bytes [off, off + len) of the logical stream into dst (host threads inside).
template <typename Fill>
int streamToDevice(PinnedRing &R, uint8_t *dDst, size_t total, cudaStream_t s, Fill fill) {
  size_t i = 0;
  for (size_t off = 0; off < total; off += R.BYTES, i++) {
    const int b = (int)(i % PinnedRing::N);
    const size_t len = std::min(R.BYTES, total - off);
    if (R.pending[b]) { MB_CUDA(cudaEventSynchronize(R.ev[b])); R.pending[b] = false; }      //  its last copy has left the host
    MB_TRY(fill(off, len, R.buf[b]));
    MB_CUDA(cudaMemcpyAsync(dDst + off, R.buf[b], len, cudaMemcpyHostToDevice, s));
    MB_CUDA(cudaEventRecord(R.ev[b], s));
    R.pending[b] = true;
  }
  return MB_OK;
}

//  consume(off, len, src) -> MB_*: take bytes [off, off + len) of the device range; the copy of the next chunk is in
//  flight while a chunk is consumed.
template <typename Consume>
int streamFromDevice(PinnedRing &R, const uint8_t *dSrc, size_t total, cudaStream_t s, Consume consume) {
  R.quiesce();
  const size_t nChunks = (total + R.BYTES - 1) / R.BYTES;
  for (size_t i = 0; i <= nChunks; i++) {
    if (i < nChunks) {
      const int b = (int)(i % PinnedRing::N);
      const size_t off = i * R.BYTES, len = std::min(R.BYTES, total - off);
      MB_CUDA(cudaMemcpyAsync(R.buf[b], dSrc + off, len, cudaMemcpyDeviceToHost, s));
      MB_CUDA(cudaEventRecord(R.ev[b], s));
    }
    if (i >= 1) {
      const size_t j = i - 1;
      const int b = (int)(j % PinnedRing::N);
      const size_t off = j * R.BYTES, len = std::min(R.BYTES, total - off);
      MB_CUDA(cudaEventSynchronize(R.ev[b]));
      MB_TRY(consume(off, len, (const uint8_t *)R.buf[b]));
    }
  }
  return MB_OK;
}

//  merylStreamWriter keeps at most this many k-mers in a block (_batchMaxKmers, kmers-writer-stream.C:127): a prefix with
//  more is written as consecutive blocks under one index entry.  MB_STREAM_BLOCK_MAX is a testing knob (16 Mi k-mers of
//  one prefix are too many for most emulator runs); the reader side must be told the same value.
u64 streamBlockMax(void) {
  const char *e = getenv("MB_STREAM_BLOCK_MAX");
  const u64 v = e ? strtoull(e, nullptr, 10) : 0;
  return v ? v : ((u64)16 << 20);
}

//  The serialised blocks behind one entry {prefix, place, nKmers} of a block index: one -- or, when the entry counts more
//  k-mers than the stream writer puts in a block, the consecutive blocks of that prefix (merylFileIndex::set adds their
//  k-mers up, kmers-files.H:116-128; the reference's reader just reads block after block, kmers-reader.C).  Their places
//  follow from the container headers: {maxBits, nSub | ..., bgn[nSub], len[nSub], data...}, data word 3 = the block's
//  k-mer count.  fetch(pos, dst, len) reads bytes of the data file.
template <typename Fetch>
int blocksOfEntry(Fetch fetch, u64 at, u64 nEntry, u64 fileBytes, u64 &nk, std::vector<UnionBlock> &tab) {
  UnionBlock ub;
  if (nEntry <= streamBlockMax()) { ub.byteOff = at;  ub.outIdx = nk;  tab.push_back(ub);  nk += nEntry;  return MB_OK; }
  std::vector<u64> len;
  for (u64 left = nEntry; left; ) {
    u64 head[2], nBlock = 0;
    if (at + 16 > fileBytes || !fetch(at, head, 16)) return MB_EFORMAT;
    const u64 nSub = head[1] & 0xffffffffull;
    if (nSub == 0 || nSub > (fileBytes - at) / 16 || at + 16 + 16 * nSub + 32 > fileBytes) return MB_EFORMAT;
    len.resize(nSub);
    if (!fetch(at + 16 + 8 * nSub, len.data(), 8 * nSub) || !fetch(at + 16 + 16 * nSub + 24, &nBlock, 8)) return MB_EFORMAT;
    u64 words = 0;
    for (u64 l : len) words += (l + 63) >> 6;
    if (nBlock == 0 || nBlock > left) return MB_EFORMAT;
    ub.byteOff = at;  ub.outIdx = nk;  tab.push_back(ub);
    nk += nBlock;  left -= nBlock;
    at += 16 + 16 * nSub + 8 * words;
  }
  return MB_OK;
}

struct FileJob {                                            //  one input data file of the group
  u32 input = 0, file = 0;
  int fd = -1;
  size_t off = 0, bytes = 0;                                //  place in the input's byte stream (8-aligned)
  std::vector<UnionBlock> tab;                              //  byteOff relative to the file, outIdx relative to its first k-mer
  u64 nk = 0;
  int rc = MB_OK;
  std::string err;
};

struct InputDb {
  std::string dir;
  const MbDbImages *mem = nullptr;                          //  instead of a directory: images in host memory (count batches)
  u32 k = 0, wPrefix = 0, wSuffix = 0;
  u64 fileBytes[64] = { 0 };                                //  .merylData sizes
  u64 fileKmers[64] = { 0 };                                //  from the per-file block indexes
  std::vector<uint8_t> index[64];                           //  the .merylIndex images: {prefix, byte offset, nKmers} per block
  const std::vector<uint8_t> &indexOf(u32 f) const { return mem ? mem->index[f] : index[f]; }
};

int openDb(const char *dir, InputDb &db) {
  db.dir = dir;
  std::vector<uint8_t> m;
  MB_TRY(readWhole(db.dir + "/merylIndex", m));
  HostBits hb;
  if (!hb.open(m.data(), m.size())) return mb_fail(MB_EFORMAT, "'%s/merylIndex' is not a meryl index", dir);
  const u64 m1 = hb.get(64), m2 = hb.get(64);
  if (m1 != 0x646e496c7972656dull || m2 != 0x33302e765f5f7865ull)
    return mb_fail(MB_EFORMAT, "'%s' is not a meryl v.03 database (kmers-reader.C:41-120)", dir);
  db.wPrefix = (u32)hb.get(32);  db.wSuffix = (u32)hb.get(32);
  const u32 nfb = (u32)hb.get(32), nbb = (u32)hb.get(32);
  const u32 flags = (u32)hb.get(32);
  if (nfb != 6 || nbb + 6 != db.wPrefix || ((db.wPrefix + db.wSuffix) & 1))
    return mb_fail(MB_EFORMAT, "'%s': unexpected index geometry (%u file bits, %u block bits, prefix %u, suffix %u)", dir, nfb, nbb, db.wPrefix, db.wSuffix);
  if (flags & 1) return mb_fail(MB_EINVAL, "'%s' is a multiset database; union-sum of multisets is not supported", dir);
  db.k = (db.wPrefix + db.wSuffix) / 2;
  if (db.k < 3 || db.k > 64) return mb_fail(MB_EFORMAT, "'%s': k=%u", dir, db.k);
  char path[4096];
  for (u32 f = 0; f < 64; f++) {
    struct stat st;
    mb_file_name(path, sizeof(path), dir, f, false);
    if (stat(path, &st) != 0) return mb_fail(MB_EIO, "cannot stat '%s': %s", path, strerror(errno));
    db.fileBytes[f] = (u64)st.st_size;
    std::vector<uint8_t> &ix = db.index[f];
    mb_file_name(path, sizeof(path), dir, f, true);
    MB_TRY(readWhole(path, ix));
    if (ix.size() % 24) return mb_fail(MB_EFORMAT, "'%s' is not a block index", path);
    for (size_t i = 0; i + 24 <= ix.size(); i += 24) { u64 nk; memcpy(&nk, ix.data() + i + 16, 8); db.fileKmers[f] += nk; }
  }
  return MB_OK;
}

int openMem(const MbDbImages *img, u32 k, u32 wPrefix, InputDb &db) {
  db.dir = "(batch in memory)";  db.mem = img;
  db.k = k;  db.wPrefix = wPrefix;  db.wSuffix = 2 * k - wPrefix;
  for (u32 f = 0; f < 64; f++) {
    db.fileBytes[f] = img->dataLen[f];
    const std::vector<uint8_t> &ix = img->index[f];
    for (size_t i = 0; i + 24 <= ix.size(); i += 24) { u64 nk; memcpy(&nk, ix.data() + i + 16, 8); db.fileKmers[f] += nk; }
  }
  return MB_OK;
}

struct Carve {
  uint8_t *base; size_t size, off = 0;
  Carve(uint8_t *b, size_t s) : base(b), size(s) {}
  void *take(size_t bytes) {                            //  256-byte aligned as an ADDRESS (the base may be any pointer)
    const size_t mis = (size_t)((uintptr_t)base & 255);
    const size_t o = (off + mis + 255) / 256 * 256 - mis;
    if (o + bytes > size) return nullptr;
    off = o + bytes;
    return base + o;
  }
};

#define MB_UCUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) { rc = mb_fail(MB_ENODEVICE, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); goto done; } \
  } while (0)
#define MB_UTRY(call)  do { rc = (call); if (rc != MB_OK) goto done; } while (0)

}  //  namespace


//  The statistics of a database as its master index holds them (merylHistogram::load, kmers-histogram.C:97-130): what
//  `meryl histogram` and `meryl statistics` print (merylOp-histogram.C:20-100).  values / occurrences may be NULL to ask for
//  the length only; host only.
extern "C" int mb_db_histogram(const char *dbDir, uint32_t *k, uint64_t *nUnique, uint64_t *nDistinct, uint64_t *nTotal,
                               uint64_t *values, uint64_t *occurrences, uint64_t cap, uint64_t *n) {
  mb_clear_error();
  if (!dbDir || !n) return mb_fail(MB_EINVAL, "mb_db_histogram: NULL argument");
  std::vector<uint8_t> m;
  MB_TRY(readWhole(std::string(dbDir) + "/merylIndex", m));
  HostBits hb;
  if (!hb.open(m.data(), m.size())) return mb_fail(MB_EFORMAT, "'%s/merylIndex' is not a meryl index", dbDir);
  const u64 m1 = hb.get(64), m2 = hb.get(64);
  if (m1 != 0x646e496c7972656dull || m2 != 0x33302e765f5f7865ull) return mb_fail(MB_EFORMAT, "'%s' is not a meryl v.03 database", dbDir);
  const u32 wP = (u32)hb.get(32), wS = (u32)hb.get(32);
  hb.get(32);  hb.get(32);  hb.get(32);                     //  file bits, block bits, flags
  const u64 nu = hb.get(64), nd = hb.get(64), nt = hb.get(64), nv = hb.get(64);
  if (k) *k = (wP + wS) / 2;
  if (nUnique) *nUnique = nu;
  if (nDistinct) *nDistinct = nd;
  if (nTotal) *nTotal = nt;
  *n = nv;
  if (values && occurrences) {
    if (cap < nv) return mb_fail(MB_EINVAL, "mb_db_histogram: the histogram has %llu entries", (unsigned long long)nv);
    for (u64 i = 0; i < nv; i++) { values[i] = hb.get(64); occurrences[i] = hb.get(64); }
  }
  return MB_OK;
}

//  Threshold of a filter given as a share of the input's k-mers (merylOperation::initializeThreshold,
//  src/meryl/merylOp-nextMer.C:65-125): fracDistinct (distinct= / d=) -> the smallest value v such that the k-mers with
//  value <= v are at least fracDistinct * numDistinct; wordFrequency (word-frequency= / f=) -> wordFrequency * numTotal,
//  and it wins if both are given.  A negative argument means "not given".  Reads the histogram of the master index only.
extern "C" int mb_db_threshold(const char *dbDir, double fracDistinct, double wordFrequency, uint64_t *threshold) {
  mb_clear_error();
  if (!dbDir || !threshold) return mb_fail(MB_EINVAL, "mb_db_threshold: NULL argument");
  std::vector<uint8_t> m;
  MB_TRY(readWhole(std::string(dbDir) + "/merylIndex", m));
  HostBits hb;
  if (!hb.open(m.data(), m.size())) return mb_fail(MB_EFORMAT, "'%s/merylIndex' is not a meryl index", dbDir);
  const u64 m1 = hb.get(64), m2 = hb.get(64);
  if (m1 != 0x646e496c7972656dull || m2 != 0x33302e765f5f7865ull) return mb_fail(MB_EFORMAT, "'%s' is not a meryl v.03 database", dbDir);
  for (int i = 0; i < 5; i++) hb.get(32);                   //  prefix, suffix, file bits, block bits, flags
  hb.get(64);                                               //  numUnique
  const u64 nDistinct = hb.get(64), nTotal = hb.get(64), nValues = hb.get(64);
  u64 t = UINT64_MAX;                                       //  merylOp.H:251
  if (fracDistinct >= 0) {
    const u64 target = (u64)(fracDistinct * (double)nDistinct);
    u64 seen = 0;
    for (u64 i = 0; i < nValues; i++) {
      const u64 value = hb.get(64), occ = hb.get(64);
      seen += occ;
      if (seen >= target) { t = value; break; }
    }
  }
  if (wordFrequency >= 0) t = (u64)(wordFrequency * (double)nTotal);
  *threshold = t;
  return MB_OK;
}

//  A database decoded on the device, one group of consecutive data files at a time (the reader side of the set operations
//  without the merge: merylFileReader's loop over blocks, kmers-reader.C:456-531).  setup(db, top, s) carves the caller's
//  own buffers from the bottom of the arena; perGroup(db, g, s) works on a group's sorted (k-mer, value) lists while they are
//  resident (the stream is synchronized after it returns); finish(s) runs after the last group.
struct DecodedGroup {
  const void *dKey = nullptr;           //  u64 or ulonglong2, all files of the group back to back
  const u32  *dVal = nullptr;
  u64         fileStart[65];            //  k-mers of the group before file f (files outside the group are empty ranges)
  const u64  *dFileStart = nullptr;     //  the same on the device
  u32         f0 = 0, f1 = 0;
  u64         nKmers = 0;
};

//  (hold: the whole database must decode as ONE group; its arena and stream are then handed to the caller instead of
//  being released -- the resident table of mb_table_open)
struct DecodeHold { uint8_t *arena = nullptr; cudaStream_t s = nullptr; DecodedGroup G; u32 k = 0; };

template <typename Setup, typename PerGroup, typename Finish>
static int decodeDbInGroups(const char *dbDir, int32_t device, uint64_t workspaceBytes, size_t fixedBytes,
                            Setup setup, PerGroup perGroup, Finish finish, DecodeHold *hold = nullptr) {
  int nDev = 0;
  if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0) return mb_fail(MB_ENODEVICE, "no CUDA device (this library has no CPU path)");
  InputDb db;
  MB_TRY(openDb(dbDir, db));
  const u32 kw = db.k <= 32 ? 1 : 2, W = 8 * kw;

  int rc = MB_OK;
  uint8_t *arena = nullptr;
  cudaStream_t s = nullptr;
  {
    if (device >= 0) MB_UCUDA(cudaSetDevice(device));
    size_t freeB = 0, totalB = 0;
    mb_arena_cache_release(-1);                      //  (a count context's cached arena would otherwise hide most of the device)
    MB_UCUDA(cudaMemGetInfo(&freeB, &totalB));
    auto groupNeed = [&](u32 f0, u32 f1) -> size_t {
      size_t need = (size_t)16 << 20;
      for (u32 f = f0; f < f1; f++) need += db.fileBytes[f] + 256 + (db.index[f].size() / 24 + db.fileKmers[f] / streamBlockMax()) * sizeof(UnionBlock) + db.fileKmers[f] * (W + 4) + 1024;
      return need;
    };
    size_t want = workspaceBytes ? (size_t)workspaceBytes : std::min(freeB > ((size_t)3 << 30) ? freeB - ((size_t)2 << 30) : freeB / 2,
                                                                     fixedBytes + groupNeed(0, 64) + ((size_t)64 << 20));
    if (want > freeB) want = freeB;
    want = want / 256 * 256;
    if (cudaMalloc((void **)&arena, want) != cudaSuccess) { cudaGetLastError(); arena = nullptr; rc = mb_fail(MB_ENOMEM, "cannot allocate a %zu byte device arena", want); goto done; }
    MB_UCUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));

    Carve top(arena, want);
    u64 *dFileStart = (u64 *)top.take(65 * 8);
    u32 *dErr = (u32 *)top.take(256);
    if (!dFileStart || !dErr) { rc = mb_fail(MB_ENOMEM, "device arena too small"); goto done; }
    MB_UCUDA(cudaMemsetAsync(dErr, 0, 256, s));
    MB_UTRY(setup(db, top, s));
    const size_t fixedOff = (top.off + 255) / 256 * 256;   //  the group carve starts here: keep its base aligned

    HostBuf host;
    u32 f0 = 0;
    while (f0 < 64) {
      u32 f1 = f0 + 1;
      if (fixedOff + groupNeed(f0, f1) > want) { rc = mb_fail(MB_ENOMEM, "data file %u needs %zu bytes on the device, %zu available", f0, groupNeed(f0, f1), want - fixedOff); goto done; }
      const u32 maxFiles = (!hold && getenv("MB_UNION_MAX_FILES")) ? (u32)strtoul(getenv("MB_UNION_MAX_FILES"), nullptr, 10) : 64;   //  testing knob
      while (f1 < 64 && f1 - f0 < maxFiles && fixedOff + groupNeed(f0, f1 + 1) <= want) f1++;
      if (hold && f1 < 64) { rc = mb_fail(MB_ENOMEM, "the decoded database needs %zu bytes on the device, %zu available: it cannot stay resident", groupNeed(0, 64), want - fixedOff); goto done; }
      Carve g(arena + fixedOff, want - fixedOff);

      //  the group's bytes and block table
      std::vector<size_t> off(f1 - f0 + 1, 0);
      for (u32 f = f0; f < f1; f++) off[f - f0 + 1] = off[f - f0] + (db.fileBytes[f] + 7) / 8 * 8;
      const size_t rawBytes = off[f1 - f0];
      if (!host.need(rawBytes + 8)) { rc = mb_fail(MB_ENOMEM, "cannot allocate %zu bytes of host memory", rawBytes); goto done; }
      std::atomic<int> bad(-1);
      parallelFor(f1 - f0, [&](size_t fi) {
        char path[4096];
        mb_file_name(path, sizeof(path), db.dir.c_str(), f0 + (u32)fi, false);
        const size_t n = (size_t)db.fileBytes[f0 + fi];
        FILE *fp = fopen(path, "rb");
        if (!fp || (n && fread(host.data() + off[fi], 1, n, fp) != n)) bad = (int)fi;
        if (fp) fclose(fp);
      });
      if (bad >= 0) { rc = mb_fail(MB_EIO, "cannot read data file %u of '%s'", f0 + (u32)bad, dbDir); goto done; }
      std::vector<UnionBlock> tab;
      DecodedGroup G;
      G.f0 = f0;  G.f1 = f1;  G.dFileStart = dFileStart;
      u64 nk = 0;
      for (u32 f = 0; f < 65; f++) {
        G.fileStart[f] = nk;
        if (f < f0 || f >= f1) continue;
        const std::vector<uint8_t> &ix = db.index[f];
        for (size_t e = 0; e + 24 <= ix.size(); e += 24) {
          u64 at, n;
          memcpy(&at, ix.data() + e + 8, 8);  memcpy(&n, ix.data() + e + 16, 8);
          if (n == 0) continue;
          if (at >= db.fileBytes[f]) { rc = mb_fail(MB_EFORMAT, "'%s': block index of data file %u points past the end of the file", dbDir, f); goto done; }
          const uint8_t *bytes = host.data() + off[f - f0];
          const size_t t0 = tab.size();
          if (blocksOfEntry([&](u64 pos, void *dst, size_t len) { memcpy(dst, bytes + pos, len); return true; }, at, n, db.fileBytes[f], nk, tab) != MB_OK) {
            rc = mb_fail(MB_EFORMAT, "'%s': block index and block headers of data file %u disagree", dbDir, f);  goto done;
          }
          for (size_t t = t0; t < tab.size(); t++) tab[t].byteOff += off[f - f0];
        }
      }
      G.nKmers = nk;
      if (nk) {
        uint8_t    *dRaw = (uint8_t *)g.take(rawBytes + 16);
        UnionBlock *dTab = (UnionBlock *)g.take(tab.size() * sizeof(UnionBlock) + 16);
        void       *dKey = g.take(nk * W + 16);
        u32        *dVal = (u32 *)g.take(nk * 4 + 16);
        if (!dRaw || !dTab || !dKey || !dVal) { rc = mb_fail(MB_ENOMEM, "internal: decoded group does not fit"); goto done; }
        MB_UCUDA(cudaMemcpyAsync(dRaw, host.data(), rawBytes, cudaMemcpyHostToDevice, s));
        MB_UCUDA(cudaMemcpyAsync(dTab, tab.data(), tab.size() * sizeof(UnionBlock), cudaMemcpyHostToDevice, s));
        MB_UCUDA(cudaMemcpyAsync(dFileStart, G.fileStart, sizeof(G.fileStart), cudaMemcpyHostToDevice, s));
        const unsigned dgrid = (unsigned)((tab.size() + 127) / 128);
        if (kw == 1) decode_kernel<1><<<dgrid, 128, 0, s>>>(dRaw, rawBytes, dTab, tab.size(), db.wSuffix, (u64 *)dKey, dVal, nk, dErr);
        else         decode_kernel<2><<<dgrid, 128, 0, s>>>(dRaw, rawBytes, dTab, tab.size(), db.wSuffix, (ulonglong2 *)dKey, dVal, nk, dErr);
        MB_UCUDA(cudaGetLastError());
        G.dKey = dKey;  G.dVal = dVal;
        if (hold) hold->G = G;
        MB_UTRY(perGroup(db, G, s));
        u32 err = 0;
        MB_UCUDA(cudaMemcpyAsync(&err, dErr, 4, cudaMemcpyDeviceToHost, s));
        MB_UCUDA(cudaStreamSynchronize(s));                  //  (the host buffers are reused by the next group)
        if (err) { rc = mb_fail(MB_EFORMAT, "corrupt block in '%s' (decoder flags 0x%x)", dbDir, err); goto done; }
      }
      f0 = f1;
    }
    MB_UTRY(finish(s));
    MB_UCUDA(cudaStreamSynchronize(s));
    if (hold) {
      if (!hold->G.dFileStart) {                              //  a database without k-mers: 64 empty lists
        MB_UCUDA(cudaMemsetAsync(dFileStart, 0, 65 * 8, s));
        MB_UCUDA(cudaStreamSynchronize(s));
        hold->G.dFileStart = dFileStart;  hold->G.f0 = 0;  hold->G.f1 = 64;
      }
      hold->k = db.k;  hold->arena = arena;  hold->s = s;
      arena = nullptr;  s = nullptr;
    }
  }
done:
  if (s) cudaStreamDestroy(s);
  if (arena) cudaFree(arena);
  return rc;
}

//  merylExactLookup::value for many k-mers at once (kernel K9 above).  values[i] = the value of k-mer i (kmerLo[i], and
//  kmerHi[i] for k > 32; first base in the top bits, as everywhere) in the database, 0 if absent or outside
//  [minValue, maxValue].  The k-mers are looked up as given: the caller passes the canonical form for canonical databases,
//  like merylExactLookup's callers do (meryl-lookup: existence.C).
extern "C" int mb_lookup(const char *dbDir, uint32_t minValue, uint32_t maxValue, const uint64_t *kmerLo, const uint64_t *kmerHi,
                         uint64_t nQueries, uint32_t *values, int32_t device, uint64_t workspaceBytes) {
  mb_clear_error();
  if (!dbDir || (nQueries && (!kmerLo || !values))) return mb_fail(MB_EINVAL, "mb_lookup: NULL argument");
  if (nQueries == 0) return MB_OK;
  u64 *dQLo = nullptr, *dQHi = nullptr;
  u32 *dOut = nullptr;
  return decodeDbInGroups(dbDir, device, workspaceBytes, (size_t)nQueries * 20 + 4096,
    [&](const InputDb &db, Carve &top, cudaStream_t s) -> int {
      const bool wide = db.k > 32;
      if (wide && !kmerHi) return mb_fail(MB_EINVAL, "mb_lookup: k=%u needs the high words of the k-mers", db.k);
      dQLo = (u64 *)top.take((size_t)nQueries * 8);
      dQHi = wide ? (u64 *)top.take((size_t)nQueries * 8) : nullptr;
      dOut = (u32 *)top.take((size_t)nQueries * 4);
      if (!dQLo || (wide && !dQHi) || !dOut) return mb_fail(MB_ENOMEM, "the %llu queries do not fit the device arena", (unsigned long long)nQueries);
      MB_CUDA(cudaMemcpyAsync(dQLo, kmerLo, (size_t)nQueries * 8, cudaMemcpyHostToDevice, s));
      if (wide) MB_CUDA(cudaMemcpyAsync(dQHi, kmerHi, (size_t)nQueries * 8, cudaMemcpyHostToDevice, s));
      MB_CUDA(cudaMemsetAsync(dOut, 0, (size_t)nQueries * 4, s));
      return MB_OK;
    },
    [&](const InputDb &db, const DecodedGroup &G, cudaStream_t s) -> int {
      const unsigned qgrid = (unsigned)((nQueries + 255) / 256);
      if (db.k <= 32) lookup_kernel<1><<<qgrid, 256, 0, s>>>((const u64 *)G.dKey, G.dVal, G.dFileStart, G.f0, G.f1, db.k, dQLo, dQHi, nQueries, minValue, maxValue, dOut);
      else            lookup_kernel<2><<<qgrid, 256, 0, s>>>((const ulonglong2 *)G.dKey, G.dVal, G.dFileStart, G.f0, G.f1, db.k, dQLo, dQHi, nQueries, minValue, maxValue, dOut);
      MB_CUDA(cudaGetLastError());
      return MB_OK;
    },
    [&](cudaStream_t s) -> int {
      MB_CUDA(cudaMemcpyAsync(values, dOut, (size_t)nQueries * 4, cudaMemcpyDeviceToHost, s));
      return MB_OK;
    });
}

//  Every k-mer of a database with its value, in order (merylFileReader::nextMer, kmers-reader.C:456-531; what `meryl print`
//  walks): the blocks are decoded on the device, the lists come back to the host a group of files at a time and are handed
//  to the callback in pieces.  kmerHi is NULL for k <= 32.  A non-zero return of the callback stops the walk (MB_EIO).
extern "C" int mb_for_each_kmer(const char *dbDir, mb_kmer_callback cb, void *arg, int32_t device, uint64_t workspaceBytes) {
  mb_clear_error();
  if (!dbDir || !cb) return mb_fail(MB_EINVAL, "mb_for_each_kmer: NULL argument");
  std::vector<u64> lo, hi, both;
  std::vector<u32> val;
  return decodeDbInGroups(dbDir, device, workspaceBytes, 4096,
    [&](const InputDb &, Carve &, cudaStream_t) -> int { return MB_OK; },
    [&](const InputDb &db, const DecodedGroup &G, cudaStream_t s) -> int {
      const u64 n = G.nKmers;
      const bool wide = db.k > 32;
      val.resize(n);
      MB_CUDA(cudaMemcpyAsync(val.data(), G.dVal, n * 4, cudaMemcpyDeviceToHost, s));
      if (!wide) {
        lo.resize(n);
        MB_CUDA(cudaMemcpyAsync(lo.data(), G.dKey, n * 8, cudaMemcpyDeviceToHost, s));
        MB_CUDA(cudaStreamSynchronize(s));
      } else {
        both.resize(2 * n);  lo.resize(n);  hi.resize(n);
        MB_CUDA(cudaMemcpyAsync(both.data(), G.dKey, n * 16, cudaMemcpyDeviceToHost, s));
        MB_CUDA(cudaStreamSynchronize(s));
        for (u64 i = 0; i < n; i++) { lo[i] = both[2 * i]; hi[i] = both[2 * i + 1]; }
      }
      const u64 PIECE = (u64)1 << 20;
      for (u64 at = 0; at < n; at += PIECE) {
        const u64 m = std::min(PIECE, n - at);
        if (cb(arg, db.k, (const uint64_t *)lo.data() + at, wide ? (const uint64_t *)hi.data() + at : nullptr, val.data() + at, m) != 0)
          return mb_fail(MB_EIO, "mb_for_each_kmer: stopped by the callback");
      }
      return MB_OK;
    },
    [&](cudaStream_t) -> int { return MB_OK; });
}

//  meryl-lookup's per-sequence numbers (existence.C:59-86): nTotal[i] = the k-mers of sequence i (bases[seqStart[i] ..
//  seqStart[i+1])), nFound[i] = those whose forward or reverse-complement form is in the database with a value in
//  [minValue, maxValue]; *nDbKmers = the database's k-mers in that range (merylExactLookup::nKmers).  The first skipFirst
//  k-mer positions of every sequence do not count as found (-10x, include-exclude.C:70).  Kernel K10 above, group of
//  files by group of files.
static int existenceRun(const char *dbDir, uint32_t minValue, uint32_t maxValue, const char *bases, const uint64_t *seqStart,
                        uint64_t nSeqs, uint32_t skipFirst, uint64_t *nTotal, uint64_t *nFound, uint64_t *nDbKmers,
                        uint8_t *flagsOut, uint32_t *valuesOut, int32_t device, uint64_t workspaceBytes) {
  mb_clear_error();
  if (!dbDir || (nSeqs && !seqStart)) return mb_fail(MB_EINVAL, "mb_existence: NULL argument");
  const u64 nBases = nSeqs ? seqStart[nSeqs] : 0;
  if (nBases && !bases) return mb_fail(MB_EINVAL, "mb_existence: NULL argument");
  for (u64 i = 0; i < nSeqs; i++) if (seqStart[i] > seqStart[i + 1]) return mb_fail(MB_EINVAL, "mb_existence: seqStart is not ascending");
  if (nSeqs && seqStart[0] != 0) return mb_fail(MB_EINVAL, "mb_existence: seqStart[0] must be 0");
  unsigned char *dBases = nullptr, *dFlag = nullptr;
  u64 *dSeqStart = nullptr, *dTot = nullptr, *dFnd = nullptr, *dDbN = nullptr;
  u32 *dValues = nullptr;
  u32 kDb = 0;
  bool ranGroup = false;
  return decodeDbInGroups(dbDir, device, workspaceBytes, (size_t)nBases * (valuesOut ? 6 : 2) + (size_t)nSeqs * 24 + 8192,
    [&](const InputDb &db, Carve &top, cudaStream_t s) -> int {
      kDb = db.k;
      dDbN = (u64 *)top.take(8);
      if (!dDbN) return mb_fail(MB_ENOMEM, "device arena too small");
      MB_CUDA(cudaMemsetAsync(dDbN, 0, 8, s));
      if (!nSeqs) return MB_OK;
      dBases    = (unsigned char *)top.take((size_t)nBases + 16);
      dFlag     = (unsigned char *)top.take((size_t)nBases + 16);
      dSeqStart = (u64 *)top.take((size_t)(nSeqs + 1) * 8);
      dTot      = (u64 *)top.take((size_t)nSeqs * 8);
      dFnd      = (u64 *)top.take((size_t)nSeqs * 8);
      if (valuesOut) dValues = (u32 *)top.take((size_t)nBases * 4 + 16);
      if (!dBases || !dFlag || !dSeqStart || !dTot || !dFnd || (valuesOut && !dValues))
        return mb_fail(MB_ENOMEM, "the %llu bases of %llu sequences do not fit the device arena", (unsigned long long)nBases, (unsigned long long)nSeqs);
      if (nBases) MB_CUDA(cudaMemcpyAsync(dBases, bases, (size_t)nBases, cudaMemcpyHostToDevice, s));
      MB_CUDA(cudaMemcpyAsync(dSeqStart, seqStart, (size_t)(nSeqs + 1) * 8, cudaMemcpyHostToDevice, s));
      MB_CUDA(cudaMemsetAsync(dFlag, 0, (size_t)nBases + 16, s));
      if (dValues) MB_CUDA(cudaMemsetAsync(dValues, 0, (size_t)nBases * 4 + 16, s));
      MB_CUDA(cudaMemsetAsync(dTot, 0, (size_t)nSeqs * 8, s));
      MB_CUDA(cudaMemsetAsync(dFnd, 0, (size_t)nSeqs * 8, s));
      return MB_OK;
    },
    [&](const InputDb &db, const DecodedGroup &G, cudaStream_t s) -> int {
      count_in_range_kernel<<<(unsigned)std::min<u64>((G.nKmers + 255) / 256, 148 * 8), 256, 0, s>>>(G.dVal, G.nKmers, minValue, maxValue, dDbN);
      MB_CUDA(cudaGetLastError());
      if (!nBases) return MB_OK;
      ranGroup = true;
      const unsigned grid = (unsigned)((nBases + 255) / 256);
      if (db.k <= 32) existence_kernel<1><<<grid, 256, 0, s>>>((const u64 *)G.dKey, G.dVal, G.dFileStart, G.f0, G.f1, db.k, dBases, nBases, dSeqStart, nSeqs, skipFirst, minValue, maxValue, dFlag, dValues);
      else            existence_kernel<2><<<grid, 256, 0, s>>>((const ulonglong2 *)G.dKey, G.dVal, G.dFileStart, G.f0, G.f1, db.k, dBases, nBases, dSeqStart, nSeqs, skipFirst, minValue, maxValue, dFlag, dValues);
      MB_CUDA(cudaGetLastError());
      return MB_OK;
    },
    [&](cudaStream_t s) -> int {
      if (nDbKmers) MB_CUDA(cudaMemcpyAsync(nDbKmers, dDbN, 8, cudaMemcpyDeviceToHost, s));
      if (!nSeqs) return MB_OK;
      if (nBases) {
        if (!ranGroup) {                                    //  a database without k-mers: the sequences' own k-mers still count
          const unsigned grid = (unsigned)((nBases + 255) / 256);
          if (kDb <= 32) existence_kernel<1><<<grid, 256, 0, s>>>(nullptr, nullptr, nullptr, 0, 0, kDb, dBases, nBases, dSeqStart, nSeqs, skipFirst, minValue, maxValue, dFlag, dValues);
          else           existence_kernel<2><<<grid, 256, 0, s>>>(nullptr, nullptr, nullptr, 0, 0, kDb, dBases, nBases, dSeqStart, nSeqs, skipFirst, minValue, maxValue, dFlag, dValues);
          MB_CUDA(cudaGetLastError());
        }
        existence_sum_kernel<<<(unsigned)((nBases + 255) / 256), 256, 0, s>>>(dFlag, nBases, dSeqStart, nSeqs, dTot, dFnd);
        MB_CUDA(cudaGetLastError());
      }
      if (nTotal) MB_CUDA(cudaMemcpyAsync(nTotal, dTot, (size_t)nSeqs * 8, cudaMemcpyDeviceToHost, s));
      if (nFound) MB_CUDA(cudaMemcpyAsync(nFound, dFnd, (size_t)nSeqs * 8, cudaMemcpyDeviceToHost, s));
      if (flagsOut && nBases)  MB_CUDA(cudaMemcpyAsync(flagsOut, dFlag, (size_t)nBases, cudaMemcpyDeviceToHost, s));
      if (valuesOut && nBases) MB_CUDA(cudaMemcpyAsync(valuesOut, dValues, (size_t)nBases * 4, cudaMemcpyDeviceToHost, s));
      return MB_OK;
    });
}

extern "C" int mb_existence(const char *dbDir, uint32_t minValue, uint32_t maxValue, const char *bases, const uint64_t *seqStart,
                            uint64_t nSeqs, uint32_t skipFirst, uint64_t *nTotal, uint64_t *nFound, uint64_t *nDbKmers,
                            int32_t device, uint64_t workspaceBytes) {
  if (nSeqs && (!nTotal || !nFound)) { mb_clear_error(); return mb_fail(MB_EINVAL, "mb_existence: NULL argument"); }
  return existenceRun(dbDir, minValue, maxValue, bases, seqStart, nSeqs, skipFirst, nTotal, nFound, nDbKmers, nullptr, nullptr, device, workspaceBytes);
}

//  The per-position form behind meryl-lookup's -bed / -wig-count / -wig-depth (src/meryl-lookup/dump.C:86-157): for position p of
//  the letters, flags[p] bit 0 = a k-mer starts here, bit 1 = the database holds its forward or reverse-complement form
//  (exists(c) || exists(n)); values[p] (may be NULL) = value(fmer) + value(rmer), once when the two are equal.
extern "C" int mb_position_lookup(const char *dbDir, uint32_t minValue, uint32_t maxValue, const char *bases, const uint64_t *seqStart,
                                  uint64_t nSeqs, uint8_t *flags, uint32_t *values, int32_t device, uint64_t workspaceBytes) {
  if (nSeqs && seqStart && seqStart[nSeqs] && !flags) { mb_clear_error(); return mb_fail(MB_EINVAL, "mb_position_lookup: NULL argument"); }
  return existenceRun(dbDir, minValue, maxValue, bases, seqStart, nSeqs, 0, nullptr, nullptr, nullptr, flags, values, device, workspaceBytes);
}

//  ---- the resident table (merylExactLookup as an object: load once, ask many times; kmers-lookup.H:83-146) ----------------
//  mb_table_open decodes the whole database into device memory and keeps it there (180 GB of HBM hold the decoded lists of a
//  human-size read database); the queries are K9 / K10 launches against the resident lists, with a scratch allocation per
//  call for the queries themselves.  A database that does not fit is MB_ENOMEM: the one-shot calls above (group by group)
//  are the fallback.
struct mb_table {
  DecodeHold h;
  int32_t    device = -1;
  u32        minValue = 0, maxValue = 0xffffffffu;
  u64        nInRange = 0;
};

extern "C" int mb_table_open(const char *dbDir, uint32_t minValue, uint32_t maxValue, int32_t device, uint64_t workspaceBytes, mb_table **out) {
  mb_clear_error();
  if (!dbDir || !out) return mb_fail(MB_EINVAL, "mb_table_open: NULL argument");
  mb_table *T = new mb_table;
  T->minValue = minValue;  T->maxValue = maxValue;
  u64 *dN = nullptr;
  const int rc = decodeDbInGroups(dbDir, device, workspaceBytes, 4096,
    [&](const InputDb &, Carve &top, cudaStream_t s) -> int {
      dN = (u64 *)top.take(8);
      if (!dN) return mb_fail(MB_ENOMEM, "device arena too small");
      MB_CUDA(cudaMemsetAsync(dN, 0, 8, s));
      return MB_OK;
    },
    [&](const InputDb &, const DecodedGroup &G, cudaStream_t s) -> int {
      count_in_range_kernel<<<(unsigned)std::min<u64>((G.nKmers + 255) / 256, 148 * 8), 256, 0, s>>>(G.dVal, G.nKmers, minValue, maxValue, dN);
      MB_CUDA(cudaGetLastError());
      return MB_OK;
    },
    [&](cudaStream_t s) -> int { MB_CUDA(cudaMemcpyAsync(&T->nInRange, dN, 8, cudaMemcpyDeviceToHost, s)); return MB_OK; },
    &T->h);
  if (rc != MB_OK) { delete T; return rc; }
  if (cudaGetDevice(&T->device) != cudaSuccess) T->device = device;
  *out = T;
  return MB_OK;
}

extern "C" void mb_table_close(mb_table *T) {
  if (!T) return;
  if (T->device >= 0) cudaSetDevice(T->device);
  if (T->h.s) { cudaStreamSynchronize(T->h.s); cudaStreamDestroy(T->h.s); }
  if (T->h.arena) cudaFree(T->h.arena);
  delete T;
}

extern "C" uint32_t mb_table_k(const mb_table *T)      { return T ? T->h.k : 0; }
extern "C" uint64_t mb_table_kmers(const mb_table *T)  { return T ? T->nInRange : 0; }

//  mb_existence / mb_position_lookup against the resident table (any of the outputs may be NULL).
extern "C" int mb_table_query(mb_table *T, const char *bases, const uint64_t *seqStart, uint64_t nSeqs, uint32_t skipFirst,
                              uint64_t *nTotal, uint64_t *nFound, uint8_t *flags, uint32_t *values) {
  mb_clear_error();
  if (!T || (nSeqs && !seqStart)) return mb_fail(MB_EINVAL, "mb_table_query: NULL argument");
  if (nSeqs == 0) return MB_OK;
  const u64 nBases = seqStart[nSeqs];
  if (seqStart[0] != 0 || (nBases && !bases)) return mb_fail(MB_EINVAL, "mb_table_query: bad sequence offsets");
  for (u64 i = 0; i < nSeqs; i++) if (seqStart[i] > seqStart[i + 1]) return mb_fail(MB_EINVAL, "mb_table_query: seqStart is not ascending");
  MB_CUDA(cudaSetDevice(T->device));
  const size_t bytes = ((size_t)nBases + 256) * 2 + (size_t)(nSeqs + 1) * 8 + 256 + (size_t)nSeqs * 16 + 512 + (values ? (size_t)nBases * 4 + 256 : 0) + 4096;
  uint8_t *scratch = nullptr;
  if (cudaMalloc((void **)&scratch, bytes) != cudaSuccess) { cudaGetLastError(); return mb_fail(MB_ENOMEM, "cannot allocate %zu bytes next to the resident table: ask in smaller batches", bytes); }
  struct Free { uint8_t *p; ~Free() { cudaFree(p); } } freeIt{ scratch };
  Carve top(scratch, bytes);
  unsigned char *dBases = (unsigned char *)top.take((size_t)nBases + 16), *dFlag = (unsigned char *)top.take((size_t)nBases + 16);
  u64 *dSeqStart = (u64 *)top.take((size_t)(nSeqs + 1) * 8), *dTot = (u64 *)top.take((size_t)nSeqs * 8), *dFnd = (u64 *)top.take((size_t)nSeqs * 8);
  u32 *dValues = values ? (u32 *)top.take((size_t)nBases * 4 + 16) : nullptr;
  if (!dBases || !dFlag || !dSeqStart || !dTot || !dFnd || (values && !dValues)) return mb_fail(MB_ENOMEM, "internal: query scratch too small");
  cudaStream_t s = T->h.s;
  const DecodedGroup &G = T->h.G;
  if (nBases) MB_CUDA(cudaMemcpyAsync(dBases, bases, (size_t)nBases, cudaMemcpyHostToDevice, s));
  MB_CUDA(cudaMemcpyAsync(dSeqStart, seqStart, (size_t)(nSeqs + 1) * 8, cudaMemcpyHostToDevice, s));
  MB_CUDA(cudaMemsetAsync(dFlag, 0, (size_t)nBases + 16, s));
  MB_CUDA(cudaMemsetAsync(dTot, 0, (size_t)nSeqs * 8, s));
  MB_CUDA(cudaMemsetAsync(dFnd, 0, (size_t)nSeqs * 8, s));
  if (dValues) MB_CUDA(cudaMemsetAsync(dValues, 0, (size_t)nBases * 4 + 16, s));
  if (nBases) {
    const unsigned grid = (unsigned)((nBases + 255) / 256);
    if (T->h.k <= 32) existence_kernel<1><<<grid, 256, 0, s>>>((const u64 *)G.dKey, G.dVal, G.dFileStart, 0, 64, T->h.k, dBases, nBases, dSeqStart, nSeqs, skipFirst, T->minValue, T->maxValue, dFlag, dValues);
    else              existence_kernel<2><<<grid, 256, 0, s>>>((const ulonglong2 *)G.dKey, G.dVal, G.dFileStart, 0, 64, T->h.k, dBases, nBases, dSeqStart, nSeqs, skipFirst, T->minValue, T->maxValue, dFlag, dValues);
    MB_CUDA(cudaGetLastError());
    existence_sum_kernel<<<grid, 256, 0, s>>>(dFlag, nBases, dSeqStart, nSeqs, dTot, dFnd);
    MB_CUDA(cudaGetLastError());
  }
  if (nTotal) MB_CUDA(cudaMemcpyAsync(nTotal, dTot, (size_t)nSeqs * 8, cudaMemcpyDeviceToHost, s));
  if (nFound) MB_CUDA(cudaMemcpyAsync(nFound, dFnd, (size_t)nSeqs * 8, cudaMemcpyDeviceToHost, s));
  if (flags && nBases)  MB_CUDA(cudaMemcpyAsync(flags, dFlag, (size_t)nBases, cudaMemcpyDeviceToHost, s));
  if (values && nBases) MB_CUDA(cudaMemcpyAsync(values, dValues, (size_t)nBases * 4, cudaMemcpyDeviceToHost, s));
  MB_CUDA(cudaStreamSynchronize(s));
  return MB_OK;
}

//  mb_lookup against the resident table.
extern "C" int mb_table_values(mb_table *T, const uint64_t *kmerLo, const uint64_t *kmerHi, uint64_t nQueries, uint32_t *values) {
  mb_clear_error();
  if (!T || (nQueries && (!kmerLo || !values))) return mb_fail(MB_EINVAL, "mb_table_values: NULL argument");
  if (nQueries == 0) return MB_OK;
  const bool wide = T->h.k > 32;
  if (wide && !kmerHi) return mb_fail(MB_EINVAL, "mb_table_values: k=%u needs the high words of the k-mers", T->h.k);
  MB_CUDA(cudaSetDevice(T->device));
  const size_t bytes = (size_t)nQueries * 20 + 4096;
  uint8_t *scratch = nullptr;
  if (cudaMalloc((void **)&scratch, bytes) != cudaSuccess) { cudaGetLastError(); return mb_fail(MB_ENOMEM, "cannot allocate %zu bytes next to the resident table: ask in smaller batches", bytes); }
  struct Free { uint8_t *p; ~Free() { cudaFree(p); } } freeIt{ scratch };
  Carve top(scratch, bytes);
  u64 *dQLo = (u64 *)top.take((size_t)nQueries * 8), *dQHi = wide ? (u64 *)top.take((size_t)nQueries * 8) : nullptr;
  u32 *dOut = (u32 *)top.take((size_t)nQueries * 4);
  if (!dQLo || (wide && !dQHi) || !dOut) return mb_fail(MB_ENOMEM, "internal: query scratch too small");
  cudaStream_t s = T->h.s;
  const DecodedGroup &G = T->h.G;
  MB_CUDA(cudaMemcpyAsync(dQLo, kmerLo, (size_t)nQueries * 8, cudaMemcpyHostToDevice, s));
  if (wide) MB_CUDA(cudaMemcpyAsync(dQHi, kmerHi, (size_t)nQueries * 8, cudaMemcpyHostToDevice, s));
  MB_CUDA(cudaMemsetAsync(dOut, 0, (size_t)nQueries * 4, s));
  const unsigned qgrid = (unsigned)((nQueries + 255) / 256);
  if (!wide) lookup_kernel<1><<<qgrid, 256, 0, s>>>((const u64 *)G.dKey, G.dVal, G.dFileStart, 0, 64, T->h.k, dQLo, dQHi, nQueries, T->minValue, T->maxValue, dOut);
  else       lookup_kernel<2><<<qgrid, 256, 0, s>>>((const ulonglong2 *)G.dKey, G.dVal, G.dFileStart, 0, 64, T->h.k, dQLo, dQHi, nQueries, T->minValue, T->maxValue, dOut);
  MB_CUDA(cudaGetLastError());
  MB_CUDA(cudaMemcpyAsync(values, dOut, (size_t)nQueries * 4, cudaMemcpyDeviceToHost, s));
  MB_CUDA(cudaStreamSynchronize(s));
  return MB_OK;
}

//  `count-suffix=ACG` (merylCommandBuilder.C:271-274; merylOp-countSimple.C:50-93, :172-199): only the k-mers whose counted
//  form ends with the given bases are kept, and the database is written with the simple counter's geometry -- prefix width
//  6 + max(0, 2k - 2s - 26), a block for every prefix.  The reference filters while it counts into a direct-indexed array of
//  4^(k-s) entries (the option exists to split a count that does not fit memory into 4^s jobs); here the count is not
//  memory-bound, so the full count is made first and this one-input operation keeps the wanted k-mers (same database).
extern "C" int mb_keep_suffix(const char *inDir, const char *suffix, const char *outDir, int32_t device, uint64_t workspaceBytes,
                              mb_union_stats *stOut) {
  mb_clear_error();
  if (!inDir || !suffix || !outDir) return mb_fail(MB_EINVAL, "mb_keep_suffix: NULL argument");
  InputDb db;
  MB_TRY(openDb(inDir, db));
  const u32 s = (u32)strlen(suffix);
  if (s == 0 || s > 32 || s >= db.k) return mb_fail(MB_EINVAL, "count-suffix '%s': 1..32 bases, shorter than k=%u", suffix, db.k);
  if (2 * db.k - 2 * s > 42) return mb_fail(MB_EINVAL, "count-suffix '%s' with k=%u: not possible (the reference's simple mode needs 2k - 2s <= 42)", suffix, db.k);
  if (2 * db.k - 2 * s < 6)  return mb_fail(MB_EINVAL, "count-suffix '%s' with k=%u leaves fewer than 3 bases", suffix, db.k);
  u64 value = 0;
  for (u32 i = 0; i < s; i++) {
    const char c = suffix[i];
    if (!strchr("ACGTacgt", c) || c == 0) return mb_fail(MB_EINVAL, "count-suffix '%s' is not made of ACGT", suffix);
    value = (value << 2) | (u64)((c >> 1) & 3);            //  kmers-tiny.H:86: A=0 C=1 T=2 G=3
  }
  const u32 psbits = 2 * db.k - 2 * s - 6;
  MbMergeEndpoints E;
  E.param2     = s == 32 ? ~0ull : (((u64)1 << (2 * s)) - 1);
  E.outPrefix  = 6 + (psbits > 20 ? psbits - 20 : 0);       //  merylOp-countSimple.C:172-175
  E.streamMode = 0;
  return mb_merge_pipeline(MB_OP_KEEP_SUFFIX, value, &inDir, 1, outDir, device, workspaceBytes, stOut, &E);
}

extern "C" int mb_union_sum(const char *const *inDirs, uint32_t nInputs, const char *outDir, int32_t device,
                            uint64_t workspaceBytes, mb_union_stats *stOut) {
  return mb_set_operation(MB_OP_UNION_SUM, inDirs, nInputs, outDir, device, workspaceBytes, stOut);
}

static int setOperationImpl(uint32_t op, uint64_t param, const char *const *inDirs, uint32_t nInputs, const char *outDir, int32_t device,
                            uint64_t workspaceBytes, mb_union_stats *stOut) {
  return mb_merge_pipeline(op, param, inDirs, nInputs, outDir, device, workspaceBytes, stOut, nullptr);
}

static void removeDb(const std::string &dir) {              //  a database directory we made ourselves: 129 files
  char path[4096];
  for (u32 f = 0; f < 64; f++)
    for (int isIdx = 0; isIdx < 2; isIdx++) { mb_file_name(path, sizeof(path), dir.c_str(), f, isIdx != 0); unlink(path); }
  unlink((dir + "/merylIndex").c_str());
  rmdir(dir.c_str());
}

extern "C" int mb_set_operation(uint32_t op, const char *const *inDirs, uint32_t nInputs, const char *outDir, int32_t device,
                                uint64_t workspaceBytes, mb_union_stats *stOut) {
  mb_clear_error();
  if (op > MB_OP_SYMMETRIC_DIFFERENCE) return mb_fail(MB_EINVAL, "mb_set_operation: unknown operation %u", op);
  if (op == MB_OP_SYMMETRIC_DIFFERENCE && nInputs > 2 && inDirs && outDir) {
    //  "the k-mers of exactly one input, with that input's value" (merylOp-nextMer.C:609-612) is not a left fold, but it is
    //  a composition of folds:  intersect( union-sum(inputs), equal-to 1 ( union(inputs) ) )  -- union counts the inputs
    //  that hold a k-mer, union-sum of a k-mer held by one input is that input's value, intersect keeps its first input's.
    const std::string base = std::string(outDir) + ".tmp-symdiff-";
    const std::string dCount = base + "count", dSum = base + "sum", dOne = base + "one";
    mb_union_stats st1, st2, st3;
    int rc = setOperationImpl(MB_OP_UNION, 0, inDirs, nInputs, dCount.c_str(), device, workspaceBytes, &st1);
    if (rc == MB_OK) rc = setOperationImpl(MB_OP_UNION_SUM, 0, inDirs, nInputs, dSum.c_str(), device, workspaceBytes, &st2);
    const char *one = dCount.c_str();
    if (rc == MB_OK) rc = setOperationImpl(MB_OP_EQUAL_TO, 1, &one, 1, dOne.c_str(), device, workspaceBytes, &st3);
    const char *two[2] = { dSum.c_str(), dOne.c_str() };
    mb_union_stats st4;
    if (rc == MB_OK) rc = setOperationImpl(MB_OP_INTERSECT, 0, two, 2, outDir, device, workspaceBytes, &st4);
    char keep[1024];
    strncpy(keep, mb_last_error(), sizeof(keep) - 1);  keep[sizeof(keep) - 1] = 0;
    removeDb(dCount);  removeDb(dSum);  removeDb(dOne);
    if (rc != MB_OK) return mb_fail(rc, "%s", keep);
    if (stOut) {
      *stOut = st4;
      stOut->nInputs = nInputs;  stOut->nInputKmers = st2.nInputKmers;  stOut->inBytes = st2.inBytes;
      stOut->nLaunches += st1.nLaunches + st2.nLaunches + st3.nLaunches;
      stOut->msTotal += st1.msTotal + st2.msTotal + st3.msTotal;
      stOut->msDecode += st1.msDecode + st2.msDecode + st3.msDecode;
      stOut->msMerge += st1.msMerge + st2.msMerge + st3.msMerge;
      stOut->msEncode += st1.msEncode + st2.msEncode + st3.msEncode;
    }
    return MB_OK;
  }
  return setOperationImpl(op, 0, inDirs, nInputs, outDir, device, workspaceBytes, stOut);
}

extern "C" int mb_value_operation(uint32_t op, uint64_t param, const char *inDir, const char *outDir, int32_t device,
                                  uint64_t workspaceBytes, mb_union_stats *stOut) {
  mb_clear_error();
  if (op < MB_OP_LESS_THAN || op > MB_OP_MODULO) return mb_fail(MB_EINVAL, "mb_value_operation: unknown operation %u", op);
  return setOperationImpl(op, param, &inDir, 1, outDir, device, workspaceBytes, stOut);
}

int mb_merge_pipeline(uint32_t op, uint64_t param, const char *const *inDirs, uint32_t nInputs, const char *outDir, int32_t device,
                      uint64_t workspaceBytes, mb_union_stats *stOut, const MbMergeEndpoints *ends) {
  mb_clear_error();
  MbMergeEndpoints dflt;
  const MbMergeEndpoints &E = ends ? *ends : dflt;
  if (E.memInputs) nInputs = E.nMem;
  if ((!inDirs && !E.memInputs) || nInputs == 0 || (!outDir && !E.outImages)) return mb_fail(MB_EINVAL, "mb_set_operation: NULL argument");
  if (op == MB_OP_SYMMETRIC_DIFFERENCE && nInputs > 2)
    return mb_fail(MB_EINVAL, "internal: symmetric-difference of more than two databases is a composition (mb_set_operation)");
  int nDev = 0;
  if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0)
    return mb_fail(MB_ENODEVICE, "mb_set_operation: no CUDA device (this library has no CPU path)");

  std::vector<InputDb> in(nInputs);
  for (u32 i = 0; i < nInputs; i++) {
    if (E.memInputs) { MB_TRY(openMem(E.memInputs[i], E.k, E.wPrefix, in[i])); continue; }
    if (!inDirs[i]) return mb_fail(MB_EINVAL, "mb_union_sum: NULL input");
    MB_TRY(openDb(inDirs[i], in[i]));
    if (in[i].k != in[0].k) return mb_fail(MB_EINVAL, "k-mer sizes differ: '%s' has k=%u, '%s' has k=%u", inDirs[0], in[0].k, inDirs[i], in[i].k);
  }
  const u32 k = in[0].k, kw = k <= 32 ? 1 : 2, W = 8 * kw;
  const u32 wP = E.outPrefix, wS = 2 * k - wP;
  //  (no suffix bits at all is fine: k = 6 under the stream writer's 12-bit prefix -- the reference binary writes such
  //   databases, and fails an assertion for k = 5 --, k = 3 in a merge of count batches, which keeps the count's geometry)
  if (2 * k < wP) return mb_fail(MB_EINVAL, "set operations need k >= 6 (the stream writer's 12-bit prefix)");
  const u64 blocksPerFile = (u64)1 << (wP - 6);

  mb_union_stats st;
  memset(&st, 0, sizeof(st));
  st.nInputs = nInputs;  st.k = k;

  int rc = MB_OK;
  uint8_t *arena = nullptr;
  cudaStream_t s = nullptr;
  cudaEvent_t ev[2] = { nullptr, nullptr }, evAll[2] = { nullptr, nullptr };
  MbHistogram hist;
  std::map<u64, u64> histMap;
  u64 nUnique = 0, nDistinct = 0, nTotal = 0;

  {
    if (device >= 0) MB_UCUDA(cudaSetDevice(device));
    size_t freeB = 0, totalB = 0;
    mb_arena_cache_release(-1);                      //  (a count context's cached arena would otherwise hide most of the device)
    MB_UCUDA(cudaMemGetInfo(&freeB, &totalB));
    if (E.arena) freeB = E.arenaBytes;
    //  what a group of files needs on the device (upper bound):
    //    raw bytes + block tables | decoded inputs | two merge buffers | encoder (meta, output, sub-block table, index) | big-value list
    auto groupNeed = [&](u32 f0, u32 f1) -> size_t {
      size_t need = (size_t)64 << 20;
      u64 nk = 0;
      for (u32 i = 0; i < nInputs; i++)
        for (u32 f = f0; f < f1; f++) {
          need += in[i].fileBytes[f] + 256;
          need += in[i].fileBytes[f] / 104 * sizeof(UnionBlock) + 4096;      //  a block is at least 104 bytes
          need += in[i].fileKmers[f] * (W + 4) + 512;
          nk   += in[i].fileKmers[f];
        }
      need += (nInputs > 2 ? 2 : 1) * (nk * (W + 4) + 1024);                //  merge output (ping-pong only when a fold has more than one step)
      need += nk * ((wS + 34 + 7) / 8 + 1) + nk / 64;                    //  encoded blocks (every block <= n(b+34) + headers) ...
      need += (size_t)(f1 - f0) * blocksPerFile * (E.streamMode ? 4096 : 512 + sizeof(MbBlockMeta) + 24 + 16);   //  ... headers, tables
      need += nk / MG_TILE * 16 + nk / MB_HIST_SMALL * 4 + 65536;
      if (E.streamMode) need += (nk / streamBlockMax() + 2) * (2 * sizeof(MbSplitPart) + 256);   //  parts of over-full prefixes: table + containers
      return need;
    };

    size_t want = workspaceBytes ? (size_t)workspaceBytes : (freeB > ((size_t)3 << 30) ? freeB - ((size_t)2 << 30) : freeB / 2);
    if (!workspaceBytes) {                                   //  no more than all 64 files at once need (allocating ~175 GB takes time)
      const size_t all = groupNeed(0, 64);
      const size_t cap = all + all / 8 + ((size_t)288 << 20);
      if (cap < want) want = cap;
    }
    if (want > freeB) want = freeB;
    want = want / 256 * 256;
    if (E.arena) { arena = E.arena;  want = E.arenaBytes / 256 * 256; }
    else if (cudaMalloc((void **)&arena, want) != cudaSuccess) { cudaGetLastError(); arena = nullptr; rc = mb_fail(MB_ENOMEM, "cannot allocate a %zu byte device arena", want); goto done; }
    MB_UCUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) { MB_UCUDA(cudaEventCreate(&ev[i])); MB_UCUDA(cudaEventCreate(&evAll[i])); }
    MB_UCUDA(cudaEventRecord(evAll[0], s));

    if (outDir && mkdir(outDir, 0777) != 0 && errno != EEXIST) { rc = mb_fail(MB_EIO, "cannot create directory '%s': %s", outDir, strerror(errno)); goto done; }

    //  small fixed tables at the bottom of the arena
    Carve top(arena, want);
    u64 *dHistSmall = (u64 *)top.take((size_t)MB_HIST_SMALL * 8);
    u64 *dCtl       = (u64 *)top.take(64 * 8);              //  [0..7] encoder totals, [8] nOut, [9] hist overflow count, [10..12] hist totals, [16] err
    u64 *dFileBase  = (u64 *)top.take(64 * 8);
    if (!dHistSmall || !dCtl || !dFileBase) { rc = mb_fail(MB_ENOMEM, "device arena too small"); goto done; }
    MB_UCUDA(cudaMemsetAsync(dHistSmall, 0, (size_t)MB_HIST_SMALL * 8, s));
    MB_UCUDA(cudaMemsetAsync(dCtl, 0, 64 * 8, s));
    MB_UCUDA(cudaMemsetAsync(dFileBase, 0, 64 * 8, s));
    const size_t fixedOff = (top.off + 255) / 256 * 256;   //  the group carve starts here: keep its base aligned

    PinnedRing ring;
    MB_UTRY(ring.init());
    HostBuf outIdx;
    u32 f0 = 0;
    while (f0 < 64) {
      if (E.onlyFiles && !E.onlyFiles[f0]) { f0++; continue; }     //  (several GPUs: another rank's file)
      u32 f1 = f0 + 1;
      if (fixedOff + groupNeed(f0, f1) > want) {
        rc = mb_fail(MB_ENOMEM, "data file %u of the inputs needs %zu bytes on the device, %zu available", f0, groupNeed(f0, f1), want - fixedOff);
        goto done;
      }
      const u32 maxFiles = getenv("MB_UNION_MAX_FILES") ? (u32)strtoul(getenv("MB_UNION_MAX_FILES"), nullptr, 10) : 64;   //  testing knob
      while (f1 < 64 && f1 - f0 < maxFiles && !(E.onlyFiles && !E.onlyFiles[f1]) && fixedOff + groupNeed(f0, f1 + 1) <= want) f1++;
      st.nGroups++;

      Carve g(arena + fixedOff, want - fixedOff);
      typedef std::pair<void *, u32 *> List;
      std::vector<List> lists(nInputs);
      std::vector<u64>  listN(nInputs, 0);
      u32 *dErr = (u32 *)(dCtl + 16);

      //  ---- read + decode every input's part of the group -------------------------------------------------
      //  The block tables come from the inputs' block indexes ({prefix, byte offset, nKmers} per block); the data files'
      //  bytes stream through the page-locked ring to the device (host threads pread one chunk while the previous one is
      //  on the bus); one decode launch per input.  The decoder checks every container against the buffer's end and
      //  every block's k-mer count against the index.
      MB_UCUDA(cudaEventRecord(ev[0], s));
      std::vector<FileJob> jobs;
      struct FdCloser { std::vector<FileJob> &v; ~FdCloser() { for (FileJob &j : v) if (j.fd >= 0) { close(j.fd); j.fd = -1; } } } fdCloser{ jobs };
      std::vector<size_t>  inBytes(nInputs, 0);
      for (u32 i = 0; i < nInputs; i++)
        for (u32 f = f0; f < f1; f++) {
          FileJob j;  j.input = i;  j.file = f;  j.off = inBytes[i];  j.bytes = (size_t)in[i].fileBytes[f];
          inBytes[i] += (j.bytes + 7) / 8 * 8;
          if (!in[i].mem && j.bytes) {
            char path[4096];
            mb_file_name(path, sizeof(path), in[i].dir.c_str(), f, false);
            j.fd = open(path, O_RDONLY);
            if (j.fd < 0) { rc = mb_fail(MB_EIO, "cannot open '%s': %s", path, strerror(errno)); jobs.push_back(std::move(j)); goto done; }
          }
          const std::vector<uint8_t> &ix = in[i].indexOf(f);
          const uint8_t *memBytes = in[i].mem ? in[i].mem->dataPtr[f] : nullptr;
          const int fd = j.fd;
          for (size_t e = 0; e + 24 <= ix.size() && rc == MB_OK; e += 24) {
            u64 at, nkB;
            memcpy(&at, ix.data() + e + 8, 8);  memcpy(&nkB, ix.data() + e + 16, 8);
            if (nkB == 0) continue;                         //  an empty block (block writer) or an absent one (stream writer)
            if (at >= j.bytes) rc = mb_fail(MB_EFORMAT, "'%s': block index of data file %u points past the end of the file", in[i].dir.c_str(), f);
            else if (blocksOfEntry([&](u64 pos, void *dst, size_t len) {
                       if (memBytes) { memcpy(dst, memBytes + pos, len); return true; }
                       return pread(fd, dst, len, (off_t)pos) == (ssize_t)len;
                     }, at, nkB, j.bytes, j.nk, j.tab) != MB_OK)
              rc = mb_fail(MB_EFORMAT, "'%s': block index and block headers of data file %u disagree", in[i].dir.c_str(), f);
          }
          jobs.push_back(std::move(j));
          if (rc != MB_OK) goto done;
        }

      std::vector<std::vector<UnionBlock>> tabs(nInputs);    //  alive until the stream has consumed them
      for (u32 i = 0; i < nInputs; i++) {
        u64 nk = 0;
        for (u32 f = f0; f < f1; f++) nk += in[i].fileKmers[f];
        uint8_t *dRaw = (uint8_t *)g.take(inBytes[i] + 16);
        void    *dKey = g.take(nk * W + 16);
        u32     *dVal = (u32 *)g.take(nk * 4 + 16);
        if (!dRaw || !dKey || !dVal) { rc = mb_fail(MB_ENOMEM, "internal: union group does not fit"); goto done; }
        std::vector<UnionBlock> &tab = tabs[i];
        u64 idx = 0;
        for (const FileJob &j : jobs) {
          if (j.input != i) continue;
          for (UnionBlock ub : j.tab) { ub.byteOff += j.off;  ub.outIdx += idx;  tab.push_back(ub); }
          idx += j.nk;
          st.h2dBytes += j.bytes;  st.inBytes += j.bytes;
        }
        if (idx != nk) { rc = mb_fail(MB_EFORMAT, "internal: block tables of '%s' disagree", in[i].dir.c_str()); goto done; }
        UnionBlock *dTab = (UnionBlock *)g.take(tab.size() * sizeof(UnionBlock) + 16);
        if (!dTab) { rc = mb_fail(MB_ENOMEM, "internal: union block table does not fit"); goto done; }
        if (!tab.empty()) {
          //  this input's files of the group, back to back (8-aligned), chunk by chunk
          size_t jFirst = 0;
          while (jobs[jFirst].input != i) jFirst++;
          const size_t jEnd = jFirst + (f1 - f0);
          rc = streamToDevice(ring, dRaw, inBytes[i], s, [&](size_t off, size_t len, uint8_t *dst) -> int {
            const size_t PIECE = ring.PIECE;
            const size_t nPieces = (len + PIECE - 1) / PIECE;
            std::atomic<int> bad(-1);
            parallelFor(nPieces, [&](size_t pi) {
              const size_t lo = off + pi * PIECE, hi = std::min(off + len, lo + PIECE);
              for (size_t ji = jFirst; ji < jEnd; ji++) {
                const FileJob &j = jobs[ji];
                const size_t a0 = std::max(lo, j.off), a1 = std::min(hi, j.off + j.bytes);
                if (a0 >= a1) continue;
                uint8_t *to = dst + (a0 - off);
                if (in[i].mem) { memcpy(to, in[i].mem->dataPtr[j.file] + (a0 - j.off), a1 - a0); continue; }
                size_t got = 0;
                while (got < a1 - a0) {
                  const ssize_t r = pread(j.fd, to + got, a1 - a0 - got, (off_t)(a0 - j.off + got));
                  if (r <= 0) { bad = (int)ji; return; }
                  got += (size_t)r;
                }
              }
            });
            if (bad >= 0) return mb_fail(MB_EIO, "cannot read data file %u of '%s'", jobs[bad].file, in[i].dir.c_str());
            return MB_OK;
          });
          if (rc != MB_OK) goto done;
          MB_UCUDA(cudaMemcpyAsync(dTab, tab.data(), tab.size() * sizeof(UnionBlock), cudaMemcpyHostToDevice, s));
          //  many short blocks (a count db at prefix width 24 has 16.8 M of them): one thread per block fills the device;
          //  few long blocks (stream-writer output, small count dbs): one warp per block.  MB_DECODE=thread|warp|auto (auto: by
          //  the block count); the warp form is green on the B200 but measured no faster at 32 K blocks (259 vs 280 ms of
          //  read + decode for 1.6 GB), so it stays opt-in until an input with 4096 long blocks has been timed.
          const char *dm = getenv("MB_DECODE");
          const bool byWarp = dm && (!strcmp(dm, "warp") || (!strcmp(dm, "auto") && tab.size() < ((size_t)1 << 18)));
          if (byWarp) {
            const unsigned grid = (unsigned)((tab.size() + DW_WARPS - 1) / DW_WARPS);
            if (kw == 1) decode_warp_kernel<1><<<grid, DW_WARPS * 32, 0, s>>>(dRaw, inBytes[i], dTab, tab.size(), in[i].wSuffix, (u64 *)dKey, dVal, idx, dErr);
            else         decode_warp_kernel<2><<<grid, DW_WARPS * 32, 0, s>>>(dRaw, inBytes[i], dTab, tab.size(), in[i].wSuffix, (ulonglong2 *)dKey, dVal, idx, dErr);
          } else {
            const unsigned grid = (unsigned)((tab.size() + 127) / 128);
            if (kw == 1) decode_kernel<1><<<grid, 128, 0, s>>>(dRaw, inBytes[i], dTab, tab.size(), in[i].wSuffix, (u64 *)dKey, dVal, idx, dErr);
            else         decode_kernel<2><<<grid, 128, 0, s>>>(dRaw, inBytes[i], dTab, tab.size(), in[i].wSuffix, (ulonglong2 *)dKey, dVal, idx, dErr);
          }
          MB_UCUDA(cudaGetLastError());
          st.nLaunches++;
          st.h2dBytes += tab.size() * sizeof(UnionBlock);
        }
        lists[i] = List(dKey, dVal);  listN[i] = idx;
        st.nInputKmers += idx;
      }
      {
        u32 err = 0;
        MB_UCUDA(cudaMemcpyAsync(&err, dErr, 4, cudaMemcpyDeviceToHost, s));
        MB_UCUDA(cudaEventRecord(ev[1], s));
        MB_UCUDA(cudaStreamSynchronize(s));
        ring.quiesce();
        for (FileJob &j : jobs) if (j.fd >= 0) { close(j.fd); j.fd = -1; }
        if (err) { rc = mb_fail(MB_EFORMAT, "corrupt block in the inputs (decoder flags 0x%x)", err); goto done; }
        float ms = 0;  cudaEventElapsedTime(&ms, ev[0], ev[1]);  st.msDecode += ms;
      }

      //  ---- fold the lists: ((A + B) + C) + ... ; a lone input is merged with an empty list (drops zero values) ----
      u64 total = 0;
      for (u32 i = 0; i < nInputs; i++) total += listN[i];
      const bool pingPong = nInputs > 2;                     //  one fold step writes one buffer
      void *mk[2] = { g.take(total * W + 16), pingPong ? g.take(total * W + 16) : nullptr };
      u32  *mv[2] = { (u32 *)g.take(total * 4 + 16), pingPong ? (u32 *)g.take(total * 4 + 16) : nullptr };
      const u64 maxTiles = total / MG_TILE + 2;
      u64 *dStatus = (u64 *)g.take((maxTiles + 2) * 8);
      if (!mk[0] || !mv[0] || (pingPong && (!mk[1] || !mv[1])) || !dStatus) { rc = mb_fail(MB_ENOMEM, "internal: union merge buffers do not fit"); goto done; }

      MB_UCUDA(cudaEventRecord(ev[0], s));
      const void *curK = lists[0].first;  const u32 *curV = lists[0].second;  u64 curN = listN[0];
      int pp = 0;
      for (u32 i = 1; i <= nInputs; i++) {
        if (i == nInputs && nInputs > 1) break;
        const void *bk = (i < nInputs) ? lists[i].first : nullptr;
        const u32  *bv = (i < nInputs) ? lists[i].second : nullptr;
        const u64   bn = (i < nInputs) ? listN[i] : 0;
        //  a lone input: union counts it (value 1), every other operator passes its values through
        MergeRule rule;
        rule.op       = (nInputs == 1 && op != MB_OP_UNION && op <= MB_OP_SYMMETRIC_DIFFERENCE) ? (u32)MB_OP_UNION_SUM : op;
        rule.param    = param;
        rule.param2   = E.param2;
        rule.firstRaw = (i == 1) ? 1 : 0;
        rule.dropZero = (op != MB_OP_INTERSECT_SUM || i + 1 >= nInputs) ? 1 : 0;
        const u64 tiles = (curN + bn + MG_TILE - 1) / MG_TILE;
        MB_UCUDA(cudaMemsetAsync(dCtl + 8, 0, 8, s));
        if (tiles) {
          MB_UCUDA(cudaMemsetAsync(dStatus, 0, (tiles + 2) * 8, s));
          if (kw == 1) merge_op_kernel<1><<<(unsigned)tiles, MG_THREADS, 0, s>>>((const u64 *)curK, curV, curN, (const u64 *)bk, bv, bn, (u64 *)mk[pp], mv[pp], dCtl + 8, dStatus + 2, dStatus, rule);
          else         merge_op_kernel<2><<<(unsigned)tiles, MG_THREADS, 0, s>>>((const ulonglong2 *)curK, curV, curN, (const ulonglong2 *)bk, bv, bn, (ulonglong2 *)mk[pp], mv[pp], dCtl + 8, dStatus + 2, dStatus, rule);
          MB_UCUDA(cudaGetLastError());
          st.nLaunches++;
        }
        u64 nOut = 0;
        MB_UCUDA(cudaMemcpyAsync(&nOut, dCtl + 8, 8, cudaMemcpyDeviceToHost, s));
        MB_UCUDA(cudaStreamSynchronize(s));
        curK = mk[pp];  curV = mv[pp];  curN = nOut;
        pp ^= 1;
      }
      MB_UCUDA(cudaEventRecord(ev[1], s));
      MB_UCUDA(cudaStreamSynchronize(s));
      { float ms = 0;  cudaEventElapsedTime(&ms, ev[0], ev[1]);  st.msMerge += ms; }
      const u64 D = curN;
      st.nOutputKmers += D;

      //  ---- value histogram + encode (stream-writer rules) + write the group's files -------------------------------
      MB_UCUDA(cudaEventRecord(ev[0], s));
      const u64 histBigCap = D / MB_HIST_SMALL + 1024;
      u32 *dHistBig = (u32 *)g.take(histBigCap * 4);
      const u64 nBlocks = (u64)(f1 - f0) * blocksPerFile;
      MbBlockMeta *dMeta = (MbBlockMeta *)g.take(nBlocks * sizeof(MbBlockMeta));
      void *dScanTmp     = g.take(mb_encode_scan_tmp_bytes(nBlocks));
      u64  *dIdx         = (u64 *)g.take(nBlocks * 24);
      if (!dHistBig || !dMeta || !dScanTmp || !dIdx) { rc = mb_fail(MB_ENOMEM, "internal: union encode tables do not fit"); goto done; }
      MB_UCUDA(cudaMemsetAsync(dCtl + 9, 0, 8, s));
      MB_UCUDA(cudaMemcpyAsync(dCtl + 8, &D, 8, cudaMemcpyHostToDevice, s));
      MB_UCUDA(mb_launch_value_histogram_counts(curV, dCtl + 8, D, dHistSmall, dHistBig, dCtl + 9, histBigCap, dCtl + 10, s));

      MbEncodeParams ep;
      ep.keys = curK;  ep.runStart = nullptr;  ep.nDistinct = D;  ep.keyWords = kw;  ep.k = k;  ep.wPrefix = wP;
      ep.blockLo = (u64)f0 * blocksPerFile;  ep.nBlocks = nBlocks;  ep.counts = curV;  ep.streamMode = E.streamMode;
      if (E.streamMode) {                                      //  prefixes of more than 16 Mi k-mers are written as several blocks
        ep.streamBlockMax = streamBlockMax();
        ep.parts = (MbSplitPart *)g.take((2 * (D / ep.streamBlockMax) + 4) * sizeof(MbSplitPart));
        if (!ep.parts) { rc = mb_fail(MB_ENOMEM, "internal: union encode tables do not fit"); goto done; }
      }
      MB_UCUDA(cudaMemsetAsync(dCtl, 0, 8 * 8, s));
      MB_UCUDA(mb_launch_encode_layout(ep, dMeta, dCtl, dScanTmp, s));
      u64 encTotals[8];
      MB_UCUDA(cudaMemcpyAsync(encTotals, dCtl, sizeof(encTotals), cudaMemcpyDeviceToHost, s));
      std::vector<MbBlockMeta> meta(f1 - f0);                  //  of the first block of every file of the group
      for (u32 f = f0; f < f1; f++)
        MB_UCUDA(cudaMemcpyAsync(&meta[f - f0], dMeta + (u64)(f - f0) * blocksPerFile, sizeof(MbBlockMeta), cudaMemcpyDeviceToHost, s));
      MB_UCUDA(cudaStreamSynchronize(s));
      st.nLaunches += 6;

      uint8_t *dOut = (uint8_t *)g.take(encTotals[0] + 16);
      u32     *dSub = (u32 *)g.take(encTotals[1] * 4 + 16);
      if (!dOut || !dSub) { rc = mb_fail(MB_ENOMEM, "internal: the encoded group (%llu bytes) does not fit", (unsigned long long)encTotals[0]); goto done; }
      MB_UCUDA(cudaMemsetAsync(dOut, 0, encTotals[0], s));
      MbEncodeChunk ch = { 0, nBlocks, 0, 0 };
      MB_UCUDA(mb_launch_encode_write(ep, ch, dMeta, dCtl, dSub, dOut, dIdx, dFileBase, s));
      st.nLaunches += 2;

      //  ---- copy back + write: the encoded bytes stream through the page-locked ring, host threads write one chunk to
      //  the files (or the images) while the next one is on the bus --------------------------------------------------------
      if (!outIdx.need(nBlocks * 24)) { rc = mb_fail(MB_ENOMEM, "cannot allocate %llu bytes of host memory", (unsigned long long)(nBlocks * 24)); goto done; }
      MB_UCUDA(cudaMemcpyAsync(outIdx.data(), dIdx, nBlocks * 24, cudaMemcpyDeviceToHost, s));
      {
        const u32 nf = f1 - f0;
        std::vector<u64> fileByte(nf + 1);                     //  the files' places in the group's encoded bytes
        for (u32 fi = 0; fi < nf; fi++) fileByte[fi] = meta[fi].outByte;
        fileByte[nf] = encTotals[0];
        std::vector<int>       fds(nf, -1);
        std::vector<uint8_t *> dest(nf, nullptr);              //  images in host memory (merge of count batches)
        int wrc = MB_OK;
        for (u32 fi = 0; fi < nf && wrc == MB_OK; fi++) {
          const u32 f = f0 + fi;
          const size_t len = (size_t)(fileByte[fi + 1] - fileByte[fi]);
          if (E.outImages) {
            MbDbImages &o = *E.outImages;
            if (E.outBuf) {
              if (*E.outUsed + len > E.outCap) { wrc = mb_fail(MB_ENOMEM, "output buffer too small: data file %u needs %zu more bytes, %zu of %zu used", f, len, *E.outUsed, E.outCap); break; }
              o.dataVec[f].clear();  dest[fi] = E.outBuf + *E.outUsed;  *E.outUsed += len;
            } else {
              o.dataVec[f].resize(len);  dest[fi] = o.dataVec[f].data();
            }
            o.dataPtr[f] = dest[fi];  o.dataLen[f] = len;
          } else {
            char path[4096];
            mb_file_name(path, sizeof(path), outDir, f, false);
            fds[fi] = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0666);
            if (fds[fi] < 0) wrc = mb_fail(MB_EIO, "cannot create '%s': %s", path, strerror(errno));
          }
        }
        if (wrc == MB_OK && !E.outImages) {
          //  a directory: one job per data file, the files in parallel (filesink.h: several host threads on ONE file gain
          //  nothing -- measured 3.4 GB/s for this step when the pieces of a 32 MB chunk were written by eight threads)
          MbFileSink *sink = mb_filesink_get();
          int dev = 0;  cudaGetDevice(&dev);
          if (cudaStreamSynchronize(s) != cudaSuccess) wrc = mb_fail(MB_ENODEVICE, "%s", cudaGetErrorString(cudaGetLastError()));
          for (u32 fi = 0; fi < nf && wrc == MB_OK; fi++)
            mb_filesink_submit(sink, dev, fds[fi], 0, dOut + fileByte[fi], (size_t)(fileByte[fi + 1] - fileByte[fi]), nullptr);
          const int drc = mb_filesink_drain(sink);
          if (wrc == MB_OK) wrc = drc;
        } else if (wrc == MB_OK)
          wrc = streamFromDevice(ring, dOut, (size_t)encTotals[0], s, [&](size_t off, size_t len, const uint8_t *src) -> int {
            //  pieces: (file, range) intersections of this chunk, at most 4 MB each
            struct Piece { u32 fi; size_t a0, a1; };
            std::vector<Piece> pieces;
            const size_t PIECE = ring.PIECE;
            u32 fi = (u32)(std::upper_bound(fileByte.begin(), fileByte.end(), (u64)off) - fileByte.begin());
            fi = fi ? fi - 1 : 0;
            for (; fi < nf && fileByte[fi] < off + len; fi++) {
              const size_t a0 = std::max<size_t>(off, fileByte[fi]), a1 = std::min<size_t>(off + len, fileByte[fi + 1]);
              for (size_t p = a0; p < a1; p += PIECE) pieces.push_back(Piece{ fi, p, std::min(a1, p + PIECE) });
            }
            std::atomic<int> bad(-1);
            parallelFor(pieces.size(), [&](size_t pi) {
              const Piece &pc = pieces[pi];
              const uint8_t *from = src + (pc.a0 - off);
              const size_t inFile = pc.a0 - (size_t)fileByte[pc.fi];
              if (dest[pc.fi]) { memcpy(dest[pc.fi] + inFile, from, pc.a1 - pc.a0); return; }
              size_t put = 0;
              while (put < pc.a1 - pc.a0) {
                const ssize_t r = pwrite(fds[pc.fi], from + put, pc.a1 - pc.a0 - put, (off_t)(inFile + put));
                if (r <= 0) { bad = (int)pc.fi; return; }
                put += (size_t)r;
              }
            });
            if (bad >= 0) return mb_fail(MB_EIO, "cannot write data file %u in '%s': %s", f0 + (u32)bad, outDir, strerror(errno));
            return MB_OK;
          });
        for (u32 fi = 0; fi < nf; fi++) if (fds[fi] >= 0 && close(fds[fi]) != 0 && wrc == MB_OK) wrc = mb_fail(MB_EIO, "cannot close data file %u in '%s': %s", f0 + fi, outDir, strerror(errno));
        if (wrc != MB_OK) { rc = wrc; goto done; }
        MB_UCUDA(cudaEventRecord(ev[1], s));
        MB_UCUDA(cudaStreamSynchronize(s));
        { float ms = 0;  cudaEventElapsedTime(&ms, ev[0], ev[1]);  st.msEncode += ms; }
        st.d2hBytes += encTotals[0] + nBlocks * 24;

        //  the block indexes (small)
        for (u32 fi = 0; fi < nf; fi++) {
          const u32 f = f0 + fi;
          const uint8_t *ix = outIdx.data() + (u64)fi * blocksPerFile * 24;
          if (E.outImages) E.outImages->index[f].assign(ix, ix + blocksPerFile * 24);
          else {
            char path[4096];
            mb_file_name(path, sizeof(path), outDir, f, true);
            MB_UTRY(writeFile(path, ix, (size_t)blocksPerFile * 24));
          }
          st.outBytes += fileByte[fi + 1] - fileByte[fi] + blocksPerFile * 24;
        }
      }

      //  big histogram values of this group
      u64 bigN = 0;
      MB_UCUDA(cudaMemcpy(&bigN, dCtl + 9, 8, cudaMemcpyDeviceToHost));
      if (bigN > histBigCap) { rc = mb_fail(MB_ENOMEM, "internal: value histogram overflow list too small"); goto done; }
      if (bigN) {
        std::vector<u32> big(bigN);
        MB_UCUDA(cudaMemcpy(big.data(), dHistBig, bigN * 4, cudaMemcpyDeviceToHost));
        for (u32 v : big) histMap[v]++;
      }
      f0 = f1;
    }

    //  ---- master index --------------------------------------------------------------------------------------
    std::vector<u64> hs(MB_HIST_SMALL);
    u64 totals[3];
    MB_UCUDA(cudaMemcpy(hs.data(), dHistSmall, (size_t)MB_HIST_SMALL * 8, cudaMemcpyDeviceToHost));
    MB_UCUDA(cudaMemcpy(totals, dCtl + 10, sizeof(totals), cudaMemcpyDeviceToHost));
    st.d2hBytes += (size_t)MB_HIST_SMALL * 8;
    nUnique = totals[0];  nDistinct = totals[1];  nTotal = totals[2];
    for (u64 v = 1; v < MB_HIST_SMALL; v++) if (hs[v]) { hist.values.push_back(v); hist.occ.push_back(hs[v]); }
    for (auto &kv : histMap) { hist.values.push_back(kv.first); hist.occ.push_back(kv.second); }
    hist.nUnique = nUnique;  hist.nDistinct = nDistinct;  hist.nTotal = nTotal;
    if (E.outHist) *E.outHist = hist;
    if (outDir) {
      std::vector<uint8_t> master;
      mb_build_master_index(k, wP, hist, master);
      MB_UTRY(writeFile(std::string(outDir) + "/merylIndex", master.data(), master.size()));
      st.outBytes += master.size();
    }
    st.nUnique = nUnique;  st.nDistinct = nDistinct;  st.nTotal = nTotal;

    MB_UCUDA(cudaEventRecord(evAll[1], s));
    MB_UCUDA(cudaStreamSynchronize(s));
    { float ms = 0;  cudaEventElapsedTime(&ms, evAll[0], evAll[1]);  st.msTotal = ms; }
  }

done:
  for (int i = 0; i < 2; i++) { if (ev[i]) cudaEventDestroy(ev[i]); if (evAll[i]) cudaEventDestroy(evAll[i]); }
  if (s) cudaStreamDestroy(s);
  if (arena && !E.arena) cudaFree(arena);
  if (rc == MB_OK && stOut) *stOut = st;
  return rc;
}
