//  K6: Elias-Fano / stuffedBits block encoder, bit-exact with the reference writer.
//
//  Replaces merylFileWriter::writeBlockToFile (src/utility/src/kmers-v1/kmers-writer.C:183-291),
//  the stuffedBits container it writes through (src/utility/src/bits/stuffedBits-v1.H:298-372,
//  -binary.C:95-147, -unary.C:85-163, stuffedBits-v1.C:65-111) and merylFileIndex
//  (src/utility/src/kmers-v1/kmers-files.H:110-155).
//
//  A block holds the n distinct k-mers that share a wPrefix-bit prefix.  With u = ceil(log2 n)
//  and b = wSuffix - u, its bit stream is a sequence of FIELDS
//      header (11 fields, 528 bits) | per k-mer: unary(hi_i - hi_{i-1}), binary(max(b-64,0)), binary(min(b,64)) | n x binary(32) counts
//  (hi_i = suffix_i >> b) and the container cuts the stream into sub-blocks of at most maxBits
//  bits such that no field straddles a cut: a field that does not fit closes the sub-block
//  where it is.  Every sub-block is stored from a fresh 64-bit word.
//
//  Closed forms make this parallel.  In "virtual" coordinates (no cuts) k-mer i's unary field
//  starts at  U(i) = 528 + hi_{i-1} + i*(1+b),  its binary fields right after the unary's d_i+1
//  bits, count i at U(n) + 32 i.  Cuts are a short greedy chain: next cut = the largest field
//  start <= previous cut + maxBits (a predecessor search: binary search over i on U(i)); a block
//  has ~5-20 of them regardless of n.  The serialised sub-block table bgn[] IS the list of cuts in
//  virtual coordinates, so once it is written every field finds its place independently:
//      physical bit = 64*wordBase[j] + (virtual - bgn[j]),   j = sub-block holding the field start.
//
//  Kernels:   layout_kernel   thread per block: extent (binary search), n, chain -> nSub, bytes
//             scan kernels    exclusive scan of (bytes, nSub) over the blocks
//             header_kernel   thread per block: container header, bgn[]/len[], wordBase[], the
//                             528-bit block header, the 24-byte index entry
//             fields_kernel   thread per distinct k-mer: its unary bit, binary bits and count,
//                             OR-ed into the zeroed output with 64-bit atomics
//
//  Stream-writer mode (set operations): a prefix without k-mers writes nothing, and a prefix with more than 16 Mi k-mers is
//  written as consecutive blocks (parts) of 16 Mi k-mers under one index entry, like merylStreamWriter::addMer
//  (kmers-writer-stream.C:127-141): each part is laid out as a block of its own, MbSplitPart keeps its place.
//
//  Bytes: reads (W+8) per distinct k-mer (+ the chain's O(log n) probes), writes ~(b+34)/8.

#include "kernels.cuh"

namespace {

const u64 HDR_BITS = 528;

struct BlockShape {
  u32 u, b, ls, rs;
  u64 maxBits;
};

__device__ __forceinline__ BlockShape block_shape(u64 n, u32 wSuffix) {
  BlockShape s;
  s.u = 0;
  u64 usum = 1;
  while (usum < n) { s.u++; usum <<= 1; }
  s.b  = wSuffix - s.u;
  s.ls = s.b <= 64 ? 0 : s.b - 64;
  s.rs = s.b <= 64 ? s.b : 64;
  u64 bs = 10 * 64 + 2 * usum + n * s.b / 16 + n * 32 / 16;
  s.maxBits = (bs & 0xfffffffffffffc00ull) + 1024;
  return s;
}

//  hi part of the suffix: (key >> b) & (2^u - 1)
template <typename Key>
__device__ __forceinline__ u64 key_hi_part(const Key &key, u32 b, u32 u) {
  const u64 v = key_shr64(key, b);                         //  2k - b <= wPrefix + u bits: the caller keeps this <= 64
  return u >= 64 ? v : (v & (((u64)1 << u) - 1));
}

//  Virtual start of k-mer i's unary field, i in [0, n]; U(n) is the start of the count region.
template <typename Key>
__device__ __forceinline__ u64 ustart(const Key *keys, u64 first, u64 i, const BlockShape &s) {
  const u64 hprev = (i == 0) ? 0 : key_hi_part(keys[first + i - 1], s.b, s.u);
  return HDR_BITS + hprev + i * (1 + (u64)s.b);
}

//  Largest field start <= T  (T >= HDR_BITS, T < total).
template <typename Key>
__device__ u64 field_start_at_or_before(const Key *keys, u64 first, u64 n, const BlockShape &s, u64 uN, u64 T) {
  if (T >= uN) {
    u64 q = (T - uN) / 32;
    if (q > n) q = n;
    return uN + 32 * q;
  }
  u64 lo = 0, hi = n;                                      //  U(lo) <= T < U(hi)
  while (hi - lo > 1) {
    const u64 mid = (lo + hi) >> 1;
    if (ustart(keys, first, mid, s) <= T) lo = mid; else hi = mid;
  }
  const u64 Ui = ustart(keys, first, lo, s);
  const u64 h  = key_hi_part(keys[first + lo], s.b, s.u);
  const u64 c1 = HDR_BITS + h + lo * (1 + (u64)s.b) + 1;   //  first binary field (after the unary's 1)
  const u64 c2 = c1 + s.ls;                                //  second binary field
  u64 E = Ui;
  if (c1 <= T)             E = c1;
  if (s.ls && c2 <= T)     E = c2;
  return E;
}

//  Walks the greedy chain of sub-blocks of one block; calls f(j, bgn, len) for each.  Returns nSub.
template <typename Key, typename F>
__device__ u32 walk_subblocks(const Key *keys, u64 first, u64 n, const BlockShape &s, F f) {
  const u64 uN    = ustart(keys, first, n, s);
  const u64 total = uN + 32 * n;
  u64 S = 0;
  u32 j = 0;
  while (total - S > s.maxBits) {
    const u64 E = field_start_at_or_before(keys, first, n, s, uN, S + s.maxBits);
    f(j, S, E - S);
    S = E;
    j++;
  }
  f(j, S, total - S);
  return j + 1;
}

//  First index in [0, D) whose (key >> wSuffix) >= prefix.
template <typename Key>
__device__ __forceinline__ u64 lower_bound_prefix(const Key *keys, u64 D, u32 wSuffix, u64 prefix) {
  u64 lo = 0, hi = D;
  while (lo < hi) {
    const u64 mid = (lo + hi) >> 1;
    if (key_shr64(keys[mid], wSuffix) < prefix) lo = mid + 1; else hi = mid;
  }
  return lo;
}

//  A stream-writer prefix that is written as several blocks (MbEncodeParams::streamBlockMax).
__device__ __forceinline__ bool is_split(const MbEncodeParams &p, u64 n) { return p.streamMode && n > p.streamBlockMax; }

//  ---- layout ---------------------------------------------------------------------------------

template <int KW>
__global__ void __launch_bounds__(128)
layout_kernel(MbEncodeParams p, MbBlockMeta *__restrict__ meta, u64 *__restrict__ totals) {
  typedef typename KeyT<KW>::type Key;
  const u64 blk = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (blk >= p.nBlocks) return;

  const Key *keys    = (const Key *)p.keys;
  const u32  wSuffix = 2 * p.k - p.wPrefix;
  const u64  prefix  = p.blockLo + blk;
  const u64  first   = lower_bound_prefix(keys, p.nDistinct, wSuffix, prefix);
  const u64  end     = lower_bound_prefix(keys, p.nDistinct, wSuffix, prefix + 1);
  const u64  n       = end - first;

  const BlockShape s = block_shape(n, wSuffix);
  u64 words = 0;
  u32 nSub  = 0;
  const bool skip = p.streamMode && n == 0;                 //  the stream writer never writes a block without k-mers
  u64 bytes = 0;
  if (is_split(p, n)) {
    //  merylStreamWriter::addMer dumps the batch when it holds streamBlockMax k-mers and goes on with the same prefix:
    //  every part is a block of its own (own shape, own sub-block chain); the parts follow each other in the file
    MbSplitPart *part = p.parts + 2 * (first / p.streamBlockMax);
    for (u64 at = 0; at < n; at += p.streamBlockMax, part++) {
      const u64 np = (n - at < p.streamBlockMax) ? n - at : p.streamBlockMax;
      const BlockShape sp = block_shape(np, wSuffix);
      u64 w = 0;
      const u32 ns = walk_subblocks(keys, first + at, np, sp, [&](u32, u64, u64 len) { w += (len + 63) >> 6; });
      part->relByte = bytes;  part->relSub = nSub;  part->nSub = ns;
      bytes += 16 + 16 * (u64)ns + 8 * w;
      nSub  += ns;
    }
  } else if (!skip) {
    nSub  = walk_subblocks(keys, first, n, s, [&](u32, u64, u64 len) { words += (len + 63) >> 6; });
    bytes = 16 + 16 * (u64)nSub + 8 * words;
  }

  MbBlockMeta m;
  m.first    = first;
  m.outByte  = 0;
  m.nSub     = nSub;
  m.subFirst = 0;
  m.bytes    = bytes;
  meta[blk]  = m;

  if (blk == 0)             totals[2] = first;             //  the chunk's distinct k-mers are [totals[2], totals[3])
  if (blk == p.nBlocks - 1) totals[3] = end;
}

//  ---- exclusive scan of (bytes, nSub) over the blocks: three small kernels ----------------------

const int SCAN_THREADS = 256;
const int SCAN_ITEMS   = 4;
const int SCAN_TILE    = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ void block_excl_scan2(u64 &a, u64 &b, u64 &totA, u64 &totB) {
  //  exclusive scan across the CTA of one (a, b) pair per thread
  __shared__ u64 wa[SCAN_THREADS / 32], wb[SCAN_THREADS / 32];
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u64 ia = a, ib = b;
  for (int o = 1; o < 32; o <<= 1) {
    const u64 va = __shfl_up_sync(0xffffffffu, ia, o), vb = __shfl_up_sync(0xffffffffu, ib, o);
    if (lane >= (u32)o) { ia += va; ib += vb; }
  }
  if (lane == 31) { wa[warp] = ia; wb[warp] = ib; }
  __syncthreads();
  u64 ba = 0, bb = 0, ta = 0, tb = 0;
  for (u32 w = 0; w < SCAN_THREADS / 32; w++) {
    if (w < warp) { ba += wa[w]; bb += wb[w]; }
    ta += wa[w]; tb += wb[w];
  }
  a = ba + ia - a;  b = bb + ib - b;
  totA = ta;  totB = tb;
  __syncthreads();
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_sums_kernel(const MbBlockMeta *__restrict__ meta, u64 nBlocks, u64 *__restrict__ tileSums) {
  const u64 base = (u64)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  u64 a = 0, b = 0;
  for (int i = 0; i < SCAN_ITEMS; i++)
    if (base + i < nBlocks) { a += meta[base + i].bytes; b += meta[base + i].nSub; }
  u64 ta, tb;
  block_excl_scan2(a, b, ta, tb);
  if (threadIdx.x == 0) { tileSums[2 * (u64)blockIdx.x] = ta; tileSums[2 * (u64)blockIdx.x + 1] = tb; }
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_offsets_kernel(u64 *__restrict__ tileSums, u64 nTiles, u64 *__restrict__ totals) {
  //  one CTA; sequential over chunks of SCAN_THREADS tiles
  __shared__ u64 carryA, carryB;
  if (threadIdx.x == 0) { carryA = 0; carryB = 0; }
  __syncthreads();
  for (u64 c = 0; c < nTiles; c += SCAN_THREADS) {
    const u64 i = c + threadIdx.x;
    u64 a = i < nTiles ? tileSums[2 * i] : 0, b = i < nTiles ? tileSums[2 * i + 1] : 0;
    u64 ta, tb;
    block_excl_scan2(a, b, ta, tb);
    if (i < nTiles) { tileSums[2 * i] = carryA + a; tileSums[2 * i + 1] = carryB + b; }
    __syncthreads();
    if (threadIdx.x == 0) { carryA += ta; carryB += tb; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { totals[0] = carryA; totals[1] = carryB; }
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(MbBlockMeta *__restrict__ meta, u64 nBlocks, const u64 *__restrict__ tileSums) {
  const u64 base = (u64)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  u64 va[SCAN_ITEMS], vb[SCAN_ITEMS];
  u64 a = 0, b = 0;
  for (int i = 0; i < SCAN_ITEMS; i++) {
    va[i] = 0; vb[i] = 0;
    if (base + i < nBlocks) { va[i] = meta[base + i].bytes; vb[i] = meta[base + i].nSub; }
    a += va[i]; b += vb[i];
  }
  u64 ta, tb;
  block_excl_scan2(a, b, ta, tb);
  a += tileSums[2 * (u64)blockIdx.x];
  b += tileSums[2 * (u64)blockIdx.x + 1];
  for (int i = 0; i < SCAN_ITEMS; i++) {
    if (base + i < nBlocks) { meta[base + i].outByte = a; meta[base + i].subFirst = (u32)b; }
    a += va[i]; b += vb[i];
  }
}

//  ---- bit scatter helpers ----------------------------------------------------------------------

//  OR `width` (1..64) bits of value into the MSB-first stream `words` at bit position pos.
__device__ __forceinline__ void or_bits(u64 *words, u64 pos, u32 width, u64 value) {
  if (width < 64) value &= (((u64)1 << width) - 1);
  const u64 w    = pos >> 6;
  const u32 room = 64 - (u32)(pos & 63);
  if (width <= room) {
    atomicOr((unsigned long long *)&words[w], (unsigned long long)(value << (room - width)));
  } else {
    atomicOr((unsigned long long *)&words[w],     (unsigned long long)(value >> (width - room)));
    atomicOr((unsigned long long *)&words[w + 1], (unsigned long long)(value << (64 - (width - room))));
  }
}

//  Container header, sub-block tables and the 528-bit block header of one serialised block of n k-mers (keys[first ...]).
template <typename Key>
__device__ void write_block_head(const Key *keys, u64 first, u64 n, const BlockShape &s, u64 prefix, u32 nSub, u64 *o, u32 *sw) {
  u64 *bgn  = o + 2;
  u64 *len  = bgn + nSub;
  u64 *data = len + nSub;

  o[0] = s.maxBits;
  o[1] = (u64)nSub | ((u64)(32 * ((nSub - 1) / 32 + 1)) << 32);

  u32 wordBase = 0;
  walk_subblocks(keys, first, n, s, [&](u32 j, u64 b0, u64 l) {
    bgn[j] = b0;
    len[j] = l;
    sw[j]  = wordBase;
    wordBase += (u32)((l + 63) >> 6);
  });

  //  the 528 header bits live at the start of sub-block 0 (maxBits >= 1024)
  atomicOr((unsigned long long *)&data[0], 0x7461446c7972656dull);
  atomicOr((unsigned long long *)&data[1], 0x0a3030656c694661ull);
  atomicOr((unsigned long long *)&data[2], (unsigned long long)prefix);
  atomicOr((unsigned long long *)&data[3], (unsigned long long)n);
  atomicOr((unsigned long long *)&data[4], (1ull << 56) | ((u64)s.u << 24) | ((u64)s.b >> 8));
  atomicOr((unsigned long long *)&data[5], ((u64)s.b & 0xff) << 56);
  atomicOr((unsigned long long *)&data[6], 1ull << 48);
}

//  ---- per-block headers ------------------------------------------------------------------------

template <int KW>
__global__ void __launch_bounds__(128)
header_kernel(MbEncodeParams p, MbEncodeChunk c, const MbBlockMeta *__restrict__ meta, const u64 *__restrict__ totals,
              u32 *__restrict__ subWord, uint8_t *__restrict__ out, u64 *__restrict__ index, const u64 *__restrict__ fileBase) {
  typedef typename KeyT<KW>::type Key;
  const u64 rel = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (rel >= c.chunkCount) return;
  const u64 blk = c.chunkFirst + rel;

  const Key *keys    = (const Key *)p.keys;
  const u32  wSuffix = 2 * p.k - p.wPrefix;
  const u64  prefix  = p.blockLo + blk;
  const MbBlockMeta m = meta[blk];
  const u64  end     = (blk + 1 < p.nBlocks) ? meta[blk + 1].first : totals[3];
  const u64  n       = end - m.first;
  const BlockShape s = block_shape(n, wSuffix);

  if (p.streamMode && n == 0) {                              //  no block; merylFileIndex::clear() (kmers-files.H:141-145)
    index[3 * rel + 0] = 0;
    index[3 * rel + 1] = ~0ull;
    index[3 * rel + 2] = 0;
    return;
  }

  if (is_split(p, n)) {
    const MbSplitPart *part = p.parts + 2 * (m.first / p.streamBlockMax);
    for (u64 at = 0; at < n; at += p.streamBlockMax, part++) {
      const u64 np = (n - at < p.streamBlockMax) ? n - at : p.streamBlockMax;
      write_block_head(keys, m.first + at, np, block_shape(np, wSuffix), prefix, part->nSub,
                       (u64 *)(out + (m.outByte + part->relByte - c.chunkByte0)), subWord + (m.subFirst + part->relSub - c.chunkSub0));
    }
  } else
    write_block_head(keys, m.first, n, s, prefix, m.nSub, (u64 *)(out + (m.outByte - c.chunkByte0)), subWord + (m.subFirst - c.chunkSub0));

  //  index entry: {prefix, byte position in its data file, nKmers}
  const u32 blockBits   = p.wPrefix - 6;
  const u64 file        = prefix >> blockBits;
  u64       fileFirst   = file << blockBits;                 //  first block prefix of the file ...
  if (fileFirst < p.blockLo) fileFirst = p.blockLo;          //  ... that is part of this chunk
  const u64 fileFirstByte = meta[fileFirst - p.blockLo].outByte;
  index[3 * rel + 0] = prefix;
  index[3 * rel + 1] = m.outByte - fileFirstByte + fileBase[file];
  index[3 * rel + 2] = n;
}

//  ---- per-k-mer fields -------------------------------------------------------------------------

template <int KW>
__global__ void __launch_bounds__(256)
fields_kernel(MbEncodeParams p, MbEncodeChunk c, const MbBlockMeta *__restrict__ meta, const u64 *__restrict__ totals,
              const u32 *__restrict__ subWord, uint8_t *__restrict__ out) {
  typedef typename KeyT<KW>::type Key;
  const u64 passEnd = totals[3];
  const u64 gEnd = (c.chunkFirst + c.chunkCount < p.nBlocks) ? meta[c.chunkFirst + c.chunkCount].first : passEnd;
  const u64 g    = meta[c.chunkFirst].first + (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= gEnd) return;

  const Key *keys    = (const Key *)p.keys;
  const u32  wSuffix = 2 * p.k - p.wPrefix;
  const Key  key     = keys[g];
  const u64  blk     = key_shr64(key, wSuffix) - p.blockLo;
  MbBlockMeta m = meta[blk];
  const u64  end     = (blk + 1 < p.nBlocks) ? meta[blk + 1].first : passEnd;
  u64        n       = end - m.first;
  if (is_split(p, n)) {                                      //  the k-mer's part of an over-full prefix is its block
    const u64 q = (g - m.first) / p.streamBlockMax;
    const MbSplitPart part = p.parts[2 * (m.first / p.streamBlockMax) + q];
    n          = (n - q * p.streamBlockMax < p.streamBlockMax) ? n - q * p.streamBlockMax : p.streamBlockMax;
    m.first   += q * p.streamBlockMax;
    m.outByte += part.relByte;
    m.subFirst += part.relSub;
    m.nSub     = part.nSub;
  }
  const u64  i       = g - m.first;
  const BlockShape s = block_shape(n, wSuffix);

  const u64 *o    = (const u64 *)(out + (m.outByte - c.chunkByte0));
  const u64 *bgn  = o + 2;
  u64       *data = (u64 *)(o + 2 + 2 * (u64)m.nSub);
  const u32 *sw   = subWord + (m.subFirst - c.chunkSub0);

  const u64 hprev = (i == 0) ? 0 : key_hi_part(keys[g - 1], s.b, s.u);
  const u64 h     = key_hi_part(key, s.b, s.u);
  const u64 Ui    = HDR_BITS + hprev + i * (1 + (u64)s.b);
  const u64 c1    = HDR_BITS + h + i * (1 + (u64)s.b) + 1;
  const u64 c2    = c1 + s.ls;

  //  sub-block of the unary field: largest j with bgn[j] <= Ui
  u32 lo = 0, hi = m.nSub;
  while (hi - lo > 1) {
    const u32 mid = (lo + hi) >> 1;
    if (bgn[mid] <= Ui) lo = mid; else hi = mid;
  }
  u32 j = lo;
  or_bits(data, 64 * (u64)sw[j] + (Ui - bgn[j]) + (h - hprev), 1, 1);          //  d zeros then the one

  if (s.ls) {
    if (j + 1 < m.nSub && bgn[j + 1] <= c1) j++;
    or_bits(data, 64 * (u64)sw[j] + (c1 - bgn[j]), s.ls, key_hi(key));
  }
  if (s.rs) {
    if (j + 1 < m.nSub && bgn[j + 1] <= c2) j++;
    or_bits(data, 64 * (u64)sw[j] + (c2 - bgn[j]), s.rs, key_lo(key));
  }

  //  the count
  const u64 uN = ustart(keys, m.first, n, s);
  const u64 V  = uN + 32 * i;
  lo = j; hi = m.nSub;                                                           //  counts come after every k-mer field
  while (hi - lo > 1) {
    const u32 mid = (lo + hi) >> 1;
    if (bgn[mid] <= V) lo = mid; else hi = mid;
  }
  const u32 cnt = p.counts ? p.counts[g] : (u32)(p.runStart[g + 1] - p.runStart[g]);
  or_bits(data, 64 * (u64)sw[lo] + (V - bgn[lo]), 32, cnt);
}

}  //  namespace


size_t mb_encode_scan_tmp_bytes(u64 nBlocks) {
  const u64 nTiles = (nBlocks + SCAN_TILE - 1) / SCAN_TILE;
  return (size_t)(2 * nTiles + 2) * sizeof(u64);
}

cudaError_t mb_launch_encode_layout(const MbEncodeParams &p, MbBlockMeta *dMeta, u64 *dTotals, void *dScanTmp, cudaStream_t s) {
  if (p.nBlocks == 0) return cudaMemsetAsync(dTotals, 0, 4 * sizeof(u64), s);
  const u64 g1 = (p.nBlocks + 127) / 128;
  if (p.keyWords == 1) layout_kernel<1><<<(unsigned)g1, 128, 0, s>>>(p, dMeta, dTotals);
  else                 layout_kernel<2><<<(unsigned)g1, 128, 0, s>>>(p, dMeta, dTotals);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;

  const u64 nTiles = (p.nBlocks + SCAN_TILE - 1) / SCAN_TILE;
  u64 *tileSums = (u64 *)dScanTmp;
  scan_tile_sums_kernel<<<(unsigned)nTiles, SCAN_THREADS, 0, s>>>(dMeta, p.nBlocks, tileSums);
  scan_tile_offsets_kernel<<<1, SCAN_THREADS, 0, s>>>(tileSums, nTiles, dTotals);
  scan_apply_kernel<<<(unsigned)nTiles, SCAN_THREADS, 0, s>>>(dMeta, p.nBlocks, tileSums);
  return cudaGetLastError();
}

cudaError_t mb_launch_encode_write(const MbEncodeParams &p, const MbEncodeChunk &c, const MbBlockMeta *dMeta, const u64 *dTotals,
                                   u32 *dSubWord, uint8_t *dOut, u64 *dIndex, const u64 *dFileBase, cudaStream_t s) {
  if (c.chunkCount == 0) return cudaSuccess;
  const u64 g1 = (c.chunkCount + 127) / 128;
  if (p.keyWords == 1) header_kernel<1><<<(unsigned)g1, 128, 0, s>>>(p, c, dMeta, dTotals, dSubWord, dOut, dIndex, dFileBase);
  else                 header_kernel<2><<<(unsigned)g1, 128, 0, s>>>(p, c, dMeta, dTotals, dSubWord, dOut, dIndex, dFileBase);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || p.nDistinct == 0) return e;
  //  one thread per distinct k-mer; the chunk's exact range is read from the block table on the device,
  //  the grid is sized by the pass total (threads past the chunk's end return at once)
  const u64 g2 = (p.nDistinct + 255) / 256;
  if (p.keyWords == 1) fields_kernel<1><<<(unsigned)g2, 256, 0, s>>>(p, c, dMeta, dTotals, dSubWord, dOut);
  else                 fields_kernel<2><<<(unsigned)g2, 256, 0, s>>>(p, c, dMeta, dTotals, dSubWord, dOut);
  return cudaGetLastError();
}
