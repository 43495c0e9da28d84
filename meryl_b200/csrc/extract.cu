//  K2 / K3: canonical k-mer extraction from the packed read store, fused with the prefix scatter.
//
//  Replaces the reference's hot loop #1:
//    kmerIterator::nextMer        src/utility/src/kmers-v1/kmers-iterator.H:63-87   (k valid bases in a row)
//    kmerTiny::addR / addL        src/utility/src/kmers-v1/kmers-tiny.H:86-87       (forward / reverse-complement)
//    insertKmers                  src/meryl/merylOp-countThreads.C:236-280          (canonical = min(f, r); prefix | suffix)
//    merylCountArray::add         src/meryl/merylCountArray.C:491-728               (append to the prefix bucket)
//
//  B200 formulation.  There is no rolling state: position p of the store starts a k-mer iff the k
//  validity bits from p are all set, and its forward value is a 2k-bit window of the packed
//  stream; the reverse complement is a bit reversal of that window.  A persistent CTA walks tiles
//  of 32*THREADS positions; each tile's store records (12 B per 32 positions, plus a 4-record
//  halo for the k-1 overhang) are brought into shared memory by one TMA bulk copy
//  (cp.async.bulk + mbarrier), double buffered so the next tile streams in under the current
//  tile's integer work.  Per thread: 32 consecutive positions.
//    1. validity of all 32 windows at once by log-step erosion of the validity bits;
//    2. all lanes walk their 32 positions in lock step, rolling forward and reverse-complement
//       k-mers (every shift is warp-uniform, no divergence, ~16 integer ops per position); the
//       canonical k-mer's top bits are its prefix bin;
//    3. k-mers of this pass's bins take their rank within (tile, bin) by a shared-memory atomic;
//       one global atomic per (tile, bin) reserves the run in the bin's segment;
//    4. the kept k-mers are produced again (rolling when the pass keeps most positions, direct
//       window extraction when it keeps few), placed bin-sorted in a 32 KB shared-memory stage
//       buffer and written out as coalesced per-bin runs.
//  The same kernel in COUNT_ONLY form is the histogram pre-pass (K2a) that sizes the segments.
//
//  Work: integer/bit ops, ~16-20 per position per phase; HBM traffic 0.375 B per position read +
//  8 (16) B per kept k-mer written -- integer-ALU bound, not HBM bound.

#include "kernels.cuh"

namespace {

//  ---- TMA bulk copy + mbarrier (sm_90+/sm_100a PTX) ------------------------------------------

#ifdef MB_CUSIM
//  Host emulator used by the CPU tests (tests/cusim): the barrier word holds the phase, the bulk copy is a
//  memcpy that completes at once.
__device__ __forceinline__ void mbar_init(u64 *bar, u32)            { *bar = 0; }
__device__ __forceinline__ void mbar_fence_init(void)               { }
__device__ __forceinline__ void mbar_expect_tx(u64 *, u32)          { }
__device__ __forceinline__ bool mbar_try_wait(u64 *bar, u32 parity) { if ((*bar & 1) != parity) return true; cusim::yield(); return false; }
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, u32 bytes, u64 *bar) { memcpy(dst, src, bytes); *bar ^= 1; cusim::g_progress++; }
__device__ __forceinline__ void fence_proxy_async(void)             { }
#else
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init(void) {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(u64 *bar, u32 parity) {
  u32 ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, u32 bytes, u64 *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async(void) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
#endif

//  ---- bit helpers ---------------------------------------------------------------------------

//  The 192-bit window of a thread: positions 0..95 relative to its first position, 2 bits each,
//  position j in bits [191-2j, 190-2j] of (w0:w1:w2).
struct Window { u64 w0, w1, w2; };

//  nbits (1..64) bits of the window ending just before bit offset `end` (nbits <= end <= 192).
__device__ __forceinline__ u64 bits_ending_at(const Window &w, u32 end, u32 nbits) {
  u64 x;
  if (end <= 64)        x = w.w0 >> (64 - end);
  else if (end < 128)   x = (w.w0 << (end - 64)) | (w.w1 >> (128 - end));
  else if (end == 128)  x = w.w1;
  else if (end < 192)   x = (w.w1 << (end - 128)) | (w.w2 >> (192 - end));
  else                  x = w.w2;
  return nbits >= 64 ? x : (x & (((u64)1 << nbits) - 1));
}

//  Reverse complement of an n-base value held in the low 2n bits (n <= 32).
__device__ __forceinline__ u64 revcomp64(u64 x, u32 n) {
  u64 y = __brevll(x) >> (64 - 2 * n);
  y = ((y >> 1) & 0x5555555555555555ull) | ((y & 0x5555555555555555ull) << 1);
  const u64 m = (n == 32) ? ~0ull : (((u64)1 << (2 * n)) - 1);
  return y ^ (0xaaaaaaaaaaaaaaaaull & m);
}

//  128-bit variant for 33 <= n <= 64: value (hi:lo) in the low 2n bits.
__device__ __forceinline__ void revcomp128(u64 hi, u64 lo, u32 n, u64 &rhi, u64 &rlo) {
  u64 a = __brevll(lo), b = __brevll(hi);            //  reversed 128 bits = (a:b)
  a = ((a >> 1) & 0x5555555555555555ull) | ((a & 0x5555555555555555ull) << 1);
  b = ((b >> 1) & 0x5555555555555555ull) | ((b & 0x5555555555555555ull) << 1);
  const u32 sh = 128 - 2 * n;                        //  0..62
  if (sh == 0) { rhi = a; rlo = b; }
  else         { rhi = a >> sh;  rlo = (b >> sh) | (a << (64 - sh)); }
  const u64 mh = (n == 64) ? ~0ull : (((u64)1 << (2 * n - 64)) - 1);
  rhi ^= 0xaaaaaaaaaaaaaaaaull & mh;
  rlo ^= 0xaaaaaaaaaaaaaaaaull;
}

//  ---- bit-sliced file-range filter ------------------------------------------------------------------
//  A pass keeps the k-mers of a contiguous range of the 64 file prefixes (top 6 bits = first 3 bases of the
//  canonical k-mer).  With F3(p) = value of bases p..p+2 and R3(p) = reverse-complement value of bases
//  p+k-3..p+k-1, the canonical k-mer's file prefix is exactly min(F3, R3) (if the two differ the smaller one
//  decides which strand is canonical; if they are equal both strands start alike).  So "is position p in
//  this pass" is a pair of 6-bit comparisons -- evaluated here for all 32 positions of a thread at once,
//  bit-sliced, instead of rolling two 2k-bit k-mers through every position only to drop most of them.
//
//  Slot word: slot p (0..31) = bits [63-2p, 62-2p]; a boolean of position p lives in the slot's top bit.

//  Slot p = 2-bit code of thread-relative position p + o  (o < 96; positions beyond the window read as 0).
__device__ __forceinline__ u64 win_slots(const Window &w, u32 o) {
  const u32 s = 2 * (o & 31);
  u64 a, b;
  if (o < 32)      { a = w.w0; b = w.w1; }
  else if (o < 64) { a = w.w1; b = w.w2; }
  else             { a = w.w2; b = 0;    }
  return s ? ((a << s) | (b >> (64 - s))) : a;
}

//  Running comparison "value >= bound" over bit planes fed most significant first (planes live in the slots' top
//  bits).  Branch-free: m = all ones where the bound has a 1 in this plane.
//    bound bit 1:  still-equal positions need a 1;             bound bit 0:  a 1 makes an equal position greater.
struct GeChain {
  u64 gt, eq;
  __device__ __forceinline__ void init(void)            { gt = 0; eq = ~0ull; }
  __device__ __forceinline__ void step(u64 x, u64 m)    { gt |= eq & x & ~m;  eq &= ~(x ^ m); }
  __device__ __forceinline__ u64  ge(void) const        { return gt | eq; }
};

//  Slot mask (top bit of each slot) of the positions whose k-mer has its file prefix in [fLo, fHi), 0 <= fLo < fHi <= 64.
//  Planes are produced and consumed one base at a time (few live registers).
__device__ __forceinline__ u64 file_range_slots(const Window &w, u32 k, u32 mode, u32 fLo, u32 fHi) {
  GeChain fl, fh, rl, rh;
  fl.init();  fh.init();  rl.init();  rh.init();
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const u64 lo1 = 0ull - (u64)((fLo >> (5 - 2 * i)) & 1), lo0 = 0ull - (u64)((fLo >> (4 - 2 * i)) & 1);
    const u64 hi1 = 0ull - (u64)((fHi >> (5 - 2 * i)) & 1), hi0 = 0ull - (u64)((fHi >> (4 - 2 * i)) & 1);
    if (mode != MB_MODE_REVERSE) {
      const u64 c = win_slots(w, i);                   //  base p + i: high bit in place, low bit one to the right
      fl.step(c, lo1);  fl.step(c << 1, lo0);
      fh.step(c, hi1);  fh.step(c << 1, hi0);
    }
    if (mode != MB_MODE_FORWARD) {
      const u64 d = win_slots(w, k - 1 - i);           //  base p + k-1-i, complemented (code ^ 2: the high bit flips)
      rl.step(~d, lo1);  rl.step(d << 1, lo0);
      rh.step(~d, hi1);  rh.step(d << 1, hi0);
    }
  }
  const u64 fGeLo = fl.ge(), rGeLo = rl.ge();
  const u64 fGeHi = fHi >= 64 ? 0 : fh.ge(), rGeHi = fHi >= 64 ? 0 : rh.ge();
  u64 m;
  if      (mode == MB_MODE_FORWARD) m = fGeLo & ~fGeHi;
  else if (mode == MB_MODE_REVERSE) m = rGeLo & ~rGeHi;
  else                              m = fGeLo & rGeLo & ~(fGeHi & rGeHi);   //  lo <= min(F, R) < hi
  return m & 0xaaaaaaaaaaaaaaaaull;
}

//  10 bits (5 bases) starting at slot q (0..31) of the 96 bits (v0:v1:v2).
__device__ __forceinline__ u32 five_bases_at(u32 v0, u32 v1, u32 v2, u32 q) {
  const u32 s = 2 * q;
  const u32 x = (s < 32) ? __funnelshift_l(v1, v0, s) : __funnelshift_l(v2, v1, s - 32);
  return x >> 22;
}

//  Reverse complement of a 5-base (10-bit) value.
__device__ __forceinline__ u32 revcomp5(u32 v) {
  u32 r = __brev(v) >> 22;
  r = ((r >> 1) & 0x155u) | ((r & 0x155u) << 1);
  return r ^ 0x2aau;
}

template <int KW> struct ExtractShape;
template <> struct ExtractShape<1> { static const int THREADS = 256; };
template <> struct ExtractShape<2> { static const int THREADS = 128; };

template <int KW, bool S32>
struct ExtractSmem {
  typedef typename KeyT<S32 ? 0 : KW>::type Key;             //  what is stored: full key, or its low 32 bits
  static const int    THREADS   = ExtractShape<KW>::THREADS;
  static const int    TILE_POS  = THREADS * 32;
  static const int    REC_U32   = (THREADS + 4) * MB_STORE_REC_WORDS;       //  records incl. halo
  static const int    REC_BYTES = REC_U32 * 4;                              //  multiple of 16
  static const size_t OFF_BAR   = 0;
  static const size_t OFF_REC0  = 128;
  static const size_t OFF_REC1  = OFF_REC0 + ((REC_BYTES + 127) / 128) * 128;
  static const size_t OFF_TABLES= OFF_REC1 + ((REC_BYTES + 127) / 128) * 128;
  //  After OFF_TABLES the layout depends on the pass (nT = bins of the pass, cap = staged keys per tile):
  //    u32 cnt[nT] | u32 binOff[nT] | u64 gofs[nT] | u16 rank[TILE_POS] | Key stage[cap] | u16 sbin[cap] (suffix keys only)
  //  A pass over a fifth of the bins needs ~45 KB instead of 86 KB: 4 CTAs per SM instead of 2, which the
  //  latency-bound rolling loop of phase A needs.
  __host__ __device__ static size_t tablesBytes(u32 nT)            { return ((size_t)nT * 16 + 15) / 16 * 16; }
  __host__ __device__ static size_t scatterBytes(u32 nT, u32 cap)  { return OFF_TABLES + tablesBytes(nT) + (size_t)TILE_POS * 2 + (size_t)cap * sizeof(Key) + (S32 ? (size_t)cap * 2 : 0) + 16; }
  __host__ __device__ static size_t countBytes(u32 nT)             { return OFF_TABLES + (size_t)nT * 4; }
};

//  Validity of the 32 windows of length k that start at this thread's positions.
//  m0..m2: validity bits of positions 0..95 (bit 31-j of m_i is position 32i+j).
__device__ __forceinline__ u32 windows_valid(u32 m0, u32 m1, u32 m2, u32 k) {
  u64 hi = ((u64)m0 << 32) | m1, lo = (u64)m2 << 32;                        //  position q at bit 127-q
  u32 run = 1;
  while (run < k) {
    const u32 s = min(run, k - run);                                        //  1..32
    const u64 shi = (hi << s) | (lo >> (64 - s)), slo = lo << s;
    hi &= shi;  lo &= slo;
    run += s;
  }
  return (u32)(hi >> 32);
}

__device__ __forceinline__ u64 kmer_key(const Window &w, u32 q, u32 k, u32 mode, u64 *) {
  const u64 f = bits_ending_at(w, 2 * (q + k), 2 * k);
  if (mode == MB_MODE_FORWARD) return f;
  const u64 r = revcomp64(f, k);
  if (mode == MB_MODE_REVERSE) return r;
  return f < r ? f : r;
}

__device__ __forceinline__ ulonglong2 kmer_key(const Window &w, u32 q, u32 k, u32 mode, ulonglong2 *) {
  const u64 flo = bits_ending_at(w, 2 * (q + k), 64);
  const u64 fhi = bits_ending_at(w, 2 * (q + k) - 64, 2 * k - 64);
  u64 rhi, rlo;
  revcomp128(fhi, flo, k, rhi, rlo);
  bool useF = (mode == MB_MODE_FORWARD);
  if (mode == MB_MODE_CANONICAL) useF = (fhi < rhi) || (fhi == rhi && flo < rlo);
  return useF ? make_ulonglong2(flo, fhi) : make_ulonglong2(rlo, rhi);
}

//  ---- lock-step rolling k-mers ------------------------------------------------------------------
//  All lanes walk their 32 positions in the same order, so every shift amount is warp-uniform
//  and there is no divergence: ~16 integer ops per position for forward + reverse complement.

template <int KW> struct Roll;

template <> struct Roll<1> {
  u64 f, r;
  __device__ __forceinline__ void init(const Window &w, u32 k) {
    f = bits_ending_at(w, 2 * k, 2 * k);
    r = revcomp64(f, k);
  }
  __device__ __forceinline__ void push(u32 code, u32 k, u64 mask) {
    f = ((f << 2) | code) & mask;
    r = (r >> 2) | ((u64)(code ^ 2) << (2 * k - 2));
  }
  __device__ __forceinline__ u64 key(u32 mode) const {
    return (mode == MB_MODE_FORWARD) ? f : (mode == MB_MODE_REVERSE) ? r : (f < r ? f : r);
  }
};

template <> struct Roll<2> {
  u64 fl, fh, rl, rh;
  __device__ __forceinline__ void init(const Window &w, u32 k) {
    fl = bits_ending_at(w, 2 * k, 64);
    fh = bits_ending_at(w, 2 * k - 64, 2 * k - 64);
    revcomp128(fh, fl, k, rh, rl);
  }
  __device__ __forceinline__ void push(u32 code, u32 k, u64 maskHi) {
    fh = ((fh << 2) | (fl >> 62)) & maskHi;
    fl = (fl << 2) | code;
    rl = (rl >> 2) | (rh << 62);
    rh = (rh >> 2) | ((u64)(code ^ 2) << (2 * k - 66));
  }
  __device__ __forceinline__ ulonglong2 key(u32 mode) const {
    bool useF = (mode == MB_MODE_FORWARD);
    if (mode == MB_MODE_CANONICAL) useF = (fh < rh) || (fh == rh && fl < rl);
    return useF ? make_ulonglong2(fl, fh) : make_ulonglong2(rl, rh);
  }
};

//  2-bit code of thread-relative position pos (0..95); pos is warp-uniform.
__device__ __forceinline__ u32 window_code(const Window &w, u32 pos) {
  const u64 word = pos < 32 ? w.w0 : pos < 64 ? w.w1 : w.w2;
  return (u32)(word >> (62 - 2 * (pos & 31))) & 3;
}

//  Stored form of a key: itself, or (suffix32) its low 32 bits.
__device__ __forceinline__ void to_stored(u64 k, u32 &o)                         { o = (u32)k; }
__device__ __forceinline__ void to_stored(u64 k, u64 &o)                         { o = k; }
__device__ __forceinline__ void to_stored(const ulonglong2 &k, ulonglong2 &o)    { o = k; }

template <int KW, bool COUNT_ONLY, bool S32>
__global__ void __launch_bounds__(ExtractShape<KW>::THREADS, 4)
extract_kernel(MbExtractParams p, typename KeyT<S32 ? 0 : KW>::type *__restrict__ out,
               const u64 *__restrict__ binStart, u64 *__restrict__ binCursor, u64 *__restrict__ binCount, u64 nTiles, u32 stageCap) {
  typedef typename KeyT<KW>::type           Key;             //  what is computed
  typedef typename KeyT<S32 ? 0 : KW>::type Stored;          //  what is written
  typedef ExtractSmem<KW, S32>              S;
  extern __shared__ __align__(128) unsigned char smem[];

  u64            *bar   = (u64 *)(smem + S::OFF_BAR);
  u32            *rec[2] = { (u32 *)(smem + S::OFF_REC0), (u32 *)(smem + S::OFF_REC1) };
  //  tables are indexed by (bin - binLo); COUNT_ONLY uses only cnt[], over all bins (binLo == 0)
  const u32       nT     = p.binHi - p.binLo;
  u32            *cnt    = (u32 *)(smem + S::OFF_TABLES);
  u32            *binOff = cnt + nT;
  u64            *gofs   = (u64 *)(smem + S::OFF_TABLES + (((size_t)nT * 8 + 7) / 8) * 8);
  unsigned short *rank   = (unsigned short *)(smem + S::OFF_TABLES + S::tablesBytes(nT));
  Stored         *stage  = (Stored *)(((uintptr_t)(rank + S::TILE_POS) + 15) & ~(uintptr_t)15);
  unsigned short *sbin   = (unsigned short *)(stage + stageCap);
  __shared__ u32 sWarpSum[ExtractShape<KW>::THREADS / 32];

  const u32 t = threadIdx.x;
  const u32 nBins = 1u << p.binBits;
  //  Does this pass keep >= 1/3 of the k-mers?  The caller knows (canonical k-mers are heavily skewed towards the
  //  low bins, so the share of bins says little); without a hint, judge by the bins.
  const bool rollAgain = p.keptPerMille ? (p.keptPerMille >= 333) : ((p.binHi - p.binLo) * 3 >= nBins);
  //  file-aligned minority pass: bit-sliced membership test instead of rolling through every position
  const u32  perFile   = p.binBits >= 6 ? (1u << (p.binBits - 6)) : 0;
  const bool prefilter = !rollAgain && p.k >= 5 && p.binBits >= 6 && p.binBits <= 10 && !p.noPrefilter &&
                         (p.binLo & (perFile - 1)) == 0 && (p.binHi & (perFile - 1)) == 0;
  const u32  fLo = perFile ? p.binLo / perFile : 0, fHi = perFile ? p.binHi / perFile : 0;
#ifdef MB_CUSIM
  if (threadIdx.x == 0 && blockIdx.x == 0 && !COUNT_ONLY && getenv("CUSIM_TRACE"))
    fprintf(stderr, "cusim: extract bins [%u,%u) of %u: %s\n", p.binLo, p.binHi, nBins, prefilter ? "bit-sliced prefilter" : rollAgain ? "rolling" : "rolling + direct");
#endif

  if (t == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
  for (u32 b = t; b < nT; b += blockDim.x) cnt[b] = 0;
  __syncthreads();

  const unsigned char *storeBytes = (const unsigned char *)p.store;

  u64 tile = blockIdx.x;
  if (t == 0 && tile < nTiles) {
    mbar_expect_tx(&bar[0], S::REC_BYTES);
    tma_bulk_g2s(rec[0], storeBytes + tile * (u64)(S::THREADS * MB_STORE_REC_WORDS * 4), S::REC_BYTES, &bar[0]);
  }


  for (u32 it = 0; tile < nTiles; it++, tile += gridDim.x) {
    const u32 buf = it & 1;

    //  Prefetch the next tile into the other buffer (it was last read two iterations ago, and
    //  every thread has passed a __syncthreads since).
    const u64 next = tile + gridDim.x;
    if (t == 0 && next < nTiles) {
      fence_proxy_async();
      mbar_expect_tx(&bar[buf ^ 1], S::REC_BYTES);
      tma_bulk_g2s(rec[buf ^ 1], storeBytes + next * (u64)(S::THREADS * MB_STORE_REC_WORDS * 4), S::REC_BYTES, &bar[buf ^ 1]);
    }

    while (!mbar_try_wait(&bar[buf], (it >> 1) & 1)) { }

    const u32 *r = rec[buf] + t * MB_STORE_REC_WORDS;
    Window w;
    w.w0 = ((u64)r[0] << 32) | r[1];
    w.w1 = ((u64)r[3] << 32) | r[4];
    w.w2 = ((u64)r[6] << 32) | r[7];
    u32 valid = windows_valid(r[2], r[5], r[8], p.k);

    //  Positions at or beyond the end of the store never start a k-mer (their records are zero),
    //  so no bound check is needed here.

    //  Phase A, lock-step over my 32 positions: canonical k-mer by rolling, its bin, and for the
    //  bins of this pass a rank inside (tile, bin).
    const u32 binShift = 2 * p.k - p.binBits;
    const u64 rollMask = (KW == 1) ? ((p.k == 32) ? ~0ull : (((u64)1 << (2 * p.k)) - 1))
                                   : ((p.k == 64) ? ~0ull : (((u64)1 << (2 * p.k - 64)) - 1));
    u32 keep = 0;
    if (COUNT_ONLY && p.k >= 5 && p.binBits <= 10 && !p.noPrefilter) {
      //  Histogram pre-pass: the bin is the top of min(F5, R5) -- 5 bases of each strand, no 2k-bit k-mers.
      //  z = reverse complement of the 36 bases from position k-5 on, so that R5 of position q is the plain
      //  10-bit field at slot 31-q of z; with q a compile-time constant every field is one funnel shift.
      if (valid) {
        const u64 ya = win_slots(w, p.k - 5), yb = win_slots(w, p.k - 5 + 32);
        const u64 rcA = revcomp64(ya, 32);
        const u32 rcB = (u32)revcomp64(yb >> 56, 4);
        const u32 z0 = (rcB << 24) | (u32)(rcA >> 40), z1 = (u32)(rcA >> 8), z2 = (u32)rcA << 24;
        const u32 f0 = (u32)(w.w0 >> 32), f1 = (u32)w.w0, f2 = (u32)(w.w1 >> 32);
        const u32 sh = 10 - p.binBits;
#pragma unroll
        for (u32 q = 0; q < 32; q++) {
          if ((valid << q) & 0x80000000u) {
            const u32 f5 = five_bases_at(f0, f1, f2, q);
            const u32 r5 = five_bases_at(z0, z1, z2, 31 - q);
            const u32 top10 = (p.mode == MB_MODE_FORWARD) ? f5 : (p.mode == MB_MODE_REVERSE) ? r5 : min(f5, r5);
            atomicAdd(&cnt[top10 >> sh], 1u);
          }
        }
      }
    } else if (!COUNT_ONLY && prefilter) {
      //  Phase A for a file-aligned pass that keeps a minority of the bins: bit-sliced membership test of all
      //  32 positions, then only the members get their exact bin (5 bases of each strand) and a rank.
      u64 cand = valid ? file_range_slots(w, p.k, p.mode, fLo, fHi) : 0;
      if (cand) {
        const u64 ya = win_slots(w, p.k - 5), yb = win_slots(w, p.k - 5 + 32);
        const u32 f0 = (u32)(w.w0 >> 32), f1 = (u32)w.w0, f2 = (u32)(w.w1 >> 32);
        const u32 r0 = (u32)(ya >> 32),   r1 = (u32)ya,   r2 = (u32)(yb >> 32);
        while (cand) {
          const u32 b = __clzll((long long)cand);      //  top set bit = slot q's top bit: 63 - b = 63 - 2q
          cand &= ~(0x8000000000000000ull >> b);
          const u32 q = b >> 1;
          if ((valid << q) & 0x80000000u) {
            const u32 f5 = five_bases_at(f0, f1, f2, q);
            const u32 r5 = revcomp5(five_bases_at(r0, r1, r2, q));
            const u32 top10 = (p.mode == MB_MODE_FORWARD) ? f5 : (p.mode == MB_MODE_REVERSE) ? r5 : min(f5, r5);
            const u32 bin = top10 >> (10 - p.binBits);
            keep |= 0x80000000u >> q;
            rank[t * 32 + q] = (unsigned short)atomicAdd(&cnt[bin - p.binLo], 1u);
          }
        }
      }
    } else if (valid) {
      Roll<KW> roll;
      roll.init(w, p.k);
      for (u32 q = 0; q < 32; q++) {
        if ((valid << q) & 0x80000000u) {
          const u32 bin = key_field(roll.key(p.mode), binShift, p.binBits);
          if (COUNT_ONLY) {
            atomicAdd(&cnt[bin - p.binLo], 1u);
          } else if (bin >= p.binLo && bin < p.binHi) {
            keep |= 0x80000000u >> q;
            rank[t * 32 + q] = (unsigned short)atomicAdd(&cnt[bin - p.binLo], 1u);
          }
        }
        roll.push(window_code(w, q + p.k), p.k, rollMask);
      }
    }
    if (!COUNT_ONLY) {
      __syncthreads();

      //  Exclusive scan of the tile's bin counts (-> place of every bin's run in the stage buffer)
      //  and, in the same sweep, one global atomic per non-empty bin to reserve its run in HBM.
      const u32 nPassBins = p.binHi - p.binLo;
      const u32 perThread = (nPassBins + blockDim.x - 1) / blockDim.x;
      const u32 b0 = p.binLo + t * perThread;
      u32 mySum = 0;
      for (u32 i = 0; i < perThread; i++) { const u32 b = b0 + i; if (b < p.binHi) mySum += cnt[b - p.binLo]; }
      u32 incl = mySum;
      for (int o = 1; o < 32; o <<= 1) {
        const u32 v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((t & 31) >= (u32)o) incl += v;
      }
      if ((t & 31) == 31) sWarpSum[t >> 5] = incl;
      __syncthreads();
      u32 run = incl - mySum, total = 0;
      for (u32 wi = 0; wi < blockDim.x / 32; wi++) { const u32 ws = sWarpSum[wi]; if (wi < (t >> 5)) run += ws; total += ws; }
      for (u32 i = 0; i < perThread; i++) {
        const u32 b = b0 + i;
        if (b < p.binHi) {
          const u32 c = cnt[b - p.binLo];
          binOff[b - p.binLo] = run;
          if (c) {
            gofs[b - p.binLo] = binStart[b] + atomicAdd((unsigned long long *)&binCursor[b], (unsigned long long)c) - run;
            cnt[b - p.binLo]  = 0;
          }
          run += c;
        }
      }
      __syncthreads();

      //  Phase B: the kept k-mers are produced again (rolling when the pass keeps most positions,
      //  direct window extraction when it keeps few) and placed bin-sorted in the stage buffer;
      //  what does not fit goes straight to its reserved slot.
      if (keep) {
        if (rollAgain) {
          Roll<KW> roll;
          roll.init(w, p.k);
          for (u32 q = 0; q < 32; q++) {
            if ((keep << q) & 0x80000000u) {
              const Key key = roll.key(p.mode);
              const u32 bin = key_field(key, binShift, p.binBits) - p.binLo;
              const u32 pos = binOff[bin] + rank[t * 32 + q];
              Stored sk;  to_stored(key, sk);
              if (pos < stageCap) { stage[pos] = sk; if (S32) sbin[pos] = (unsigned short)bin; } else out[gofs[bin] + pos] = sk;
            }
            roll.push(window_code(w, q + p.k), p.k, rollMask);
          }
        } else {
          for (u32 v = keep; v; ) {
            const u32 q = __clz(v);
            v &= ~(0x80000000u >> q);
            const Key key = kmer_key(w, q, p.k, p.mode, (Key *)0);
            const u32 bin = key_field(key, binShift, p.binBits) - p.binLo;
            const u32 pos = binOff[bin] + rank[t * 32 + q];
            Stored sk;  to_stored(key, sk);
            if (pos < stageCap) { stage[pos] = sk; if (S32) sbin[pos] = (unsigned short)bin; } else out[gofs[bin] + pos] = sk;
          }
        }
      }
      __syncthreads();
      //  Copy-out: consecutive threads, consecutive staged keys -> every bin's run is written coalesced.
      const u32 nStaged = total < stageCap ? total : stageCap;
      for (u32 j = t; j < nStaged; j += blockDim.x) {
        const Stored sk = stage[j];
        const u32 bin = S32 ? (u32)sbin[j] : key_field(sk, binShift, p.binBits) - p.binLo;
        out[gofs[bin] + j] = sk;
      }
    }
    __syncthreads();                                  //  all reads of rec[buf] / gbase done
  }

  if (COUNT_ONLY) {
    __syncthreads();
    for (u32 b = t; b < nBins; b += blockDim.x)
      if (cnt[b]) atomicAdd((unsigned long long *)&binCount[b], (unsigned long long)cnt[b]);   //  COUNT_ONLY: binLo == 0
  }
}

//  ---- dense form of a filtered pass (k <= 32) ----------------------------------------------------------------------
//
//  extract_kernel's filtered pass walks a thread's members in a divergent loop (a thread has 0..32 members among its 32
//  positions, the warp runs as long as its busiest lane: about half the lanes idle) and does everything per member twice
//  (rank in phase A, k-mer in phase B).  Here the bit-sliced membership test stays per thread, but the members of the
//  whole tile are COMPACTED into a list of positions first (one block scan of the per-thread member counts), and then
//  consecutive lanes take consecutive members: every member gets its k-mer by direct window extraction from the
//  shared-memory records ONCE, its rank inside (tile, bin) from one returning shared-memory atomic issued by full warps, and
//  keeps (stored key, bin, rank) in registers over the bin scan; placement in the bin-sorted stage buffer and the coalesced
//  copy-out are as before.  ITEMS = members per thread and round (CAP = 256 ITEMS members per round; a tile with more
//  members takes another round -- rare, the host sizes ITEMS for 1.3x the expected share).
//  Per tile of 8192 positions at one fifth kept: ~830 warp instructions per thread row instead of ~1350.
template <bool S32, int ITEMS>
__global__ void __launch_bounds__(256, ITEMS <= 12 ? 4 : ITEMS <= 16 ? 3 : 2)
extract_dense_kernel(MbExtractParams p, typename KeyT<S32 ? 0 : 1>::type *__restrict__ out,
                     const u64 *__restrict__ binStart, u64 *__restrict__ binCursor, u64 nTiles) {
  typedef typename KeyT<S32 ? 0 : 1>::type Stored;
  typedef ExtractSmem<1, S32>              S;
  const int THREADS = 256, CAP = THREADS * ITEMS;
  extern __shared__ __align__(128) unsigned char smem[];

  u64            *bar    = (u64 *)(smem + S::OFF_BAR);
  u32            *rec[2] = { (u32 *)(smem + S::OFF_REC0), (u32 *)(smem + S::OFF_REC1) };
  const u32       nT     = p.binHi - p.binLo;
  u32            *cnt    = (u32 *)(smem + S::OFF_TABLES);
  u32            *binOff = cnt + nT;
  u64            *gofs   = (u64 *)(smem + S::OFF_TABLES + (((size_t)nT * 8 + 7) / 8) * 8);
  unsigned short *plist  = (unsigned short *)(smem + S::OFF_TABLES + S::tablesBytes(nT));   //  CAP member positions of this round ...
  unsigned short *sbin   = plist;                                                          //  ... later the staged keys' bins (S32)
  Stored         *stage  = (Stored *)(((uintptr_t)(plist + CAP) + 15) & ~(uintptr_t)15);
  __shared__ u32 sWarpSum[THREADS / 32];

  const u32 t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const u32 perFile  = 1u << (p.binBits - 6);
  const u32 fLo = p.binLo / perFile, fHi = p.binHi / perFile;
  const u32 binShift = 2 * p.k - p.binBits;

#ifdef MB_CUSIM
  if (threadIdx.x == 0 && blockIdx.x == 0 && getenv("CUSIM_TRACE"))
    fprintf(stderr, "cusim: extract bins [%u,%u) of %u: dense, %d members per thread and round\n", p.binLo, p.binHi, 1u << p.binBits, ITEMS);
#endif
  if (t == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
  for (u32 b = t; b < nT; b += THREADS) cnt[b] = 0;
  __syncthreads();

  const unsigned char *storeBytes = (const unsigned char *)p.store;
  u64 tile = blockIdx.x;
  if (t == 0 && tile < nTiles) {
    mbar_expect_tx(&bar[0], S::REC_BYTES);
    tma_bulk_g2s(rec[0], storeBytes + tile * (u64)(THREADS * MB_STORE_REC_WORDS * 4), S::REC_BYTES, &bar[0]);
  }

  for (u32 it = 0; tile < nTiles; it++, tile += gridDim.x) {
    const u32 buf = it & 1;
    const u64 next = tile + gridDim.x;
    if (t == 0 && next < nTiles) {                      //  rec[buf ^ 1] was last read before the barrier that ended the previous tile
      fence_proxy_async();
      mbar_expect_tx(&bar[buf ^ 1], S::REC_BYTES);
      tma_bulk_g2s(rec[buf ^ 1], storeBytes + next * (u64)(THREADS * MB_STORE_REC_WORDS * 4), S::REC_BYTES, &bar[buf ^ 1]);
    }
    while (!mbar_try_wait(&bar[buf], (it >> 1) & 1)) { }

    //  members of this pass among my 32 positions, as a slot mask (slot q's top bit = bit 63 - 2q)
    u64 cand = 0;
    {
      const u32 *r = rec[buf] + t * MB_STORE_REC_WORDS;
      const u32 valid = windows_valid(r[2], r[5], r[8], p.k);
      if (valid) {
        Window w;
        w.w0 = ((u64)r[0] << 32) | r[1];
        w.w1 = ((u64)r[3] << 32) | r[4];
        w.w2 = ((u64)r[6] << 32) | r[7];
        u64 v = valid;                                  //  bit 31 - q  ->  bit 63 - 2q
        v = (v | (v << 16)) & 0x0000ffff0000ffffull;
        v = (v | (v << 8))  & 0x00ff00ff00ff00ffull;
        v = (v | (v << 4))  & 0x0f0f0f0f0f0f0f0full;
        v = (v | (v << 2))  & 0x3333333333333333ull;
        v = (v | (v << 1))  & 0x5555555555555555ull;
        cand = file_range_slots(w, p.k, p.mode, fLo, fHi) & (v << 1);
      }
    }
    //  the members as a 32-bit mask (bit 31 - q = position q): the loops below walk it with 32-bit instructions
    u32 cand32;
    {
      u64 x = (cand >> 1) & 0x5555555555555555ull;       //  bit 2 (31 - q)
      x = (x | (x >> 1))  & 0x3333333333333333ull;
      x = (x | (x >> 2))  & 0x0f0f0f0f0f0f0f0full;
      x = (x | (x >> 4))  & 0x00ff00ff00ff00ffull;
      x = (x | (x >> 8))  & 0x0000ffff0000ffffull;
      cand32 = (u32)(x | (x >> 16));
    }
    const u32 mine = (u32)__popc(cand32);
    u32 incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
      const u32 x = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (u32)o) incl += x;
    }
    if (lane == 31) sWarpSum[warp] = incl;
    __syncthreads();
    u32 base = incl - mine, total = 0;
    for (u32 wi = 0; wi < THREADS / 32; wi++) { const u32 ws = sWarpSum[wi]; if (wi < warp) base += ws; total += ws; }

    for (u32 round0 = 0; round0 < total; round0 += CAP) {          //  (uniform over the CTA; one round unless the tile is unusually full)
      const u32 nR = min((u32)CAP, total - round0);
      {                                                            //  my members whose place in the tile's list falls into this round
        u32 idx = base - round0;                                   //  (unsigned: places before this round wrap far above CAP)
        for (u32 m = cand32; m; idx++) {
          const u32 q = __clz(m);
          m &= ~(0x80000000u >> q);
          if (idx < (u32)CAP) plist[idx] = (unsigned short)(t * 32 + q);
        }
      }
      __syncthreads();

      //  dense: consecutive lanes, consecutive members
      Stored sk[ITEMS];
      u32    br[ITEMS];                                            //  bin (relative to the pass) << 16 | rank inside (tile, bin)
#pragma unroll
      for (int i = 0; i < ITEMS; i++) {
        const u32 j = (u32)i * THREADS + t;
        if (j < nR) {
          const u32 pos = plist[j];
          const u32 *r  = rec[buf] + (pos >> 5) * MB_STORE_REC_WORDS;
          const u32 q2  = 2 * (pos & 31);
          const u64 a   = ((u64)r[0] << 32) | r[1], b = ((u64)r[3] << 32) | r[4];
          const u64 f   = (q2 ? ((a << q2) | (b >> (64 - q2))) : a) >> (64 - 2 * p.k);              //  k bases from the position on
          u64 key = f;
          if (p.mode != MB_MODE_FORWARD) {
            const u64 rc = revcomp64(f, p.k);
            key = (p.mode == MB_MODE_REVERSE) ? rc : (f < rc ? f : rc);
          }
          const u32 bin = (u32)(key >> binShift) - p.binLo;
          to_stored(key, sk[i]);
          br[i] = (bin << 16) | atomicAdd(&cnt[bin], 1u);
        }
      }
      __syncthreads();

      //  the round's bin counts -> place of every bin's run in the stage buffer, and one global atomic per non-empty
      //  bin reserves its run in the bin's segment
      {
        const u32 perThread = (nT + THREADS - 1) / THREADS;
        const u32 b0 = t * perThread;
        u32 mySum = 0;
        for (u32 i = 0; i < perThread; i++) { const u32 b = b0 + i; if (b < nT) mySum += cnt[b]; }
        u32 inc2 = mySum;
        for (int o = 1; o < 32; o <<= 1) {
          const u32 x = __shfl_up_sync(0xffffffffu, inc2, o);
          if (lane >= (u32)o) inc2 += x;
        }
        __syncthreads();                                           //  (sWarpSum is read above by everybody)
        if (lane == 31) sWarpSum[warp] = inc2;
        __syncthreads();
        u32 run = inc2 - mySum;
        for (u32 wi = 0; wi < warp; wi++) run += sWarpSum[wi];
        for (u32 i = 0; i < perThread; i++) {
          const u32 b = b0 + i;
          if (b < nT) {
            const u32 c = cnt[b];
            binOff[b] = run;
            if (c) {
              gofs[b] = binStart[p.binLo + b] + atomicAdd((unsigned long long *)&binCursor[p.binLo + b], (unsigned long long)c) - run;
              cnt[b]  = 0;
            }
            run += c;
          }
        }
      }
      __syncthreads();

      //  bin-sorted stage buffer (the list of positions is dead: its memory takes the staged keys' bins)
#pragma unroll
      for (int i = 0; i < ITEMS; i++) {
        const u32 j = (u32)i * THREADS + t;
        if (j < nR) {
          const u32 bin = br[i] >> 16;
          const u32 pos = binOff[bin] + (br[i] & 0xffffu);
          stage[pos] = sk[i];
          if (S32) sbin[pos] = (unsigned short)bin;
        }
      }
      __syncthreads();

      //  copy-out: consecutive threads, consecutive staged keys -> every bin's run is written coalesced
      for (u32 j = t; j < nR; j += THREADS) {
        const Stored k2 = stage[j];
        const u32 bin = S32 ? (u32)sbin[j] : (u32)(key_shr64(k2, binShift)) - p.binLo;
        out[gofs[bin] + j] = k2;
      }
      __syncthreads();                                             //  stage / list / tables are reused by the next round or tile
    }
    __syncthreads();                                               //  (tiles without members: all reads of rec[buf] are done)
  }
}

template <bool S32, int ITEMS>
cudaError_t launch_dense(const MbExtractParams &p, void *out, const u64 *binStart, u64 *binCursor, cudaStream_t s) {
  typedef ExtractSmem<1, S32> S;
  typedef typename KeyT<S32 ? 0 : 1>::type Stored;
  const u32 nT = p.binHi - p.binLo;
  const size_t smem = S::OFF_TABLES + S::tablesBytes(nT) + (size_t)256 * ITEMS * 2 + 16 + (size_t)256 * ITEMS * sizeof(Stored) + 16;
  auto kern = extract_dense_kernel<S32, ITEMS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 0, perSm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, 256, smem);
  if (e != cudaSuccess) return e;
  if (perSm < 1) perSm = 1;
  const u64 nRecords = p.nPositions / 32;
  const u64 nTiles   = (nRecords + 255) / 256;
  if (nTiles == 0) return cudaSuccess;
  u64 grid = (u64)sms * perSm;
  if (grid > nTiles) grid = nTiles;
  kern<<<(unsigned)grid, 256, smem, s>>>(p, (Stored *)out, binStart, binCursor, nTiles);
  return cudaGetLastError();
}

//  Does this pass take the dense form?  k <= 32, file-aligned bins, a minority (or a small majority) of the k-mers kept.
static bool dense_pass(const MbExtractParams &p) {
  const bool off = getenv("MB_EXTRACT_DENSE") && getenv("MB_EXTRACT_DENSE")[0] == '0';      //  A/B: the divergent form of extract_kernel
  if (off || p.k > 32 || p.k < 5 || p.binBits < 6 || p.binBits > 10 || p.noPrefilter) return false;
  const u32 perFile = 1u << (p.binBits - 6), nBins = 1u << p.binBits;
  if ((p.binLo & (perFile - 1)) || (p.binHi & (perFile - 1)) || p.binHi - p.binLo >= nBins) return false;
  const u32 kept = p.keptPerMille ? p.keptPerMille : (p.binHi - p.binLo) * 1000 / nBins;
  return kept <= 600;
}

template <bool S32>
cudaError_t launch_dense_sized(const MbExtractParams &p, void *out, const u64 *binStart, u64 *binCursor, cudaStream_t s) {
  const u32 kept = p.keptPerMille ? p.keptPerMille : (p.binHi - p.binLo) * 1000 / (1u << p.binBits);
  u32 want = (u32)(8192ull * kept / 1000 * 13 / 10);                //  1.3 x the expected members of a tile (every position a k-mer)
  if (getenv("MB_DENSE_ITEMS")) want = 256 * (u32)strtoul(getenv("MB_DENSE_ITEMS"), nullptr, 10);   //  testing knob: small values force several rounds per tile
  if (want <= 256 * 8)  return launch_dense<S32, 8 >(p, out, binStart, binCursor, s);
  if (want <= 256 * 12) return launch_dense<S32, 12>(p, out, binStart, binCursor, s);
  if (want <= 256 * 16) return launch_dense<S32, 16>(p, out, binStart, binCursor, s);
  return launch_dense<S32, 24>(p, out, binStart, binCursor, s);
}

template <int KW, bool COUNT_ONLY, bool S32>
cudaError_t launch(const MbExtractParams &p, void *out, const u64 *binStart, u64 *binCursor, u64 *binCount, cudaStream_t s) {
  typedef ExtractSmem<KW, S32> S;
  const u32 nBins = 1u << p.binBits, nT = p.binHi - p.binLo;
  //  stage capacity per tile: everything when the pass keeps all bins, else twice the expected share
  //  (what does not fit is written directly -- slower, never wrong)
  u32 stageCap = S::TILE_POS;
  if (nT < nBins) {
    u64 want = p.keptPerMille ? 2ull * S::TILE_POS * p.keptPerMille / 1000 : 2ull * S::TILE_POS * nT / nBins;
    if (want < 1024) want = 1024;
    want = (want + 255) / 256 * 256;
    if (want < stageCap) stageCap = (u32)want;
  }
  if (sizeof(typename S::Key) * stageCap > 32768) stageCap = 32768 / sizeof(typename S::Key);
  const size_t smem = COUNT_ONLY ? S::countBytes(nT) : S::scatterBytes(nT, stageCap);
  auto kern = extract_kernel<KW, COUNT_ONLY, S32>;

  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;

  int dev = 0, sms = 0, perSm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, S::THREADS, smem);
  if (e != cudaSuccess) return e;
  if (perSm < 1) perSm = 1;

  const u64 nRecords = p.nPositions / 32;
  const u64 nTiles   = (nRecords + S::THREADS - 1) / S::THREADS;
  if (nTiles == 0) return cudaSuccess;
  u64 grid = (u64)sms * perSm;                                   //  persistent: a multiple of the SM count
  if (grid > nTiles) grid = nTiles;

  kern<<<(unsigned)grid, S::THREADS, smem, s>>>(p, (typename KeyT<S32 ? 0 : KW>::type *)out, binStart, binCursor, binCount, nTiles, stageCap);
  return cudaGetLastError();
}

}  //  namespace


cudaError_t mb_launch_bin_histogram(const MbExtractParams &p, u64 *dBinCount, cudaStream_t s) {
  return (p.k <= 32) ? launch<1, true, false>(p, nullptr, nullptr, nullptr, dBinCount, s)
                     : launch<2, true, false>(p, nullptr, nullptr, nullptr, dBinCount, s);
}

cudaError_t mb_launch_extract(const MbExtractParams &p, void *dKeys, const u64 *dBinStart, u64 *dBinCursor, cudaStream_t s) {
  if (dense_pass(p)) {
    if (p.suffix32) {
      if (2 * p.k - p.binBits > 32) return cudaErrorInvalidValue;
      return launch_dense_sized<true>(p, dKeys, dBinStart, dBinCursor, s);
    }
    return launch_dense_sized<false>(p, dKeys, dBinStart, dBinCursor, s);
  }
  if (p.suffix32) {
    if (p.k > 32 || 2 * p.k - p.binBits > 32) return cudaErrorInvalidValue;
    return launch<1, false, true>(p, dKeys, dBinStart, dBinCursor, nullptr, s);
  }
  return (p.k <= 32) ? launch<1, false, false>(p, dKeys, dBinStart, dBinCursor, nullptr, s)
                     : launch<2, false, false>(p, dKeys, dBinStart, dBinCursor, nullptr, s);
}
