//  K1: ASCII base stream -> 2-bit packed store with a validity bit per position.
//
//  Replaces the per-base work of kmerIterator::isACGT + the 2-bit encode inside
//  kmerTiny::addR  (src/utility/src/kmers-v1/kmers-iterator.H:98-103, kmers-tiny.H:86):
//  code = (ascii >> 1) & 3  (A=0 C=1 T=2 G=3, either case), anything that is not ACGTacgt -- the
//  '.' record breaks of the work buffer (merylOp-countThreads.C:214-215), N, IUPAC codes -- is an
//  invalid position that no k-mer may cover.
//
//  One thread packs one 32-position record: 32 B in (two 128-bit loads), 12 B out.
//  HBM-bound: 1 B + 0.375 B per position.

#include "kernels.cuh"

namespace {

__device__ __forceinline__ void pack4(u32 w, u64 &bases, u32 &valid) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const u32 c = (w >> (8 * i)) & 0xff;
    const u32 u = c & 0xdf;                                      //  fold case
    const u32 ok = (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T');
    bases = (bases << 2) | ((c >> 1) & 3);
    valid = (valid << 1) | ok;
  }
}

__global__ void __launch_bounds__(256)
pack_kernel(const uint8_t *__restrict__ ascii, u64 nBytes, u32 *__restrict__ store, u64 firstRecord, u64 nRecords) {
  const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nRecords) return;

  const u64 base = r * 32;
  u64 bases = 0;
  u32 valid = 0;

  if (base + 32 <= nBytes) {
    const uint4 a = __ldg((const uint4 *)(ascii + base));
    const uint4 b = __ldg((const uint4 *)(ascii + base + 16));
    pack4(a.x, bases, valid);  pack4(a.y, bases, valid);  pack4(a.z, bases, valid);  pack4(a.w, bases, valid);
    pack4(b.x, bases, valid);  pack4(b.y, bases, valid);  pack4(b.z, bases, valid);  pack4(b.w, bases, valid);
  } else {
    for (u32 j = 0; j < 32; j++) {                               //  ragged tail: pad with invalid positions
      const u32 c  = (base + j < nBytes) ? ascii[base + j] : '.';
      const u32 u  = c & 0xdf;
      const u32 ok = (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T');
      bases = (bases << 2) | ((c >> 1) & 3);
      valid = (valid << 1) | ok;
    }
  }

  u32 *rec = store + (firstRecord + r) * MB_STORE_REC_WORDS;
  rec[0] = (u32)(bases >> 32);
  rec[1] = (u32)bases;
  rec[2] = valid;
}

}  //  namespace


cudaError_t mb_launch_pack(const uint8_t *dAscii, u64 nBytes, u32 *dStore, u64 firstRecord, cudaStream_t s) {
  const u64 nRecords = (nBytes + 31) / 32;
  if (nRecords == 0) return cudaSuccess;
  const u32 threads = 256;
  const u64 blocks  = (nRecords + threads - 1) / threads;
  pack_kernel<<<(unsigned)blocks, threads, 0, s>>>(dAscii, nBytes, dStore, firstRecord, nRecords);
  return cudaGetLastError();
}
