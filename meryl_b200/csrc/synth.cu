//  Seeded synthetic read generator, on the device (bench / test input; SURVEY.md 8(d) shapes).
//
//  Reads of readLen bases are drawn from a virtual uniform-random genome of genomeLen bases
//  (base g = hash(seed, g): nothing is stored), start uniform, strand 50/50, substitution errors
//  at rate err.  Output is text, produced 16 bytes per thread so stores are 128-bit and coalesced:
//    format 0   work-buffer form  "<bases>."                       (readLen+1 bytes per read)
//    format 1   4-line FASTQ      "@r%012llu\n<bases>\n+\n<I...>\n"  (2*readLen+19 bytes per read)
#include "common.h"
#include "kernels.cuh"

namespace {

__device__ __forceinline__ u64 mix64(u64 x) {                 //  splitmix64 finaliser
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

struct SynthParams {
  u64 firstRead, nReads, genomeLen, seed;
  u32 readLen, errThreshold;                                  //  err * 2^32
  u32 format, recBytes;
};

__device__ __forceinline__ u32 read_base(const SynthParams &p, u64 r, u32 j, u64 start, bool rev) {
  const u64 g = rev ? start + (p.readLen - 1 - j) : start + j;
  u32 c = (u32)mix64(p.seed ^ (g * 0x2545f4914f6cdd1dull)) & 3;        //  genome base, ACGT order
  if (rev) c = 3 - c;
  const u64 h = mix64((p.seed + 0x51ed27) ^ ((r * p.readLen + j) * 0x9e3779b97f4a7c15ull));
  if ((u32)h < p.errThreshold) c = (c + 1 + (u32)((h >> 32) % 3)) & 3;
  return c;
}

__device__ __forceinline__ unsigned char synth_byte(const SynthParams &p, u64 r, u32 off, u64 start, bool rev) {
  const u32 L = p.readLen;
  if (p.format == 0)
    return off < L ? "ACGT"[read_base(p, r, off, start, rev)] : '.';
  //  FASTQ
  if (off == 0) return '@';
  if (off == 1) return 'r';
  if (off < 14) {                                             //  12 decimal digits, zero padded
    u64 v = r;
    for (u32 i = off; i < 13; i++) v /= 10;
    return (unsigned char)('0' + v % 10);
  }
  if (off == 14) return '\n';
  off -= 15;
  if (off < L)  return "ACGT"[read_base(p, r, off, start, rev)];
  if (off == L) return '\n';
  off -= L + 1;
  if (off == 0) return '+';
  if (off == 1) return '\n';
  off -= 2;
  return off < L ? 'I' : '\n';
}

__global__ void __launch_bounds__(256)
synth_kernel(SynthParams p, unsigned char *__restrict__ out, u64 nBytes) {
  const u64 b0 = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (b0 >= nBytes) return;
  u64 r   = b0 / p.recBytes;
  u32 off = (u32)(b0 - r * p.recBytes);
  unsigned char v[16];
  u64 rr = ~0ull, start = 0;  bool rev = false;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    if (r != rr) {
      rr = r;
      const u64 gr = p.firstRead + r;
      const u64 h  = mix64((p.seed + 0x7f4a7c15) ^ (gr * 0xd6e8feb86659fd93ull));
      start = mix64(h) % (p.genomeLen - p.readLen + 1);
      rev   = (h >> 63) != 0;
    }
    v[i] = (b0 + i < nBytes) ? synth_byte(p, p.firstRead + r, off, start, rev) : 0;
    if (++off == p.recBytes) { off = 0; r++; }
  }
  if (b0 + 16 <= nBytes) {
    *(uint4 *)(out + b0) = *(uint4 *)v;
  } else {
    for (int i = 0; i < 16 && b0 + i < nBytes; i++) out[b0 + i] = v[i];
  }
}

}  //  namespace

extern "C" uint64_t mb_synth_record_bytes(uint32_t readLen, uint32_t format) {
  return format == 0 ? (uint64_t)readLen + 1 : 2 * (uint64_t)readLen + 19;
}

extern "C" int mb_synth_reads_device(void *dOut, uint64_t firstRead, uint64_t nReads, uint32_t readLen,
                                     uint64_t genomeLen, double err, uint64_t seed, uint32_t format, void *stream) {
  mb_clear_error();
  if (!dOut && nReads)                         return mb_fail(MB_EINVAL, "mb_synth_reads_device: NULL output");
  if (format > 1 || readLen == 0 || genomeLen < readLen) return mb_fail(MB_EINVAL, "mb_synth_reads_device: bad arguments");
  if (((uintptr_t)dOut & 15) != 0)             return mb_fail(MB_EINVAL, "mb_synth_reads_device: output must be 16-byte aligned");
  if (nReads == 0) return MB_OK;
  SynthParams p;
  p.firstRead = firstRead;  p.nReads = nReads;  p.genomeLen = genomeLen;  p.seed = seed;
  p.readLen = readLen;  p.format = format;
  p.errThreshold = err <= 0 ? 0 : err >= 1 ? 0xffffffffu : (u32)(err * 4294967296.0);
  p.recBytes = (u32)mb_synth_record_bytes(readLen, format);
  const u64 nBytes = nReads * p.recBytes;
  const u64 threads = (nBytes + 15) / 16;
  synth_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, (unsigned char *)dOut, nBytes);
  MB_CUDA(cudaGetLastError());
  return MB_OK;
}
