//  P1/P2 on the device: FASTA/FASTQ file bytes -> base stream, as a parallel finite-state machine.
//
//  The host parser (parser.cpp; the reference's bufSeqFile::loadBases + work-buffer builder, src/utility/src/sequence/
//  bufSeqFile-v1.C:554-595, src/meryl/merylOp-countThreads.C:139-231) is a byte-level machine with 11 states that consumes
//  one byte at a time and emits at most one character per byte (the byte itself inside a sequence word, or the '.' that ends
//  a record).  Its transition table is not restated here: it is TABULATED from SeqParser itself at start-up (one byte of each
//  of the 8 byte classes fed from each state), so the two parsers cannot drift apart.
//
//  Parallel form (the file is handled in blocks of <= 64 MB, cut anywhere -- the state and the unaligned tail of the
//  emitted stream carry over):
//    K11a parse_map_kernel    a thread per 256-byte chunk runs the machine from EVERY state at once: end state and number of
//                             emitted characters per start state (chunks staged through shared memory, coalesced loads)
//    K11b parse_group_kernel  a thread per 64 chunks composes their maps;  parse_walk_kernel (one thread) walks the group maps
//                             from the block's true start state;  parse_spread_kernel gives every chunk its true start state
//                             and output offset
//    K11c parse_emit_kernel   a thread per chunk runs the machine once more from the true state, compacting the emitted
//                             characters in place in shared memory; the CTA then writes its run of output bytes coalesced
//  Bytes per input byte: 2 reads (K11a, K11c) + <= 1 write + 40/256 (maps).  The stream goes straight to K1 (pack) in pieces
//  that are multiples of 32 positions.  Experimental until it has run on a GPU: MB_DEVICE_PARSE=1 routes mb_add_file_bytes /
//  mb_add_file through it (engine.cu).

#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "engine.h"
#include "kernels.cuh"

namespace {

const int PS_STATES  = 12;                                 //  SeqParser's 11 + an absorbing error state
const int PS_ERR     = 11;
const int PS_CLASSES = 8;
const int PS_CHUNK   = 256;                                //  bytes per thread
const int PS_THREADS = 128;                                //  chunks per CTA
const int PS_PITCH   = PS_CHUNK + 4;                       //  shared-memory pitch of a chunk: lanes hit different banks
const int PS_GROUP   = 64;                                 //  chunks per group

struct ParseTable { unsigned char t[PS_STATES][PS_CLASSES]; };   //  bits 0-3 next state, bit 4 emit the byte, bit 5 emit '.'

struct ChunkMap { unsigned short cnt[PS_STATES]; unsigned char end[PS_STATES]; unsigned char pad[4]; };   //  40 bytes
struct GroupMap { u32 cnt[PS_STATES]; unsigned char end[PS_STATES]; unsigned char pad[4]; };               //  64 bytes
struct ParseCtl { u64 startOffset; u32 startState; u32 endState; u64 endOffset; };

__host__ __device__ __forceinline__ u32 byte_class(unsigned char c) {
  if (c == '\n') return 0;
  if (c == '\r') return 1;
  if (c == ' ' || c == '\t' || c == '\v' || c == '\f') return 2;
  if (c == '>') return 3;
  if (c == '@') return 4;
  if (c == '+') return 5;
  if (c >= '!' && c <= '~') return 6;
  return 7;
}

//  The CTA's 32 KB of input into shared memory, chunk c at c * PS_PITCH (bytes past the end of the block read as LF and are
//  never walked).
__device__ __forceinline__ void load_tile(const unsigned char *__restrict__ raw, u64 nBytes, u64 tileStart, unsigned char *tile) {
  const u32 *raw32 = (const u32 *)(raw + tileStart);       //  raw and tileStart are 16-byte aligned
  for (u32 w = threadIdx.x; w < PS_THREADS * PS_CHUNK / 4; w += PS_THREADS) {
    const u64 at = tileStart + (u64)w * 4;
    u32 v = 0x0a0a0a0au;
    if (at + 4 <= nBytes) v = raw32[w];
    else if (at < nBytes) { v = 0; for (u32 i = 0; i < 4; i++) v |= (u32)((at + i < nBytes) ? raw[at + i] : (unsigned char)'\n') << (8 * i); }
    const u32 b = w * 4;
    *(u32 *)(tile + (b / PS_CHUNK) * PS_PITCH + (b % PS_CHUNK)) = v;
  }
}

__global__ void __launch_bounds__(PS_THREADS)
parse_map_kernel(const unsigned char *__restrict__ raw, u64 nBytes, ParseTable tab, ChunkMap *__restrict__ maps) {
  __shared__ __align__(16) unsigned char tile[PS_THREADS * PS_PITCH];
  __shared__ unsigned char stab[PS_STATES * PS_CLASSES];     //  the table in shared memory: every lane indexes it differently
  if (threadIdx.x < PS_STATES * PS_CLASSES) stab[threadIdx.x] = tab.t[threadIdx.x / PS_CLASSES][threadIdx.x % PS_CLASSES];
  const u64 tileStart = (u64)blockIdx.x * PS_THREADS * PS_CHUNK;
  load_tile(raw, nBytes, tileStart, tile);
  __syncthreads();
  const u64 chunkStart = tileStart + (u64)threadIdx.x * PS_CHUNK;
  if (chunkStart >= nBytes) return;
  const u32 len = (u32)(nBytes - chunkStart < (u64)PS_CHUNK ? nBytes - chunkStart : (u64)PS_CHUNK);
  const unsigned char *my = tile + threadIdx.x * PS_PITCH;
  u32 cur[PS_STATES], cnt[PS_STATES];
#pragma unroll
  for (int s = 0; s < PS_STATES; s++) { cur[s] = s; cnt[s] = 0; }
  for (u32 i = 0; i < len; i++) {
    const u32 c = byte_class(my[i]);
#pragma unroll
    for (int s = 0; s < PS_STATES; s++) {
      const u32 t = stab[cur[s] * PS_CLASSES + c];
      cur[s] = t & 15;
      cnt[s] += (t >> 4) ? 1 : 0;
    }
  }
  ChunkMap m;
#pragma unroll
  for (int s = 0; s < PS_STATES; s++) { m.cnt[s] = (unsigned short)cnt[s]; m.end[s] = (unsigned char)cur[s]; }
  m.pad[0] = m.pad[1] = m.pad[2] = m.pad[3] = 0;
  maps[chunkStart / PS_CHUNK] = m;
}

__global__ void __launch_bounds__(128)
parse_group_kernel(const ChunkMap *__restrict__ maps, u64 nChunks, GroupMap *__restrict__ groups, u64 nGroups) {
  const u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nGroups) return;
  GroupMap out;
  for (int s = 0; s < PS_STATES; s++) { out.end[s] = (unsigned char)s; out.cnt[s] = 0; }
  const u64 c0 = g * PS_GROUP, c1 = c0 + PS_GROUP < nChunks ? c0 + PS_GROUP : nChunks;
  for (u64 c = c0; c < c1; c++)
    for (int s = 0; s < PS_STATES; s++) {
      const u32 e = out.end[s];
      out.cnt[s] += maps[c].cnt[e];
      out.end[s]  = maps[c].end[e];
    }
  out.pad[0] = out.pad[1] = out.pad[2] = out.pad[3] = 0;
  groups[g] = out;
}

__global__ void parse_walk_kernel(const GroupMap *__restrict__ groups, u64 nGroups, ParseCtl *ctl,
                                  unsigned char *__restrict__ groupState, u64 *__restrict__ groupOffset) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  u32 st  = ctl->startState;
  u64 off = ctl->startOffset;
  for (u64 g = 0; g < nGroups; g++) {
    groupState[g] = (unsigned char)st;  groupOffset[g] = off;
    off += groups[g].cnt[st];
    st   = groups[g].end[st];
  }
  ctl->endState = st;  ctl->endOffset = off;
}

__global__ void __launch_bounds__(128)
parse_spread_kernel(const ChunkMap *__restrict__ maps, u64 nChunks, const unsigned char *__restrict__ groupState,
                    const u64 *__restrict__ groupOffset, u64 nGroups, const ParseCtl *__restrict__ ctl,
                    unsigned char *__restrict__ chunkState, u64 *__restrict__ chunkOffset /* nChunks + 1 */) {
  const u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nGroups) return;
  u32 st  = groupState[g];
  u64 off = groupOffset[g];
  const u64 c0 = g * PS_GROUP, c1 = c0 + PS_GROUP < nChunks ? c0 + PS_GROUP : nChunks;
  for (u64 c = c0; c < c1; c++) {
    chunkState[c] = (unsigned char)st;  chunkOffset[c] = off;
    off += maps[c].cnt[st];
    st   = maps[c].end[st];
  }
  if (c1 == nChunks) chunkOffset[nChunks] = off;           //  (== ctl->endOffset)
}

__global__ void __launch_bounds__(PS_THREADS)
parse_emit_kernel(const unsigned char *__restrict__ raw, u64 nBytes, ParseTable tab, const unsigned char *__restrict__ chunkState,
                  const u64 *__restrict__ chunkOffset, u64 nChunks, unsigned char *__restrict__ out) {
  __shared__ __align__(16) unsigned char tile[PS_THREADS * PS_PITCH];
  __shared__ u32 rel[PS_THREADS + 1];
  __shared__ unsigned char stab[PS_STATES * PS_CLASSES];
  if (threadIdx.x < PS_STATES * PS_CLASSES) stab[threadIdx.x] = tab.t[threadIdx.x / PS_CLASSES][threadIdx.x % PS_CLASSES];
  const u64 tileStart = (u64)blockIdx.x * PS_THREADS * PS_CHUNK;
  load_tile(raw, nBytes, tileStart, tile);
  const u64 firstChunk = (u64)blockIdx.x * PS_THREADS;
  const u64 myChunk = firstChunk + threadIdx.x;
  const u64 base = chunkOffset[firstChunk];
  rel[threadIdx.x] = (u32)(chunkOffset[myChunk < nChunks ? myChunk : nChunks] - base);
  if (threadIdx.x == 0) { const u64 last = firstChunk + PS_THREADS < nChunks ? firstChunk + PS_THREADS : nChunks; rel[PS_THREADS] = (u32)(chunkOffset[last] - base); }
  __syncthreads();
  if (myChunk < nChunks) {
    const u64 chunkStart = myChunk * PS_CHUNK;
    const u32 len = (u32)(nBytes - chunkStart < (u64)PS_CHUNK ? nBytes - chunkStart : (u64)PS_CHUNK);
    unsigned char *my = tile + threadIdx.x * PS_PITCH;
    u32 st = chunkState[myChunk], o = 0;
    for (u32 i = 0; i < len; i++) {                         //  compaction in place: o <= i
      const unsigned char b = my[i];
      const u32 t = stab[st * PS_CLASSES + byte_class(b)];
      st = t & 15;
      if (t & 16)      my[o++] = b;
      else if (t & 32) my[o++] = '.';
    }
  }
  __syncthreads();
  const u32 total = rel[PS_THREADS];
  for (u32 f = threadIdx.x; f < total; f += PS_THREADS) {   //  consecutive threads, consecutive output bytes
    u32 lo = 0, hi = PS_THREADS;                            //  last chunk with rel[chunk] <= f
    while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (rel[mid] <= f) lo = mid; else hi = mid; }
    out[base + f] = tile[lo * PS_PITCH + (f - rel[lo])];
  }
}

}  //  namespace


//  ---- host side ---------------------------------------------------------------------------------------------------

struct MbDeviceParser {
  cudaStream_t   s        = nullptr;
  size_t         block    = (size_t)64 << 20;              //  bytes of the file per round (MB_DPARSE_BLOCK: testing knob)
  unsigned char *dRaw     = nullptr, *dOut = nullptr, *dChunkState = nullptr, *dGroupState = nullptr;
  ChunkMap      *dMaps    = nullptr;
  GroupMap      *dGroups  = nullptr;
  u64           *dChunkOffset = nullptr, *dGroupOffset = nullptr;
  ParseCtl      *dCtl     = nullptr;
  ParseTable     tab;
  u32            state    = 0;                             //  SeqParser::A
  size_t         carry    = 0;                             //  emitted characters not yet handed on (< 32), at the front of dOut
  u64            launches = 0;
};

static int buildParseTable(ParseTable &tab) {
  static const unsigned char rep[PS_CLASSES] = { '\n', '\r', ' ', '>', '@', '+', 'A', 0x01 };
  for (int st = 0; st < PS_STATES; st++)
    for (int c = 0; c < PS_CLASSES; c++) {
      if (st == PS_ERR) { tab.t[st][c] = PS_ERR; continue; }
      std::string got;
      SeqParser q([](void *a, const char *b, size_t n) -> int { ((std::string *)a)->append(b, n); return 0; }, &got);
      q.setState((SeqParser::State)st);
      const int rc = q.feed(&rep[c], 1);
      if (rc != MB_OK) { tab.t[st][c] = PS_ERR; continue; }
      if (got.size() > 1 || (unsigned)q.state() >= (unsigned)PS_ERR) return mb_fail(MB_EINVAL, "internal: the host parser emits %zu characters for one byte", got.size());
      unsigned char e = (unsigned char)q.state();
      if (got.size() == 1) e |= (got[0] == '.' && rep[c] != '.') ? 32 : 16;
      tab.t[st][c] = e;
    }
  mb_clear_error();
  return MB_OK;
}

int mb_dparse_create(MbDeviceParser **out, cudaStream_t s) {
  MbDeviceParser *P = new MbDeviceParser;
  P->s = s;
  if (getenv("MB_DPARSE_BLOCK")) P->block = (size_t)strtoull(getenv("MB_DPARSE_BLOCK"), nullptr, 10);
  P->block = (P->block + PS_THREADS * PS_CHUNK - 1) / (PS_THREADS * PS_CHUNK) * (PS_THREADS * PS_CHUNK);
  if (P->block == 0) P->block = PS_THREADS * PS_CHUNK;
  int rc = buildParseTable(P->tab);
  const size_t nChunks = P->block / PS_CHUNK, nGroups = (nChunks + PS_GROUP - 1) / PS_GROUP;
  if (rc == MB_OK &&
      (cudaMalloc((void **)&P->dRaw, P->block + 64) != cudaSuccess || cudaMalloc((void **)&P->dOut, P->block + 256) != cudaSuccess ||
       cudaMalloc((void **)&P->dMaps, nChunks * sizeof(ChunkMap)) != cudaSuccess || cudaMalloc((void **)&P->dGroups, nGroups * sizeof(GroupMap)) != cudaSuccess ||
       cudaMalloc((void **)&P->dChunkState, nChunks + 64) != cudaSuccess || cudaMalloc((void **)&P->dChunkOffset, (nChunks + 1) * 8) != cudaSuccess ||
       cudaMalloc((void **)&P->dGroupState, nGroups + 64) != cudaSuccess || cudaMalloc((void **)&P->dGroupOffset, nGroups * 8) != cudaSuccess ||
       cudaMalloc((void **)&P->dCtl, sizeof(ParseCtl)) != cudaSuccess)) {
    cudaGetLastError();
    rc = mb_fail(MB_ENOMEM, "cannot allocate the device parser's buffers (%zu byte blocks)", P->block);
  }
  if (rc != MB_OK) { mb_dparse_destroy(P); return rc; }
  *out = P;
  return MB_OK;
}

void mb_dparse_destroy(MbDeviceParser *P) {
  if (!P) return;
  void *all[] = { P->dRaw, P->dOut, P->dMaps, P->dGroups, P->dChunkState, P->dChunkOffset, P->dGroupState, P->dGroupOffset, P->dCtl };
  for (void *p : all) if (p) cudaFree(p);
  delete P;
}

u64 mb_dparse_launches(const MbDeviceParser *P) { return P->launches; }

//  sink(arg, dStream, n): n characters of the base stream in device memory (n a multiple of 32 except for the last call of
//  mb_dparse_finish); the memory is reused when the sink returns.
int mb_dparse_feed(MbDeviceParser *P, const uint8_t *data, size_t n, MbDeviceSink sink, void *arg) {
  for (size_t at = 0; at < n; at += P->block) {
    const size_t len = n - at < P->block ? n - at : P->block;
    const u64 nChunks = (len + PS_CHUNK - 1) / PS_CHUNK, nGroups = (nChunks + PS_GROUP - 1) / PS_GROUP;
    const unsigned tiles = (unsigned)((nChunks + PS_THREADS - 1) / PS_THREADS);
    ParseCtl ctl;
    ctl.startOffset = P->carry;  ctl.startState = P->state;  ctl.endState = 0;  ctl.endOffset = 0;
    MB_CUDA(cudaMemcpyAsync(P->dRaw, data + at, len, cudaMemcpyHostToDevice, P->s));
    MB_CUDA(cudaMemcpyAsync(P->dCtl, &ctl, sizeof(ctl), cudaMemcpyHostToDevice, P->s));
    parse_map_kernel<<<tiles, PS_THREADS, 0, P->s>>>(P->dRaw, len, P->tab, P->dMaps);
    parse_group_kernel<<<(unsigned)((nGroups + 127) / 128), 128, 0, P->s>>>(P->dMaps, nChunks, P->dGroups, nGroups);
    parse_walk_kernel<<<1, 32, 0, P->s>>>(P->dGroups, nGroups, P->dCtl, P->dGroupState, P->dGroupOffset);
    parse_spread_kernel<<<(unsigned)((nGroups + 127) / 128), 128, 0, P->s>>>(P->dMaps, nChunks, P->dGroupState, P->dGroupOffset, nGroups, P->dCtl,
                                                                          P->dChunkState, P->dChunkOffset);
    parse_emit_kernel<<<tiles, PS_THREADS, 0, P->s>>>(P->dRaw, len, P->tab, P->dChunkState, P->dChunkOffset, nChunks, P->dOut);
    MB_CUDA(cudaGetLastError());
    MB_CUDA(cudaMemcpyAsync(&ctl, P->dCtl, sizeof(ctl), cudaMemcpyDeviceToHost, P->s));
    MB_CUDA(cudaStreamSynchronize(P->s));
    P->launches += 5;
    if (ctl.endState == (u32)PS_ERR)
      return mb_fail(MB_EFORMAT, "a byte that is neither visible nor white space inside a sequence (the reference does not terminate on it)");
    P->state = ctl.endState;
    const size_t have = (size_t)ctl.endOffset, whole = have / 32 * 32;
    if (whole) {
      MB_TRY(sink(arg, (const char *)P->dOut, whole));
      if (have > whole) MB_CUDA(cudaMemcpyAsync(P->dOut, P->dOut + whole, have - whole, cudaMemcpyDeviceToDevice, P->s));   //  (< 32 bytes, behind the first 32)
    }
    P->carry = have - whole;
  }
  return MB_OK;
}

//  End of the file: the record break the work-buffer builder adds (merylOp-countThreads.C:196), then whatever is left.
int mb_dparse_finish(MbDeviceParser *P, MbDeviceSink sink, void *arg) {
  MB_CUDA(cudaMemsetAsync(P->dOut + P->carry, '.', 1, P->s));
  MB_CUDA(cudaStreamSynchronize(P->s));
  const size_t n = P->carry + 1;
  P->carry = 0;  P->state = 0;
  return sink(arg, (const char *)P->dOut, n);
}
