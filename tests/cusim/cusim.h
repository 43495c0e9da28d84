//  cusim: a tiny host-side SIMT emulator, TEST INFRASTRUCTURE ONLY.
//
//  The container the kernels are developed in has no GPU.  To check the indexing / synchronisation
//  logic of a kernel before it ever reaches a B200, tests/cusim/translate.py rewrites a .cu file into
//  plain C++ (launch chevrons -> cusim::launch, `extern __shared__` -> cusim::dyn_smem) and this header
//  supplies the CUDA vocabulary the kernels use: every CUDA thread of a block is a fiber (own stack,
//  cooperative switch), __syncthreads / warp collectives are rendezvous points between fibers, blocks
//  run one after the other in launch order (so a decoupled look-back always finds its predecessors
//  published).  "Device memory" is host memory.  Nothing under meryl_b200/ links or includes this.
#pragma once

#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <functional>
#include <algorithm>

#define MB_CUSIM 1

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static
#define __constant__ static

//  ---- vector types ------------------------------------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct dim3  { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct __attribute__((aligned(8)))  uint2      { unsigned x, y; };
struct __attribute__((aligned(16))) uint4      { unsigned x, y, z, w; };
struct __attribute__((aligned(16))) ulonglong2 { unsigned long long x, y; };
static inline uint2      make_uint2(unsigned a, unsigned b)                         { uint2 r; r.x = a; r.y = b; return r; }
static inline uint4      make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { uint4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
static inline ulonglong2 make_ulonglong2(unsigned long long a, unsigned long long b) { ulonglong2 r; r.x = a; r.y = b; return r; }

//  ---- host runtime vocabulary (only what the launch wrappers touch) ------------------------------
typedef int  cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorLaunchFailure = 719 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaDevAttrMultiProcessorCount = 16 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaHostRegisterDefault = 0 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void *devicePointer, *hostPointer; };
struct cudaDeviceProp { char name[256]; int major, minor, multiProcessorCount; size_t totalGlobalMem; };
namespace cusim { void *dev_alloc(size_t bytes); void dev_free(void *p); }
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n)          { *p = (T *)cusim::dev_alloc(n); return *p ? cudaSuccess : 2; }
static inline cudaError_t cudaFree(void *p)                                       { cusim::dev_free(p); return cudaSuccess; }
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t n)      { *p = (T *)malloc(n ? n : 1); return cudaSuccess; }
template <class T> static inline cudaError_t cudaHostAlloc(T **p, size_t n, unsigned) { *p = (T *)malloc(n ? n : 1); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void *p)                                   { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostRegister(void *, size_t, unsigned)              { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void *)                              { return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *p, int v, size_t n)                    { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int)                                      { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n)                              { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize(void)                             { return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned)    { *s = (cudaStream_t)1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s)                       { *s = (cudaStream_t)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t)                         { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t)                     { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e)                         { *e = (cudaEvent_t)1; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned)      { *e = (cudaEvent_t)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t)                           { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t)                       { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t)                    { *f = (size_t)2560 << 20; *t = (size_t)4 << 30; return cudaSuccess; }   //  a "2.5 GB" device
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *) { memset(a, 0, sizeof(*a)); return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)         { memset(p, 0, sizeof(*p)); strcpy(p->name, "cusim"); p->major = 10; p->multiProcessorCount = 4; p->totalGlobalMem = (size_t)4 << 30; return cudaSuccess; }
static inline cudaError_t cudaGetLastError(void)                                  { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError(void)                               { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t)                         { return "cusim"; }
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t)              { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d)                                   { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int *v, int, int)                { *v = 4; return cudaSuccess; }   //  "4 SMs"
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, size_t) { return cudaSuccess; }
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 2; return cudaSuccess; }

//  ---- the emulator --------------------------------------------------------------------------------
namespace cusim {

struct WarpBarrier { unsigned mask, arrived, gen; };   //  one rendezvous in flight (keyed by its lane mask)
struct Warp {
  unsigned char slot[32][16];
  unsigned      pred;                 //  ballot bits
  unsigned      alive;                //  lanes that exist and have not exited
  WarpBarrier   bar[8];               //  disjoint lane sets of a diverged warp may be in different collectives at once
};

struct Fiber {
  uint3     tid;
  unsigned  lane, warpId, linear;
  Warp     *warp;
  void     *sp;                       //  saved stack pointer
  bool      done;
};

extern Fiber   *g_cur;
extern uint3    g_blockIdx;
extern dim3     g_blockDim, g_gridDim;
extern unsigned long long g_clock;    //  advances at every fiber switch
extern unsigned long long g_progress; //  bumped by every event that can unblock somebody

void  launch(dim3 grid, dim3 block, size_t dynSmem, const std::function<void()> &body);

//  Several "ranks" (one GPU process each, in the product) inside this one process: every rank runs its host code as a
//  fiber; rank_yield() lets the others run (the emulated collectives of csrc/exchange.cu wait with it).  Kernel launches
//  stay synchronous inside a rank.
void  run_ranks(unsigned world, const std::function<void(unsigned rank)> &body);
void  rank_yield(void);
bool  in_rank(void);
void *dyn_smem(void);
void  yield(void);                    //  let the other fibers of the block run
void  block_barrier(void);
void  warp_barrier(unsigned mask);
unsigned long long launches(void);    //  kernels launched so far (for tests)

template <class T> static inline T warp_exchange(unsigned mask, const T &v, int srcLane) {
  static_assert(sizeof(T) <= 16, "shuffle payload too large");
  Fiber *f = g_cur;
  memcpy(f->warp->slot[f->lane], &v, sizeof(T));
  warp_barrier(mask);
  T r = v;
  if (srcLane >= 0 && srcLane < 32 && ((f->warp->alive & mask) >> srcLane & 1)) memcpy(&r, f->warp->slot[srcLane], sizeof(T));
  warp_barrier(mask);
  return r;
}

}  //  namespace cusim

#define threadIdx (cusim::g_cur->tid)
#define blockIdx  (cusim::g_blockIdx)
#define blockDim  (cusim::g_blockDim)
#define gridDim   (cusim::g_gridDim)
#define warpSize  32

static inline void __syncthreads(void)                          { cusim::block_barrier(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu)      { cusim::warp_barrier(mask); }
static inline void __threadfence(void)                          { cusim::g_progress++; }
static inline void __threadfence_block(void)                    { }
static inline long long clock64(void)                           { return (long long)cusim::g_clock; }

template <class T> static inline T __shfl_sync(unsigned mask, T v, int srcLane, int width = 32) {
  const int lane = (int)cusim::g_cur->lane;
  return cusim::warp_exchange(mask, v, (lane / width) * width + (srcLane % width));
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const int lane = (int)cusim::g_cur->lane;
  return cusim::warp_exchange(mask, v, (lane % width) >= (int)delta ? lane - (int)delta : -1);
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const int lane = (int)cusim::g_cur->lane;
  return cusim::warp_exchange(mask, v, (lane % width) + (int)delta < width ? lane + (int)delta : -1);
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int laneMask, int width = 32) {
  const int lane = (int)cusim::g_cur->lane;
  return cusim::warp_exchange(mask, v, lane ^ laneMask);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  cusim::Fiber *f = cusim::g_cur;
  unsigned char p = pred ? 1 : 0;
  f->warp->slot[f->lane][0] = p;
  cusim::warp_barrier(mask);
  unsigned r = 0;
  for (int l = 0; l < 32; l++) if (((f->warp->alive & mask) >> l & 1) && f->warp->slot[l][0]) r |= 1u << l;
  cusim::warp_barrier(mask);
  return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
static inline unsigned __match_any_sync(unsigned mask, unsigned v) {
  cusim::Fiber *f = cusim::g_cur;
  memcpy(f->warp->slot[f->lane], &v, 4);
  cusim::warp_barrier(mask);
  unsigned r = 0;
  for (int l = 0; l < 32; l++) {
    unsigned o;  memcpy(&o, f->warp->slot[l], 4);
    if (((f->warp->alive & mask) >> l & 1) && o == v) r |= 1u << l;
  }
  cusim::warp_barrier(mask);
  return r;
}
static inline unsigned __activemask(void) { return cusim::g_cur->warp->alive; }

//  ---- integer intrinsics --------------------------------------------------------------------------
static inline int      __popc(unsigned v)                 { return __builtin_popcount(v); }
static inline int      __popcll(unsigned long long v)     { return __builtin_popcountll(v); }
static inline int      __clz(int v)                       { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int      __clzll(long long v)               { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int      __ffs(int v)                       { return __builtin_ffs(v); }
static inline int      __ffsll(long long v)               { return __builtin_ffsll(v); }
static inline unsigned __brev(unsigned v) {
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
  return __builtin_bswap32(v);
}
static inline unsigned long long __brevll(unsigned long long v) {
  return ((unsigned long long)__brev((unsigned)v) << 32) | __brev((unsigned)(v >> 32));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) {
  s &= 31;  return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) {
  s &= 31;  return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
  const unsigned long long v = ((unsigned long long)b << 32) | a;
  unsigned r = 0;
  for (int i = 0; i < 4; i++) {
    const unsigned s = (sel >> (4 * i)) & 0xf;
    unsigned byte = (unsigned)(v >> (8 * (s & 7))) & 0xff;
    if (s & 8) byte = (byte & 0x80) ? 0xff : 0;
    r |= byte << (8 * i);
  }
  return r;
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
template <class T> static inline T __ldg(const T *p)    { return *p; }
using std::min;
using std::max;

//  ---- atomics (one OS thread: plain read-modify-write) -----------------------------------------------
#define CUSIM_ATOMIC(name, T, expr)                                              \
  static inline T name(T *p, T v) { const T old = *p; *p = (expr); cusim::g_progress++; return old; }
CUSIM_ATOMIC(atomicAdd,  unsigned,           old + v)
CUSIM_ATOMIC(atomicAdd,  int,                old + v)
CUSIM_ATOMIC(atomicAdd,  unsigned long long, old + v)
CUSIM_ATOMIC(atomicSub,  unsigned,           old - v)
CUSIM_ATOMIC(atomicOr,   unsigned,           old | v)
CUSIM_ATOMIC(atomicOr,   unsigned long long, old | v)
CUSIM_ATOMIC(atomicAnd,  unsigned,           old & v)
CUSIM_ATOMIC(atomicAnd,  unsigned long long, old & v)
CUSIM_ATOMIC(atomicXor,  unsigned,           old ^ v)
CUSIM_ATOMIC(atomicMax,  unsigned,           old > v ? old : v)
CUSIM_ATOMIC(atomicMax,  unsigned long long, old > v ? old : v)
CUSIM_ATOMIC(atomicMin,  unsigned,           old < v ? old : v)
CUSIM_ATOMIC(atomicMin,  unsigned long long, old < v ? old : v)
CUSIM_ATOMIC(atomicExch, unsigned,           v)
CUSIM_ATOMIC(atomicExch, unsigned long long, v)
#undef CUSIM_ATOMIC
static inline unsigned atomicCAS(unsigned *p, unsigned cmp, unsigned v) {
  const unsigned old = *p;  if (old == cmp) *p = v;  cusim::g_progress++;  return old;
}
static inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long cmp, unsigned long long v) {
  const unsigned long long old = *p;  if (old == cmp) *p = v;  cusim::g_progress++;  return old;
}
