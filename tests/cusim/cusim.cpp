//  cusim runtime: fibers + block scheduler.  TEST INFRASTRUCTURE ONLY (see cusim.h).
#include "cusim.h"

#include <sys/mman.h>
#include <time.h>
#include <vector>

#if !defined(__x86_64__)
#error "cusim's context switch is written for x86-64"
#endif

//  cusim_switch(&fromSp, toSp): save callee-saved registers on the current stack, switch stacks.
extern "C" void cusim_switch(void **fromSp, void *toSp);
asm(".text\n"
    ".globl cusim_switch\n"
    ".type cusim_switch,@function\n"
    "cusim_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size cusim_switch,.-cusim_switch\n");

namespace cusim {

Fiber   *g_cur = nullptr;
uint3    g_blockIdx;
dim3     g_blockDim, g_gridDim;
unsigned long long g_clock = 0, g_progress = 0;

namespace {

const size_t kStackBytes = (size_t)256 << 10;
const unsigned kMaxThreads = 1024;

unsigned char *g_stacks = nullptr;
void          *g_schedSp = nullptr;
std::vector<Fiber> g_fibers;
std::vector<Warp>  g_warps;
const std::function<void()> *g_body = nullptr;
unsigned g_alive = 0, g_barArrived = 0, g_barGen = 0;
unsigned char *g_dynSmem = nullptr;
size_t   g_dynCap = 0;
unsigned long long g_launches = 0;

void fiber_main() {
  (*g_body)();
  Fiber *f = g_cur;
  f->done = true;
  f->warp->alive &= ~(1u << f->lane);
  g_alive--;
  g_progress++;
  //  an exited thread no longer takes part in barriers: release whoever is waiting for it
  if (g_alive && g_barArrived == g_alive) { g_barArrived = 0; g_barGen++; }
  for (WarpBarrier &b : f->warp->bar)
    if (b.arrived && b.arrived >= (unsigned)__builtin_popcount(f->warp->alive & b.mask)) { b.arrived = 0; b.mask = 0; b.gen++; }
  cusim_switch(&f->sp, g_schedSp);
  abort();                                              //  a finished fiber is never resumed
}

void prepare(Fiber &f, unsigned idx) {
  unsigned char *top = g_stacks + (size_t)(idx + 1) * kStackBytes;
  void **sp = (void **)top;
  *--sp = nullptr;                                      //  fake return address of fiber_main (keeps the ABI alignment)
  *--sp = (void *)fiber_main;
  for (int i = 0; i < 6; i++) *--sp = nullptr;         //  rbp rbx r12 r13 r14 r15
  f.sp = sp;
}

}  //  namespace

void *dyn_smem(void) { return g_dynSmem; }

//  "Device" allocations: anonymous mappings (large arenas stay untouched until used); small ones are
//  filled with garbage, like fresh device memory, so a kernel cannot rely on zeroed buffers.
void *dev_alloc(size_t bytes) {
  //  [ header page | the allocation, its END (rounded up to 256 bytes) on a page boundary | inaccessible page ]: running over
  //  the end of a device allocation stops the process here, like the smaller ones compute-sanitizer would report
  const size_t body  = ((bytes ? bytes : 1) + 4095) / 4096 * 4096;
  const size_t total = 4096 + body + 4096;
  unsigned char *p = (unsigned char *)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (p == MAP_FAILED) return nullptr;
  if (mprotect(p + 4096 + body, 4096, PROT_NONE) != 0) { munmap(p, total); return nullptr; }
  const size_t rounded = ((bytes ? bytes : 1) + 255) / 256 * 256;
  unsigned char *q = p + 4096 + body - rounded;
  ((size_t *)p)[0] = total;
  //  fresh device memory is not zero: poison it (large arenas only when CUSIM_POISON_ALL is set -- touching gigabytes is slow)
  static const bool poisonAll = getenv("CUSIM_POISON_ALL") != nullptr;
  if (poisonAll || bytes <= ((size_t)64 << 20)) memset(q, 0xa5, bytes);
  ((void **)(q - 8))[0] = p;                              //  (the slack in front of q: where dev_free finds the mapping)
  return q;
}
void dev_free(void *q) {
  if (!q) return;
  unsigned char *p = (unsigned char *)((void **)((unsigned char *)q - 8))[0];
  munmap(p, ((size_t *)p)[0]);
}
unsigned long long launches(void) { return g_launches; }

void yield(void) {
  g_clock++;
  Fiber *f = g_cur;
  cusim_switch(&f->sp, g_schedSp);
}

void block_barrier(void) {
  const unsigned gen = g_barGen;
  if (++g_barArrived == g_alive) { g_barArrived = 0; g_barGen++; g_progress++; return; }
  while (g_barGen == gen) yield();
}

void warp_barrier(unsigned mask) {
  Warp *w = g_cur->warp;
  const unsigned expect = (unsigned)__builtin_popcount(w->alive & mask);
  WarpBarrier *b = nullptr;
  for (WarpBarrier &x : w->bar) if (x.arrived && x.mask == mask) { b = &x; break; }      //  join the rendezvous of my lane set
  if (!b) for (WarpBarrier &x : w->bar) if (!x.arrived) { b = &x; b->mask = mask; break; }
  if (!b) { fprintf(stderr, "cusim: more than 8 warp collectives in flight in one warp\n"); abort(); }
  const unsigned gen = b->gen;
  if (++b->arrived >= expect) { b->arrived = 0; b->mask = 0; b->gen++; g_progress++; return; }
  while (b->gen == gen) yield();
}

void launch(dim3 grid, dim3 block, size_t dynSmem, const std::function<void()> &body) {
  const unsigned nThreads = block.x * block.y * block.z;
  if (nThreads == 0 || nThreads > kMaxThreads) { fprintf(stderr, "cusim: bad block size %u\n", nThreads); abort(); }
  if (g_cur) { fprintf(stderr, "cusim: nested launch\n"); abort(); }
  if (!g_stacks) {
    g_stacks = (unsigned char *)mmap(nullptr, kStackBytes * kMaxThreads, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (g_stacks == MAP_FAILED) { perror("cusim: mmap"); abort(); }
  }
  if (dynSmem > (size_t)227 << 10) { fprintf(stderr, "cusim: %zu bytes of dynamic shared memory exceed the 227 KB of an SM\n", dynSmem); abort(); }
  //  Dynamic shared memory ends at an inaccessible page: a kernel that runs over the size it was launched with (rounded up
  //  to 128 bytes, the alignment the sources declare) stops here instead of scribbling over whatever a real SM has there.
  {
    static unsigned char *region = nullptr;               //  256 KB + one guard page, mapped once
    const size_t cap = (size_t)256 << 10;
    if (!region) {
      region = (unsigned char *)mmap(nullptr, cap + 4096, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
      if (region == MAP_FAILED) { perror("cusim: mmap"); abort(); }
      if (mprotect(region + cap, 4096, PROT_NONE) != 0) { perror("cusim: mprotect"); abort(); }
    }
    const size_t rounded = (dynSmem + 127) / 128 * 128;
    g_dynSmem = region + cap - rounded;
    g_dynCap  = rounded;
  }
  g_launches++;
  g_body = &body;
  static const bool trace = getenv("CUSIM_TRACE") != nullptr;
  struct timespec t0;  clock_gettime(CLOCK_MONOTONIC, &t0);
  g_blockDim = block;  g_gridDim = grid;
  const unsigned nWarps = (nThreads + 31) / 32;
  g_fibers.resize(nThreads);
  g_warps.resize(nWarps);

  for (unsigned bz = 0; bz < grid.z; bz++)
  for (unsigned by = 0; by < grid.y; by++)
  for (unsigned bx = 0; bx < grid.x; bx++) {
    g_blockIdx.x = bx;  g_blockIdx.y = by;  g_blockIdx.z = bz;
    if (g_dynSmem) memset(g_dynSmem, 0xcd, g_dynCap);   //  shared memory starts as garbage
    for (unsigned w = 0; w < nWarps; w++) { g_warps[w].alive = 0; memset(g_warps[w].bar, 0, sizeof(g_warps[w].bar)); }
    for (unsigned t = 0; t < nThreads; t++) {
      Fiber &f = g_fibers[t];
      f.linear = t;
      f.tid.x = t % block.x;  f.tid.y = (t / block.x) % block.y;  f.tid.z = t / (block.x * block.y);
      f.lane = t & 31;  f.warpId = t >> 5;  f.warp = &g_warps[t >> 5];
      f.warp->alive |= 1u << f.lane;
      f.done = false;
      prepare(f, t);
    }
    g_alive = nThreads;  g_barArrived = 0;  g_barGen = 0;
    unsigned long long lastProgress = g_progress, idleRounds = 0;
    while (g_alive) {
      for (unsigned t = 0; t < nThreads; t++) {
        Fiber &f = g_fibers[t];
        if (f.done) continue;
        g_cur = &f;
        cusim_switch(&g_schedSp, f.sp);
      }
      g_cur = nullptr;
      if (g_progress != lastProgress) { lastProgress = g_progress; idleRounds = 0; }
      else if (++idleRounds > 200000) {
        fprintf(stderr, "cusim: block (%u,%u,%u) makes no progress (%u threads alive): deadlock or a wait on a block that has not run\n", bx, by, bz, g_alive);
        abort();
      }
    }
    g_cur = nullptr;
  }
  g_body = nullptr;
  if (trace) {
    struct timespec t1;  clock_gettime(CLOCK_MONOTONIC, &t1);
    fprintf(stderr, "cusim: launch %llu grid (%u,%u,%u) block %u smem %zu: %.3f s\n", g_launches, grid.x, grid.y, grid.z, nThreads, dynSmem,
            (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec));
  }
}

//  ---- rank fibers ---------------------------------------------------------------------------------------------

namespace {
struct RankFiber { void *sp = nullptr; bool done = false; unsigned char *stack = nullptr; };
std::vector<RankFiber> g_ranks;
void *g_rankSchedSp = nullptr;
int   g_curRank = -1;
const std::function<void(unsigned)> *g_rankBody = nullptr;
const size_t kRankStack = (size_t)64 << 20;

void rank_main() {
  (*g_rankBody)((unsigned)g_curRank);
  RankFiber &r = g_ranks[g_curRank];
  r.done = true;
  cusim_switch(&r.sp, g_rankSchedSp);
  abort();
}
}  //  namespace

bool in_rank(void) { return g_curRank >= 0; }

void rank_yield(void) {
  if (g_curRank < 0) return;
  RankFiber &r = g_ranks[g_curRank];
  cusim_switch(&r.sp, g_rankSchedSp);
}

void run_ranks(unsigned world, const std::function<void(unsigned rank)> &body) {
  if (g_curRank >= 0) { fprintf(stderr, "cusim: nested run_ranks\n"); abort(); }
  g_ranks.assign(world, RankFiber());
  g_rankBody = &body;
  for (unsigned i = 0; i < world; i++) {
    RankFiber &r = g_ranks[i];
    r.stack = (unsigned char *)mmap(nullptr, kRankStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (r.stack == MAP_FAILED) { perror("cusim: mmap"); abort(); }
    void **sp = (void **)(r.stack + kRankStack);
    *--sp = nullptr;
    *--sp = (void *)rank_main;
    for (int k = 0; k < 6; k++) *--sp = nullptr;
    r.sp = sp;
  }
  unsigned live = world;
  unsigned long long rounds = 0;
  while (live) {
    for (unsigned i = 0; i < world; i++) {
      if (g_ranks[i].done) continue;
      g_curRank = (int)i;
      cusim_switch(&g_rankSchedSp, g_ranks[i].sp);
      g_curRank = -1;
      if (g_ranks[i].done) live--;
    }
    if (++rounds > 50000000ull) { fprintf(stderr, "cusim: ranks make no progress (a collective some rank never joins?)\n"); abort(); }
  }
  for (unsigned i = 0; i < world; i++) munmap(g_ranks[i].stack, kRankStack);
  g_ranks.clear();
  g_rankBody = nullptr;
}

}  //  namespace cusim
