//  Multi-GPU: routing extracted k-mers to the GPU that owns their data file.
//
//  The reference has no distributed mode; the shard unit is its own output partition: the 64 data
//  files (top 6 bits of the k-mer, `merylFileWriter::fileNumber`,
//  src/utility/src/kmers-v1/kmers-writer.C:157-178).  Every file's blocks, index and bytes depend only
//  on the k-mers with that prefix, so after ONE all-to-all of packed k-mers per pass each GPU sorts,
//  counts and encodes its own files with no further communication.
//
//  Ownership: canonical k-mers are linearly skewed across files (file 0 holds ~2x the mean, file 63
//  almost nothing), so file f is paired with 63-f and pairs are dealt round-robin:
//      owner(f) = min(f, 63-f) mod worldSize.
//  With that map the j-th owned file of every rank together form the contiguous file range
//  [j*world, (j+1)*world), so a pass over "slots [a,b) of every rank" is a contiguous prefix range
//  for the extraction kernel.
//
//  NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy torch already loaded when the
//  caller is bench.py), so the library has no link-time dependency on it; the communicator is
//  bootstrapped from a 128-byte unique id that the caller broadcasts with whatever it has
//  (torch.distributed in bench.py).  One process per GPU.

#include <dlfcn.h>
#include <string.h>
#include <vector>

#include "common.h"
#include "engine.h"
#include "exchange.h"

namespace {

typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void *ncclComm_p;
enum { kNcclUint8 = 1, kNcclUint64 = 5 };                 //  ncclDataType_t values (nccl.h)

struct NcclApi {
  void *handle = nullptr;
  int  (*GetUniqueId)(ncclUniqueId_t *) = nullptr;
  int  (*CommInitRank)(ncclComm_p *, int, ncclUniqueId_t, int) = nullptr;
  int  (*CommDestroy)(ncclComm_p) = nullptr;
  int  (*GroupStart)(void) = nullptr;
  int  (*GroupEnd)(void) = nullptr;
  int  (*Send)(const void *, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
  int  (*Recv)(void *, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
  int  (*AllGather)(const void *, void *, size_t, int, ncclComm_p, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};

NcclApi g_nccl;

int loadNccl(void) {
  if (g_nccl.handle) return MB_OK;
  const char *names[] = { "libnccl.so.2", "libnccl.so", nullptr };
  void *h = nullptr;
  for (int i = 0; names[i] && !h; i++) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!h) return mb_fail(MB_ECOMM, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(field, name) do { *(void **)&g_nccl.field = dlsym(h, name); if (!g_nccl.field) return mb_fail(MB_ECOMM, "libnccl lacks %s", name); } while (0)
  SYM(GetUniqueId, "ncclGetUniqueId");    SYM(CommInitRank, "ncclCommInitRank");   SYM(CommDestroy, "ncclCommDestroy");
  SYM(GroupStart, "ncclGroupStart");      SYM(GroupEnd, "ncclGroupEnd");
  SYM(Send, "ncclSend");                  SYM(Recv, "ncclRecv");                   SYM(AllGather, "ncclAllGather");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.handle = h;
  return MB_OK;
}

#define MB_NCCL(call) do { int r_ = (call); if (r_ != 0) return mb_fail(MB_ECOMM, "%s:%d: %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); } while (0)

}  //  namespace


struct MbComm {
  ncclComm_p comm = nullptr;
  u32 rank = 0, world = 1;
  std::vector<unsigned char> peerHandle;                     //  world x 64 bytes: the CUDA-IPC handle behind peerPtr[r]
  std::vector<void *>        peerPtr;                        //  rank r's arena in this process (nullptr: not mapped)
};

#ifdef MB_CUSIM
//  Host emulator (tests/cusim): the ranks are fibers of one process, "device" pointers are host pointers, so a
//  collective is a rendezvous: every rank deposits its arguments, the last one to arrive moves the data for all.
namespace {
struct SimHub {
  u32 world = 0, arrived = 0;  u64 gen = 0;
  struct Args { const void *a = nullptr; void *b = nullptr; const u64 *c = nullptr; const u64 *d = nullptr; u64 n = 0; u32 x = 0, y = 0, z = 0; } args[32];
} g_hub;

template <typename F> int simCollective(MbComm *m, const SimHub::Args &mine, F moveAll) {
  g_hub.args[m->rank] = mine;
  const u64 gen = g_hub.gen;
  if (++g_hub.arrived == m->world) { moveAll(); g_hub.arrived = 0; g_hub.gen++; }
  else while (g_hub.gen == gen) cusim::rank_yield();
  return MB_OK;
}
}  //  namespace

extern "C" int mb_sim_comm_attach(mb_ctx *ctx) {             //  instead of mb_comm_unique_id + mb_comm_init
  const mb_config &cfg = mb_ctx_config(ctx);
  MbComm *m = new MbComm();
  m->rank = cfg.rank;  m->world = cfg.worldSize;
  g_hub.world = cfg.worldSize;
  mb_ctx_set_comm(ctx, m);
  return MB_OK;
}
#endif

//  The communicator is created once per process and shared by every context: NCCL sets up its
//  peer-to-peer channels lazily (seconds for the first send/recv), which must not be paid per job.
static MbComm *g_processComm = nullptr;

void mb_comm_destroy(MbComm *m) {
  if (!m || m == g_processComm) return;                      //  the process-wide communicator lives until exit
#ifndef MB_CUSIM
  for (void *p : m->peerPtr) if (p) cudaIpcCloseMemHandle(p);
#endif
  if (m->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(m->comm);
  delete m;
}

u32 mb_comm_rank(const MbComm *m)  { return m ? m->rank : 0; }
u32 mb_comm_world(const MbComm *m) { return m ? m->world : 1; }

u32 mb_file_owner(u32 file, u32 world) {
  const u32 p = file < 63 - file ? file : 63 - file;
  return p % world;
}

extern "C" uint32_t mb_file_owner_of(uint32_t file, uint32_t worldSize) {
  return worldSize ? mb_file_owner(file & 63, worldSize) : 0;
}

extern "C" int mb_comm_unique_id(uint8_t id[128]) {
  mb_clear_error();
  if (!id) return mb_fail(MB_EINVAL, "mb_comm_unique_id: NULL argument");
  MB_TRY(loadNccl());
  ncclUniqueId_t u;
  MB_NCCL(g_nccl.GetUniqueId(&u));
  memcpy(id, u.internal, 128);
  return MB_OK;
}

extern "C" int mb_comm_init(mb_ctx *ctx, const uint8_t id[128]) {
  mb_clear_error();
  if (!ctx || !id) return mb_fail(MB_EINVAL, "mb_comm_init: NULL argument");
  const mb_config &cfg = mb_ctx_config(ctx);
  if (cfg.worldSize < 2) return mb_fail(MB_EINVAL, "mb_comm_init: context was created with worldSize %u", cfg.worldSize);
  if (64 % cfg.worldSize != 0 || cfg.worldSize > 32) return mb_fail(MB_EINVAL, "mb_comm_init: worldSize %u must divide 64 and be <= 32", cfg.worldSize);
  if (cfg.rank >= cfg.worldSize) return mb_fail(MB_EINVAL, "mb_comm_init: rank %u >= worldSize %u", cfg.rank, cfg.worldSize);
  MB_TRY(loadNccl());
  MB_CUDA(cudaSetDevice(mb_ctx_device(ctx)));
  if (g_processComm) {
    if (g_processComm->rank != cfg.rank || g_processComm->world != cfg.worldSize)
      return mb_fail(MB_ECOMM, "mb_comm_init: this process already joined a communicator as rank %u of %u", g_processComm->rank, g_processComm->world);
    mb_ctx_set_comm(ctx, g_processComm);
    return MB_OK;
  }
  ncclUniqueId_t u;
  memcpy(u.internal, id, 128);
  MbComm *m = new MbComm();
  m->rank = cfg.rank;  m->world = cfg.worldSize;
  const int r = g_nccl.CommInitRank(&m->comm, (int)cfg.worldSize, u, (int)cfg.rank);
  if (r != 0) { delete m; return mb_fail(MB_ECOMM, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); }

  //  Warm-up: one word to and from every peer, so the channels exist before the first timed exchange.
  u64 *d = nullptr;
  MB_CUDA(cudaMalloc((void **)&d, 2 * sizeof(u64) * cfg.worldSize));
  MB_CUDA(cudaMemset(d, 0, 2 * sizeof(u64) * cfg.worldSize));
  cudaStream_t s = mb_ctx_stream(ctx);
  MB_NCCL(g_nccl.GroupStart());
  for (u32 p = 0; p < cfg.worldSize; p++) {
    MB_NCCL(g_nccl.Send(d + p, 1, kNcclUint64, (int)p, m->comm, s));
    MB_NCCL(g_nccl.Recv(d + cfg.worldSize + p, 1, kNcclUint64, (int)p, m->comm, s));
  }
  MB_NCCL(g_nccl.GroupEnd());
  MB_CUDA(cudaStreamSynchronize(s));
  cudaFree(d);

  g_processComm = m;
  mb_ctx_set_comm(ctx, m);
  return MB_OK;
}

//  All-gather `n` u64 per rank (device pointers; dAll holds world*n).
int mb_comm_allgather_u64(MbComm *m, const u64 *dMine, u64 *dAll, size_t n, cudaStream_t s) {
#ifdef MB_CUSIM
  SimHub::Args a;  a.a = dMine;  a.b = dAll;  a.n = n;
  return simCollective(m, a, [&]() {
    for (u32 dst = 0; dst < m->world; dst++)
      for (u32 src = 0; src < m->world; src++)
        memcpy((u64 *)g_hub.args[dst].b + (size_t)src * n, g_hub.args[src].a, n * sizeof(u64));
  });
#endif
  MB_NCCL(g_nccl.AllGather(dMine, dAll, n, kNcclUint64, m->comm, s));
  return MB_OK;
}

//  The per-pass all-to-all.  `files` are the contiguous files [fileLo, fileHi) of this pass.
//  sendBuf holds this rank's extracted keys grouped by file (sendStart[f - fileLo] in keys, counts
//  C[me][f]); file f goes to owner(f).  recvBuf receives, for every file this rank owns (ascending),
//  the chunks of ranks 0..world-1 back to back, so each owned file ends up as one contiguous segment.
//  C is the [world][64] matrix of per-rank file counts (host).  keyWords u64 per key.
int mb_comm_exchange_keys(MbComm *m, u32 keyWords, const void *sendBuf, void *recvBuf,
                          u32 fileLo, u32 fileHi, const u64 *sendStart, const u64 *C, cudaStream_t s) {
  const u32 me = m->rank, world = m->world;
  const u64 *sb = (const u64 *)sendBuf;
  u64       *rb = (u64 *)recvBuf;
#ifdef MB_CUSIM
  {
    SimHub::Args a;  a.a = sendBuf;  a.b = recvBuf;  a.c = sendStart;  a.d = C;  a.x = fileLo;  a.y = fileHi;  a.z = keyWords;
    return simCollective(m, a, [&]() {
      //  what the grouped ncclSend / ncclRecv pairs below do, for every (file, source, owner)
      for (u32 dst = 0; dst < world; dst++) {
        u64 recvAt = 0;
        for (u32 f = fileLo; f < fileHi; f++) {
          if (mb_file_owner(f, world) != dst) continue;
          for (u32 src = 0; src < world; src++) {
            const u64 c = C[(size_t)src * 64 + f];
            if (c) memcpy((u64 *)g_hub.args[dst].b + recvAt * keyWords,
                          (const u64 *)g_hub.args[src].a + g_hub.args[src].c[f - fileLo] * keyWords, c * keyWords * sizeof(u64));
            recvAt += c;
          }
        }
      }
    });
  }
#endif
  MB_NCCL(g_nccl.GroupStart());
  u64 recvAt = 0;
  for (u32 f = fileLo; f < fileHi; f++) {
    const u32 own = mb_file_owner(f, world);
    const u64 mine = C[(size_t)me * 64 + f];
    if (mine) MB_NCCL(g_nccl.Send(sb + sendStart[f - fileLo] * keyWords, mine * keyWords, kNcclUint64, (int)own, m->comm, s));
    if (own == me) {
      for (u32 src = 0; src < world; src++) {
        const u64 c = C[(size_t)src * 64 + f];
        if (c) MB_NCCL(g_nccl.Recv(rb + recvAt * keyWords, c * keyWords, kNcclUint64, (int)src, m->comm, s));
        recvAt += c;
      }
    }
  }
  MB_NCCL(g_nccl.GroupEnd());
  return MB_OK;
}

//  The same all-to-all for u32 suffix keys (k <= 21: the bits below the prefix bin fit one word, the bin is implied by
//  the segment a key sits in).  The unit is the prefix BIN: bins [binLo, binHi) of this pass, bin b belongs to the owner
//  of its data file b / binsPerFile.  sendBuf holds this rank's suffixes grouped by bin (sendStart[b - binLo] in keys,
//  counts Cb[me][b]); recvBuf receives, for every bin of every file this rank owns (ascending), the chunks of ranks
//  0..world-1 back to back, so every owned bin ends up as one contiguous sort segment.  Cb is the [world][nBins] matrix
//  of per-rank bin counts (host).  One message per (bin, peer): 4 bytes per k-mer on the wire instead of 8.
int mb_comm_exchange_suffixes(MbComm *m, const void *sendBuf, void *recvBuf, u32 binLo, u32 binHi, u32 binsPerFile,
                              const u64 *sendStart, const u64 *Cb, u32 nBins, cudaStream_t s) {
  const u32 me = m->rank, world = m->world;
  const u32 *sb = (const u32 *)sendBuf;
  u32       *rb = (u32 *)recvBuf;
#ifdef MB_CUSIM
  {
    SimHub::Args a;  a.a = sendBuf;  a.b = recvBuf;  a.c = sendStart;  a.d = Cb;  a.x = binLo;  a.y = binHi;  a.z = binsPerFile;  a.n = nBins;
    return simCollective(m, a, [&]() {
      for (u32 dst = 0; dst < world; dst++) {
        u64 recvAt = 0;
        for (u32 b = binLo; b < binHi; b++) {
          if (mb_file_owner(b / binsPerFile, world) != dst) continue;
          for (u32 src = 0; src < world; src++) {
            const u64 c = Cb[(size_t)src * nBins + b];
            if (c) memcpy((u32 *)g_hub.args[dst].b + recvAt, (const u32 *)g_hub.args[src].a + g_hub.args[src].c[b - binLo], c * sizeof(u32));
            recvAt += c;
          }
        }
      }
    });
  }
#endif
  //  One group per data file (every rank walks the files in the same order): a group holds at most binsPerFile sends
  //  and binsPerFile * world receives, instead of thousands of operations in one group.
  u64 recvAt = 0;
  for (u32 f = binLo / binsPerFile; f < binHi / binsPerFile; f++) {
    const u32 own = mb_file_owner(f, world);
    MB_NCCL(g_nccl.GroupStart());
    for (u32 b = f * binsPerFile; b < (f + 1) * binsPerFile; b++) {
      const u64 mine = Cb[(size_t)me * nBins + b];
      if (mine) MB_NCCL(g_nccl.Send(sb + sendStart[b - binLo], mine * sizeof(u32), kNcclUint8, (int)own, m->comm, s));
      if (own == me) {
        for (u32 src = 0; src < world; src++) {
          const u64 c = Cb[(size_t)src * nBins + b];
          if (c) MB_NCCL(g_nccl.Recv(rb + recvAt, c * sizeof(u32), kNcclUint8, (int)src, m->comm, s));
          recvAt += c;
        }
      }
    }
    MB_NCCL(g_nccl.GroupEnd());
  }
  return MB_OK;
}

//  ---- extraction fused with the exchange: peer memory over NVLink -------------------------------------------------
//  Every rank's arena is mapped into every other rank's address space (CUDA IPC; one process per GPU), so the
//  extraction kernel of a pass can store a kept k-mer straight into the receive segment of the GPU that owns its file:
//  the k-mers cross NVLink tile by tile while the extraction runs, no send buffer and no separate all-to-all.
//  dScratch: device memory for (world + 1) * 8 u64.  peerBase[r] = rank r's arena in this process (own arena for r == me).
int mb_comm_open_peer_arenas(MbComm *m, void *myArena, u64 *dScratch, uint8_t **peerBase, cudaStream_t s) {
  const u32 me = m->rank, world = m->world;
#ifdef MB_CUSIM
  {
    SimHub::Args a;  a.a = myArena;  a.b = (void *)peerBase;
    return simCollective(m, a, [&]() {
      for (u32 dst = 0; dst < world; dst++)
        for (u32 src = 0; src < world; src++) ((uint8_t **)g_hub.args[dst].b)[src] = (uint8_t *)g_hub.args[src].a;
    });
  }
#else
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t mine;
  MB_CUDA(cudaIpcGetMemHandle(&mine, myArena));
  std::vector<cudaIpcMemHandle_t> all(world);
  MB_CUDA(cudaMemcpyAsync(dScratch, &mine, 64, cudaMemcpyHostToDevice, s));
  MB_TRY(mb_comm_allgather_u64(m, dScratch, dScratch + 8, 8, s));
  MB_CUDA(cudaMemcpyAsync(all.data(), dScratch + 8, (size_t)world * 64, cudaMemcpyDeviceToHost, s));
  MB_CUDA(cudaStreamSynchronize(s));
  //  A mapping outlives the job (measured: opening and closing a 170 GB peer arena costs ~1.8 s, per peer and job): the
  //  ranks keep their arenas between contexts (engine.cu), so the same handle comes back and the mapping is reused; a
  //  different handle (the peer allocated a new arena) replaces it.
  if (m->peerHandle.size() != (size_t)world * 64) { m->peerHandle.assign((size_t)world * 64, 0); m->peerPtr.assign(world, nullptr); }
  for (u32 r = 0; r < world; r++) peerBase[r] = nullptr;
  for (u32 r = 0; r < world; r++) {
    if (r == me) { peerBase[r] = (uint8_t *)myArena; continue; }
    if (m->peerPtr[r] && memcmp(&m->peerHandle[(size_t)r * 64], &all[r], 64) == 0) { peerBase[r] = (uint8_t *)m->peerPtr[r]; continue; }
    if (m->peerPtr[r]) { cudaIpcCloseMemHandle(m->peerPtr[r]); m->peerPtr[r] = nullptr; }
    void *p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return mb_fail(MB_ECOMM, "cannot map the arena of rank %u into rank %u: %s", r, me, cudaGetErrorString(e));
    }
    m->peerPtr[r] = p;
    memcpy(&m->peerHandle[(size_t)r * 64], &all[r], 64);
    peerBase[r] = (uint8_t *)p;
  }
  return MB_OK;
#endif
}

//  Unmap the peers' arenas (call after a collective that follows the last use, and before any rank frees its arena).
int mb_comm_close_peer_arenas(MbComm *m, uint8_t **peerBase) {
  //  the mappings stay with the communicator (see above); a communicator that is destroyed closes them
  for (u32 r = 0; r < m->world; r++) peerBase[r] = nullptr;
  return MB_OK;
}

//  Gather variable-length u64 arrays on rank 0 (device staging provided by the caller):
//  every rank sends `n` words from dSend; rank 0 receives rank r's words at dRecv + offsets[r].
int mb_comm_gather_u64(MbComm *m, const u64 *dSend, u64 n, u64 *dRecv, const u64 *counts, cudaStream_t s) {
#ifdef MB_CUSIM
  {
    SimHub::Args a;  a.a = dSend;  a.b = dRecv;  a.n = n;
    return simCollective(m, a, [&]() {
      u64 at = 0;
      for (u32 r = 1; r < m->world; r++) {
        if (g_hub.args[r].n) memcpy((u64 *)g_hub.args[0].b + at, g_hub.args[r].a, g_hub.args[r].n * sizeof(u64));
        at += g_hub.args[r].n;
      }
    });
  }
#endif
  MB_NCCL(g_nccl.GroupStart());
  if (m->rank != 0) {
    if (n) MB_NCCL(g_nccl.Send(dSend, n, kNcclUint64, 0, m->comm, s));
  } else {
    u64 at = 0;
    for (u32 r = 1; r < m->world; r++) {
      if (counts[r]) MB_NCCL(g_nccl.Recv(dRecv + at, counts[r], kNcclUint64, (int)r, m->comm, s));
      at += counts[r];
    }
  }
  MB_NCCL(g_nccl.GroupEnd());
  return MB_OK;
}
