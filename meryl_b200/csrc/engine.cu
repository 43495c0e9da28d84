//  Host orchestration of the device-resident `meryl count` (the role of
//  merylOperation::countThreads, src/meryl/merylOp-countThreads.C:385-474, and of the
//  merylBlockWriter it drives).
//
//  One context == one GPU.  All device memory is one arena, carved by the host:
//
//    [ packed read store | zero pad ][ control block ][ value histogram ] ... pass buffers ...  [ ingest staging x2 ]
//
//  Ingest (mb_add_*): the base stream is collected in pinned host buffers, copied to the device and
//  packed (K1) into the store, asynchronously, while the host keeps parsing.
//  Count (mb_count):
//    1. K2a bin histogram over the store -> exact number of k-mers per prefix bin (top 10 bits);
//    2. passes: consecutive bin ranges whose k-mers fit the key buffers.  A prefix is never split
//       between passes, so no merge step exists and counts equal the reference's single batch;
//    3. per pass: K2/K3 extract+scatter -> K4 segmented radix sort -> K5 run-length count ->
//       value histogram -> K6 encode -> device-to-host copy of the encoded blocks and index entries.
//  The host only appends the encoded bytes to the 64 file images and builds the master index.

#include <algorithm>
#include <map>
#include <mutex>
#include <memory>
#include <string>
#include <thread>
#include <stdio.h>
#include <string.h>
#include <unistd.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>

#include "common.h"
#include "kernels.cuh"
#include "engine.h"
#include "exchange.h"
#include "filesink.h"


namespace {

const size_t kAlign      = 256;
const size_t kStageBytes = (size_t)64 << 20;        //  ingest staging buffer (multiple of 32)

inline size_t alignUp(size_t x, size_t a = kAlign) { return (x + a - 1) / a * a; }

struct EventTimer {                                  //  device time of a stream section (events released on every way out)
  cudaEvent_t a = nullptr, b = nullptr;
  EventTimer() {}
  EventTimer(const EventTimer &) = delete;
  EventTimer &operator=(const EventTimer &) = delete;
  ~EventTimer()                       { destroy(); }
  void init()                         { cudaEventCreate(&a); cudaEventCreate(&b); }
  void destroy()                      { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); a = b = nullptr; }
  void start(cudaStream_t s)          { cudaEventRecord(a, s); }
  void stop (cudaStream_t s)          { cudaEventRecord(b, s); }
  float ms()                          { float m = 0; cudaEventSynchronize(b); cudaEventElapsedTime(&m, a, b); return m; }
};

}  //  namespace


struct mb_ctx {
  mb_config   cfg;
  int         device     = 0;
  u32         keyWords   = 1;
  u32         binBits    = 10;
  cudaStream_t stream    = nullptr;

  uint8_t    *arena      = nullptr;
  size_t      arenaBytes = 0;

  //  store
  u64         nRecords   = 0;                        //  records filled so far
  size_t      storeLimit = 0;                        //  bytes the store may grow to during ingest

  //  ingest staging
  char       *hStage[2]  = { nullptr, nullptr };
  uint8_t    *dStage[2]  = { nullptr, nullptr };
  cudaEvent_t evH2D[2]   = { nullptr, nullptr };     //  hStage[i] has been copied
  cudaEvent_t evIngest[2]= { nullptr, nullptr };     //  first chunk queued / last chunk packed (st.msPack)
  bool        ingestStarted = false;
  size_t      fill       = 0;
  int         cur        = 0;
  bool        inRecord   = false;                    //  bases since the last '.'
  uint8_t     hpLast     = 0;                        //  MB_FLAG_COMPRESS: last letter kept in this record (lower case), 0 at a record start
  std::vector<char> hpBuf;

  //  results
  bool        counted    = false;
  MbDbImages  db;
  MbHistogram hist;
  mb_stats    st;
  std::vector<u64> cLo, cHi;                         //  MB_FLAG_KEEP_COUNTED
  std::vector<u32> cCnt;
  std::vector<u64> sortedLo, sortedHi;               //  debug: all k-mers, sorted (stopAfterSort)
  bool        stopAfterSort = false;

  uint8_t    *outBuf     = nullptr;                  //  caller-supplied host buffer for the data images
  size_t      outCap     = 0, outUsed = 0;
  std::string outDir;                                //  mb_set_output_dir: the data files are streamed there while counting

  MbComm     *comm       = nullptr;                  //  multi-GPU (exchange.cu)

  //  Count batches (the reference's finishBatch / mergeBatches, kmers-writer-block.C:172-386): when the packed reads
  //  reach batchLimit records the store is counted at the next record break, its encoded images are set aside in host
  //  memory and the store starts again; mb_count merges the batches (union.cu pipeline, saturating sums).
  std::vector<std::unique_ptr<MbDbImages>> batches;
  u64         batchLimit = 0;                        //  records; 0 = never
  bool        closing    = false;
  //  several GPUs: every rank takes part in every batch (runCount is collective).  A round's first all-gather carries two flags:
  //  "I count this as a batch" and "nothing follows this round on my side" (mb_count); the answers:
  bool        finalRound = false;                    //  set by mb_count
  bool        roundAnyBatch = false, roundAllFinal = true;
};


//  ------------------------------------------------------------------------------------------
//  creation / destruction
//  ------------------------------------------------------------------------------------------

extern "C" int mb_create(const mb_config *cfg, mb_ctx **out) {
  mb_clear_error();
  if (!cfg || !out)                       return mb_fail(MB_EINVAL, "mb_create: NULL argument");
  if (cfg->k < 3 || cfg->k > 64)          return mb_fail(MB_EINVAL, "mb_create: k=%u outside 3..64", cfg->k);
  if (cfg->mode > MB_MODE_REVERSE)        return mb_fail(MB_EINVAL, "mb_create: bad mode %u", cfg->mode);
  if (cfg->wPrefix < 6 || cfg->wPrefix > 2 * cfg->k || cfg->wPrefix > 40)
    return mb_fail(MB_EINVAL, "mb_create: wPrefix=%u unusable for k=%u (need 6..min(2k,40))", cfg->wPrefix, cfg->k);

  int nDev = 0;
  if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0)
    return mb_fail(MB_ENODEVICE, "mb_create: no CUDA device (this library has no CPU path)");

  std::unique_ptr<mb_ctx, void (*)(mb_ctx *)> guard(new mb_ctx(), mb_destroy);   //  released on success, destroyed on any early return
  mb_ctx *c = guard.get();
  c->cfg = *cfg;
  if (cfg->device >= 0) { c->device = cfg->device; MB_CUDA(cudaSetDevice(c->device)); }
  else                  { MB_CUDA(cudaGetDevice(&c->device)); }

  cudaDeviceProp prop;
  MB_CUDA(cudaGetDeviceProperties(&prop, c->device));
  if (prop.major < 10) {
    return mb_fail(MB_ENODEVICE, "mb_create: device %d is sm_%d%d; this build is sm_100a only", c->device, prop.major, prop.minor);
  }

  c->keyWords = cfg->k <= 32 ? 1 : 2;
  c->binBits  = mb_bin_bits(cfg->wPrefix);
  memset(&c->st, 0, sizeof(c->st));
  c->st.keyBytes = 8 * c->keyWords;
  for (u32 f = 0; f < 64; f++) c->db.owned[f] = true;

  MB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; i++) {
    MB_CUDA(cudaEventCreateWithFlags(&c->evH2D[i], cudaEventDisableTiming));
    MB_CUDA(cudaEventCreate(&c->evIngest[i]));
  }
  //  The device arena is taken at the first use (ensureArena), i.e. AFTER mb_comm_init in a multi-GPU
  //  job: NCCL allocates its own channel buffers (gigabytes with 8 peers) and must find room.

  *out = guard.release();
  return MB_OK;
}


//  A process keeps ONE arena per device between contexts (and its two page-locked staging buffers): a service that counts job
//  after job allocates its ~170 GB once, and -- what made this necessary -- the peers' CUDA-IPC mappings of an arena stay valid
//  from job to job (measured on two B200s: mapping and unmapping a 170 GB peer arena costs ~1.8 s per job, as much as the
//  whole count).  mb_destroy hands the arena back here; mb_release_device_memory() (or MB_NO_ARENA_CACHE=1) frees it; the
//  other entry points that size their own arenas from the free memory (set operations, lookup) release it first.
namespace {
struct ArenaSlot { uint8_t *arena = nullptr; size_t bytes = 0; char *hStage[2] = { nullptr, nullptr }; };
std::mutex g_arenaMu;
ArenaSlot  g_arenaSlot[64];
bool arenaCacheOn(void) { return getenv("MB_NO_ARENA_CACHE") == nullptr; }
void freeSlot(ArenaSlot &a) {
  if (a.arena) cudaFree(a.arena);
  for (int i = 0; i < 2; i++) if (a.hStage[i]) cudaFreeHost(a.hStage[i]);
  a = ArenaSlot();
}
}  //  namespace

void mb_arena_cache_release(int device) {                 //  device < 0: all
  std::lock_guard<std::mutex> lk(g_arenaMu);
  int cur = 0;  cudaGetDevice(&cur);
  for (int d = 0; d < 64; d++)
    if ((device < 0 || d == device) && g_arenaSlot[d].arena) { cudaSetDevice(d); freeSlot(g_arenaSlot[d]); }
  cudaSetDevice(cur);
}

extern "C" void mb_release_device_memory(void) { mb_arena_cache_release(-1); }

extern "C" void mb_destroy(mb_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->comm) mb_comm_destroy(c->comm);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (int i = 0; i < 2; i++) {
    if (c->evH2D[i])  cudaEventDestroy(c->evH2D[i]);
    if (c->evIngest[i]) cudaEventDestroy(c->evIngest[i]);
  }
  {
    ArenaSlot mine;
    mine.arena = c->arena;  mine.bytes = c->arenaBytes;  mine.hStage[0] = c->hStage[0];  mine.hStage[1] = c->hStage[1];
    std::lock_guard<std::mutex> lk(g_arenaMu);
    if (c->arena && arenaCacheOn() && c->device >= 0 && c->device < 64 && cudaPeekAtLastError() == cudaSuccess) {
      ArenaSlot &a = g_arenaSlot[c->device];
      if (a.arena && a.bytes >= mine.bytes) freeSlot(mine);                //  keep the larger one
      else { freeSlot(a); a = mine; }
    } else freeSlot(mine);
  }
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}


//  ------------------------------------------------------------------------------------------
//  ingest
//  ------------------------------------------------------------------------------------------

static int ensureArena(mb_ctx *c) {
  if (c->arena) return MB_OK;
  MB_CUDA(cudaSetDevice(c->device));
  if (c->device >= 0 && c->device < 64) {                  //  the arena the previous context on this device left behind
    std::lock_guard<std::mutex> lk(g_arenaMu);
    ArenaSlot &a = g_arenaSlot[c->device];
    if (a.arena) {
      const size_t wantB = c->cfg.workspaceBytes / kAlign * kAlign;
      if (arenaCacheOn() && (c->cfg.workspaceBytes == 0 ? a.bytes >= ((size_t)8 << 30) : a.bytes == wantB)) {
        c->arena = a.arena;  c->arenaBytes = a.bytes;
        for (int i = 0; i < 2; i++) c->hStage[i] = a.hStage[i];
        a = ArenaSlot();
      } else freeSlot(a);
    }
  }
  size_t freeB = 0, totalB = 0;
  MB_CUDA(cudaMemGetInfo(&freeB, &totalB));
  const size_t margin = (size_t)(c->cfg.worldSize > 1 ? 4 : 2) << 30;
  size_t want = c->cfg.workspaceBytes;
  if (want == 0) want = freeB > margin + ((size_t)1 << 30) ? freeB - margin : freeB / 2;
  if (want > freeB) want = freeB;
  want = want / kAlign * kAlign;
  if (!c->arena) {
    if (want < ((size_t)256 << 20))
      return c->cfg.workspaceBytes && c->cfg.workspaceBytes <= freeB
               ? mb_fail(MB_EINVAL, "workspaceBytes = %llu: the device arena needs at least 256 MB", (unsigned long long)c->cfg.workspaceBytes)
               : mb_fail(MB_ENOMEM, "only %zu bytes of device memory free", freeB);
    if (cudaMalloc((void **)&c->arena, want) != cudaSuccess) {
      cudaGetLastError();
      c->arena = nullptr;
      return mb_fail(MB_ENOMEM, "cannot allocate a %zu byte device arena", want);
    }
    c->arenaBytes = want;
  }
  for (int i = 0; i < 2; i++) {
    if (!c->hStage[i]) MB_CUDA(cudaMallocHost((void **)&c->hStage[i], kStageBytes));
    c->dStage[i] = c->arena + c->arenaBytes - (size_t)(i + 1) * kStageBytes;
  }
  c->storeLimit = c->arenaBytes - 2 * kStageBytes - ((size_t)64 << 20);
  //  One batch = as many reads as leave half of the arena to the key buffers of a pass.  (With several GPUs a batch is closed
  //  by all ranks together, see mb_count.)  Not for the kernel-only / debugging forms that keep no images.
  if (!(c->cfg.flags & (MB_FLAG_NO_IMAGES | MB_FLAG_KEEP_COUNTED)) && !c->stopAfterSort) {
    c->batchLimit = c->storeLimit / 2 / (MB_STORE_REC_WORDS * 4);
    if (getenv("MB_MAX_BATCH_BASES")) c->batchLimit = std::max<u64>(1, strtoull(getenv("MB_MAX_BATCH_BASES"), nullptr, 10) / 32);   //  testing knob
  }
  return MB_OK;
}

static int storeReserve(mb_ctx *c, u64 moreRecords) {
  const size_t need = (size_t)(c->nRecords + moreRecords + MB_STORE_PAD_RECORDS) * MB_STORE_REC_WORDS * 4;
  if (need > c->storeLimit)
    return mb_fail(MB_ENOMEM, "input too large for one device: the packed read store needs %zu bytes, %zu available "
                              "(shard the input over more GPUs)", need, c->storeLimit);
  return MB_OK;
}

//  Ship the filled part of the current pinned buffer: H2D, then K1 into the store.
static int flushStage(mb_ctx *c) {
  if (c->fill == 0) return MB_OK;
  const int    i = c->cur;
  const size_t n = c->fill;
  const u64 recs = (n + 31) / 32;
  MB_TRY(storeReserve(c, recs));
  MB_CUDA(cudaSetDevice(c->device));
  if (!c->ingestStarted) { c->ingestStarted = true; MB_CUDA(cudaEventRecord(c->evIngest[0], c->stream)); }
  MB_CUDA(cudaMemcpyAsync(c->dStage[i], c->hStage[i], n, cudaMemcpyHostToDevice, c->stream));
  MB_CUDA(cudaEventRecord(c->evH2D[i], c->stream));
  MB_CUDA(mb_launch_pack(c->dStage[i], n, (u32 *)c->arena, c->nRecords, c->stream));
  c->st.nLaunches++;
  c->st.h2dBytes += n;
  c->nRecords    += recs;
  c->fill = 0;
  c->cur ^= 1;
  MB_CUDA(cudaEventSynchronize(c->evH2D[c->cur]));      //  the other pinned buffer must have left the host
  return MB_OK;
}

//  The copy into the page-locked staging buffer was the serial part of the file ingest (one thread behind sixteen parsing
//  threads): large pieces are copied by several threads.
static void stageCopy(char *dst, const char *src, size_t n) {
  static const unsigned hw = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
  if (n < ((size_t)8 << 20) || hw < 2) { memcpy(dst, src, n); return; }
  const size_t per = (n / hw + 4095) / 4096 * 4096;
  std::vector<std::thread> th;
  for (size_t at = per; at < n; at += per) th.emplace_back([=]() { memcpy(dst + at, src + at, std::min(per, n - at)); });
  memcpy(dst, src, std::min(per, n));
  for (auto &t : th) t.join();
}

static int appendRaw(mb_ctx *c, const char *p, size_t n) {
  MB_TRY(ensureArena(c));
  while (n) {
    size_t room = kStageBytes - c->fill;
    if (room == 0) { MB_TRY(flushStage(c)); room = kStageBytes; }
    const size_t m = n < room ? n : room;
    stageCopy(c->hStage[c->cur] + c->fill, p, m);
    c->fill += m;  p += m;  n -= m;
  }
  return MB_OK;
}

static int closeBatch(mb_ctx *c);

static inline bool batchDue(const mb_ctx *c) {
  return c->batchLimit && !c->closing && c->nRecords + c->fill / 32 >= c->batchLimit;
}

//  The ingest's append: once a batch is due, the stream is cut after the next record break ('.': no k-mer spans it).
static int appendHost(mb_ctx *c, const char *p, size_t n) {
  MB_TRY(ensureArena(c));
  while (n) {
    size_t m = n;
    if (batchDue(c)) {
      const char *dot = (const char *)memchr(p, '.', n);
      m = dot ? (size_t)(dot - p) + 1 : n;
      MB_TRY(appendRaw(c, p, m));
      if (dot) MB_TRY(closeBatch(c));
    } else {
      if (c->batchLimit && !c->closing) {                  //  no further than the point where the batch becomes due
        const u64 room = (c->batchLimit - (c->nRecords + c->fill / 32)) * 32;
        if (m > room) m = (size_t)room;
      }
      MB_TRY(appendRaw(c, p, m));
    }
    p += m;  n -= m;
  }
  return MB_OK;
}

static int parserSink(void *arg, const char *bases, size_t n) {
  mb_ctx *c = (mb_ctx *)arg;
  if (n) c->inRecord = (bases[n - 1] != '.');
  if (!(c->cfg.flags & MB_FLAG_COMPRESS)) return appendHost(c, bases, n);
  //  homopolymer compression (merylInput.C:258-262): a run of one letter keeps its first letter; the run state
  //  carries over the pieces of a record (the reference's `skip` argument) and ends with the record.
  //  (Two corners of the reference are not reproduced: a piece boundary of its 2 MB work buffers inside a run of
  //  MIXED case is compared case-sensitively there, and a piece that consists only of the previous piece's last
  //  letter yields one undefined byte there.)
  c->hpBuf.resize(n);
  size_t m = 0;
  uint8_t last = c->hpLast;
  for (size_t i = 0; i < n; i++) {
    const uint8_t b = (uint8_t)bases[i];
    if (b == '.') { c->hpBuf[m++] = '.'; last = 0; continue; }
    if ((b | 0x20) == last) continue;
    c->hpBuf[m++] = (char)b;
    last = b | 0x20;
  }
  c->hpLast = last;
  return appendHost(c, c->hpBuf.data(), m);
}

//  Large buffers that are already page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory)
//  skip the staging memcpy: the DMA engine reads them in place, 64 MB at a time.
static int appendPinnedDirect(mb_ctx *c, const char *&p, uint64_t &n) {
  if (n < 4 * kStageBytes || c->inRecord) return MB_OK;
  MB_TRY(ensureArena(c));
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return MB_OK; }
  if (at.type != cudaMemoryTypeHost) return MB_OK;
  MB_TRY(flushStage(c));                                 //  stream is at a record boundary: padding is harmless
  MB_CUDA(cudaSetDevice(c->device));
  if (!c->ingestStarted) { c->ingestStarted = true; MB_CUDA(cudaEventRecord(c->evIngest[0], c->stream)); }
  const char *const p0 = p;
  int i = 0;
  while (n >= kStageBytes && !batchDue(c)) {
    const u64 recs = kStageBytes / 32;
    MB_TRY(storeReserve(c, recs));
    MB_CUDA(cudaMemcpyAsync(c->dStage[i], p, kStageBytes, cudaMemcpyHostToDevice, c->stream));
    MB_CUDA(mb_launch_pack(c->dStage[i], kStageBytes, (u32 *)c->arena, c->nRecords, c->stream));
    c->st.nLaunches++;
    c->st.h2dBytes += kStageBytes;
    c->nRecords    += recs;
    p += kStageBytes;  n -= kStageBytes;  i ^= 1;
  }
  MB_CUDA(cudaStreamSynchronize(c->stream));             //  the caller may touch its buffer when we return
  if (p != p0) c->inRecord = (p[-1] != '.');
  return MB_OK;
}

extern "C" int mb_add_bases(mb_ctx *c, const char *bases, uint64_t len, int endOfSequence) {
  mb_clear_error();
  if (!c || (!bases && len))  return mb_fail(MB_EINVAL, "mb_add_bases: NULL argument");
  if (c->counted)             return mb_fail(MB_ESTATE, "mb_add_bases: already counted");
  for (;;) {
    MB_TRY(appendPinnedDirect(c, bases, len));            //  stops early when a batch is due
    if (!len || !batchDue(c)) break;
    //  up to and including the next record break through the staged path: that closes the batch
    const char *dot = (const char *)memchr(bases, '.', len);
    const size_t head = dot ? (size_t)(dot - bases) + 1 : (size_t)len;
    MB_TRY(appendHost(c, bases, head));
    bases += head;  len -= head;
    c->inRecord = (dot == nullptr);
    if (!dot) break;
  }
  MB_TRY(appendHost(c, bases, len));
  if (len) c->inRecord = (bases[len - 1] != '.');
  if (endOfSequence) { const char dot = '.'; MB_TRY(appendHost(c, &dot, 1)); c->inRecord = false; }
  return MB_OK;
}

extern "C" int mb_add_bases_device(mb_ctx *c, const char *dBases, uint64_t len) {
  mb_clear_error();
  if (!c || (!dBases && len))  return mb_fail(MB_EINVAL, "mb_add_bases_device: NULL argument");
  if (c->counted)              return mb_fail(MB_ESTATE, "mb_add_bases_device: already counted");
  if (((uintptr_t)dBases & 15) != 0) return mb_fail(MB_EINVAL, "mb_add_bases_device: pointer must be 16-byte aligned");
  MB_TRY(ensureArena(c));
  //  A device chunk holds whole records: the stream is padded to a record boundary before and after.
  if (c->inRecord) { const char dot = '.'; MB_TRY(appendHost(c, &dot, 1)); c->inRecord = false; }
  MB_TRY(flushStage(c));
  if (batchDue(c)) MB_TRY(closeBatch(c));                 //  (a device chunk holds whole records)
  const u64 recs = (len + 31) / 32;
  MB_TRY(storeReserve(c, recs));
  MB_CUDA(cudaSetDevice(c->device));
  if (!c->ingestStarted) { c->ingestStarted = true; MB_CUDA(cudaEventRecord(c->evIngest[0], c->stream)); }
  MB_CUDA(mb_launch_pack((const uint8_t *)dBases, len, (u32 *)c->arena, c->nRecords, c->stream));
  MB_CUDA(cudaStreamSynchronize(c->stream));            //  the caller may reuse dBases when we return
  c->st.nLaunches++;
  c->nRecords += recs;
  return MB_OK;
}

extern "C" int mb_add_file_bytes(mb_ctx *c, const uint8_t *data, uint64_t len) {
  mb_clear_error();
  if (!c || (!data && len))  return mb_fail(MB_EINVAL, "mb_add_file_bytes: NULL argument");
  if (c->counted)            return mb_fail(MB_ESTATE, "mb_add_file_bytes: already counted");
  if (!SeqParser::sniff(data, len < 65536 ? len : 65536))
    return mb_fail(MB_EFORMAT, "input is not FASTA or FASTQ");
  SeqParser ps(parserSink, c);
  const size_t BLK = (size_t)256 << 20;
  for (uint64_t at = 0; at < len; at += BLK) MB_TRY(mb_parse_parallel(parserSink, c, ps, data + at, (size_t)std::min<uint64_t>(BLK, len - at)));
  return ps.finish();
}

extern "C" int mb_add_file(mb_ctx *c, const char *path) {
  mb_clear_error();
  if (!c || !path)  return mb_fail(MB_EINVAL, "mb_add_file: NULL argument");
  if (c->counted)   return mb_fail(MB_ESTATE, "mb_add_file: already counted");

  //  Compressed inputs go through the same external tools the reference uses
  //  (compressedFileReader, src/utility/src/files/compressed-v1-reading.C:90-116).
  const size_t L = strlen(path);
  const char *tool = nullptr;
  if      (L > 3 && !strcmp(path + L - 3, ".gz"))   tool = "gzip -dc";
  else if (L > 4 && !strcmp(path + L - 4, ".bz2"))  tool = "bzip2 -dc";
  else if (L > 3 && !strcmp(path + L - 3, ".xz"))   tool = "xz -dc";
  else if (L > 4 && !strcmp(path + L - 4, ".zst"))  tool = "zstd -dc";

  //  A plain file is mapped and parsed in place, block by block, on several threads (the page cache is read by the
  //  parsing threads themselves); pipes and stdin go through a read buffer.
  if (!tool && strcmp(path, "-") != 0) {
    const int fd = open(path, O_RDONLY);
    struct stat stt;
    if (fd >= 0 && fstat(fd, &stt) == 0 && S_ISREG(stt.st_mode) && stt.st_size > 0) {
      const size_t sz = (size_t)stt.st_size;
      void *m = mmap(nullptr, sz, PROT_READ, MAP_PRIVATE, fd, 0);
      close(fd);
      if (m != MAP_FAILED) {
        madvise(m, sz, MADV_SEQUENTIAL);
        const uint8_t *d = (const uint8_t *)m;
        int rc = MB_OK;
        if (!SeqParser::sniff(d, sz < 65536 ? sz : 65536)) rc = mb_fail(MB_EFORMAT, "'%s' is not FASTA or FASTQ", path);
        SeqParser ps(parserSink, c);
        const size_t BLK = (size_t)1 << 30;                  //  (16 threads x 64 MB: pieces large enough for the parallel staging copy)
        for (size_t at = 0; rc == MB_OK && at < sz; at += BLK) rc = mb_parse_parallel(parserSink, c, ps, d + at, std::min(BLK, sz - at));
        munmap(m, sz);
        if (rc != MB_OK) return rc;
        return ps.finish();
      }
    } else if (fd >= 0) close(fd);
  }

  FILE *f = nullptr;
  if (tool) {
    if (access(path, R_OK) != 0) return mb_fail(MB_EIO, "cannot read '%s'", path);
    std::string cmd = std::string(tool) + " '";             //  the path inside single quotes, a quote in it as '\''
    for (const char *q = path; *q; q++) { if (*q == '\'') cmd += "'\\''"; else cmd += *q; }
    cmd += "'";
    f = popen(cmd.c_str(), "r");
  } else if (!strcmp(path, "-")) {
    f = stdin;
  } else {
    f = fopen(path, "rb");
  }
  if (!f) return mb_fail(MB_EIO, "cannot open '%s'", path);

  const size_t CH = (size_t)128 << 20;                    //  a block is parsed on several threads (parseParallel)
  std::vector<uint8_t> buf(CH);
  SeqParser ps(parserSink, c);
  bool first = true;
  int  rc = MB_OK;
  for (;;) {
    const size_t n = fread(buf.data(), 1, CH, f);
    if (first) {
      first = false;
      if (!SeqParser::sniff(buf.data(), n < 65536 ? n : 65536)) { rc = mb_fail(MB_EFORMAT, "'%s' is not FASTA or FASTQ", path); break; }
    }
    if (n == 0) break;
    if ((rc = mb_parse_parallel(parserSink, c, ps, buf.data(), n)) != MB_OK) break;
    if (n < CH) break;
  }
  if (tool) pclose(f); else if (f != stdin) fclose(f);
  if (rc != MB_OK) return rc;
  return ps.finish();
}


//  One of nShards pieces of a plain FASTA/FASTQ file (multi-GPU: rank r ingests shard r of every input; the reads are split
//  evenly).  The file is cut at byte i * size / nShards, moved forward to the next line that starts a record
//  of the file's kind (the kind of its first record: '>'; or '@' if the line after next starts with '+', since a quality
//  line may start with '@' too).  Whether that guess was
//  right shows at the END of the previous piece: its parser must stand at a record boundary there.  Shard 0 starts at the
//  start of the file, so if every shard reports success every cut was a true record start (by induction) and the shards
//  together are exactly the file; a shard whose end is not a boundary fails with MB_EFORMAT and the job must not be trusted.
static size_t shardCut(const uint8_t *d, size_t n, uint32_t i, uint32_t nShards) {
  if (i == 0) return 0;
  if (i >= nShards) return n;
  size_t p = (size_t)((unsigned __int128)n * i / nShards);
  while (p < n) {
    const uint8_t *lf = (const uint8_t *)memchr(d + p, '\n', n - p);
    if (!lf || (size_t)(lf - d) + 1 >= n) return n;
    p = (size_t)(lf - d) + 1;
    if (d[p] == '>' && d[0] == '>') {                      //  (in a FASTQ file a line that starts with '>' is a quality line)
      //  ... unless the last non-blank line in front of it is a header too: after the header of a record without bases the
      //  reference's grammar reads the next line as sequence whatever it starts with -- not a record start, try the next one
      size_t q = p;
      while (q > 0 && (d[q - 1] == '\n' || d[q - 1] == '\r')) q--;            //  end of the last non-blank line
      while (q > 0 && d[q - 1] != '\n') q--;                                   //  its start
      if (d[q] != '>' || q == p) return p;
      continue;
    }
    if (d[p] == '@' && d[0] == '@') {
      const uint8_t *l1 = (const uint8_t *)memchr(d + p, '\n', n - p);
      const uint8_t *l2 = l1 ? (const uint8_t *)memchr(l1 + 1, '\n', n - (size_t)(l1 + 1 - d)) : nullptr;
      if (l2 && (size_t)(l2 + 1 - d) < n && l2[1] == '+') return p;
    }
  }
  return n;
}

extern "C" int mb_add_file_shard(mb_ctx *c, const char *path, uint32_t shard, uint32_t nShards) {
  mb_clear_error();
  if (!c || !path || nShards == 0 || shard >= nShards)  return mb_fail(MB_EINVAL, "mb_add_file_shard: bad argument");
  if (c->counted)   return mb_fail(MB_ESTATE, "mb_add_file_shard: already counted");
  if (nShards == 1) return mb_add_file(c, path);
  const size_t L = strlen(path);
  if (!strcmp(path, "-") || (L > 3 && (!strcmp(path + L - 3, ".gz") || !strcmp(path + L - 3, ".xz"))) ||
      (L > 4 && (!strcmp(path + L - 4, ".bz2") || !strcmp(path + L - 4, ".zst"))))
    return mb_fail(MB_EINVAL, "'%s': only plain files can be split into shards (a compressed stream has no record index)", path);
  const int fd = open(path, O_RDONLY);
  struct stat stt;
  if (fd < 0 || fstat(fd, &stt) != 0 || !S_ISREG(stt.st_mode)) { if (fd >= 0) close(fd); return mb_fail(MB_EIO, "cannot open '%s'", path); }
  const size_t sz = (size_t)stt.st_size;
  if (sz == 0) { close(fd); return MB_OK; }
  void *m = mmap(nullptr, sz, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (m == MAP_FAILED) return mb_fail(MB_EIO, "cannot map '%s'", path);
  const uint8_t *d = (const uint8_t *)m;
  int rc = MB_OK;
  if (!SeqParser::sniff(d, sz < 65536 ? sz : 65536)) rc = mb_fail(MB_EFORMAT, "'%s' is not FASTA or FASTQ", path);
  const size_t lo = shardCut(d, sz, shard, nShards), hi = shardCut(d, sz, shard + 1, nShards);
  SeqParser ps(parserSink, c);
  const size_t BLK = (size_t)256 << 20;
  for (size_t at = lo; rc == MB_OK && at < hi; at += BLK) rc = mb_parse_parallel(parserSink, c, ps, d + at, std::min(BLK, hi - at));
  if (rc == MB_OK && hi < sz) {                           //  the next shard starts here: is it a record boundary?
    const SeqParser::State st = ps.state();
    if (!(st == SeqParser::A || st == SeqParser::Q || st == SeqParser::G))
      rc = mb_fail(MB_EFORMAT, "'%s': shard %u of %u does not end at a record boundary (multi-line FASTQ?); count this file on one GPU", path, shard, nShards);
  }
  munmap(m, sz);
  if (rc != MB_OK) return rc;
  return (lo < hi || shard == 0) ? ps.finish() : MB_OK;
}


//  ------------------------------------------------------------------------------------------
//  counting
//  ------------------------------------------------------------------------------------------

namespace {

struct Carver {                                       //  bump allocator over a region of the arena
  uint8_t *base; size_t size, off = 0;
  Carver(uint8_t *b, size_t s) : base(b), size(s) {}
  void *take(size_t bytes) {
    const size_t o = alignUp(off);
    if (o + bytes > size) return nullptr;
    off = o + bytes;
    return base + o;
  }
  size_t left() const { const size_t o = alignUp(off); return o >= size ? 0 : size - o; }
};

struct Control {                                      //  small device-side tables, one block in the arena
  u64 binCount [MB_MAX_BINS];
  u64 binStart [MB_MAX_BINS + 1];
  u64 binCursor[MB_MAX_BINS];
  u64 segStart [MB_MAX_BINS + 1];
  u64 segTile  [MB_MAX_BINS + 1];
  u64 fileBase [64];
  u64 fileCount[72];                                  //  multi-GPU: this rank's k-mers per file (+ the round's flags, see runCount)
  u64 allFileCount[32 * 72];                          //  multi-GPU: every rank's
  u64 numDistinct;
  u64 encTotals[4];
  u64 histTotals[3];
  u64 histBigN;
};

//  One pass = one contiguous range of prefix bins, extracted, (exchanged,) sorted, counted, encoded.
struct Pass {
  u32 binLo = 0, binHi = 0;
  u64 nLocal = 0;                                     //  keys this rank extracts
  u64 nKeys  = 0;                                     //  keys this rank sorts (== nLocal on one GPU)
  std::vector<u64> segCount;                          //  sort segments, in memory order
  std::vector<u32> segBin;                            //  multi-GPU u32 suffix keys: prefix bin of every segment (else binLo + i)
  std::vector<std::pair<u64, u64>> encRanges;         //  (first block prefix, number of blocks) to encode
};

}  //  namespace


//  Several GPUs: rank 0 collects the other ranks' value histograms for the master index (collective: every rank calls it).
//  dScratch: device memory for the exchange.
static int gatherHistograms(mb_ctx *c, uint8_t *dScratch, size_t scratchBytes) {
  const u32 world = mb_comm_world(c->comm), me = mb_comm_rank(c->comm);
  cudaStream_t s = c->stream;
  {
      const u64 myWords = 2 * (u64)c->hist.values.size();
      std::vector<u64> cnt(world, 0);
      Carver hc(dScratch, scratchBytes);
      u64 *dCnt = (u64 *)hc.take(8 * (size_t)(world + 1));
      if (!dCnt) return mb_fail(MB_ENOMEM, "internal: no room for the histogram exchange");
      MB_CUDA(cudaMemcpyAsync(dCnt, &myWords, 8, cudaMemcpyHostToDevice, s));
      MB_TRY(mb_comm_allgather_u64(c->comm, dCnt, dCnt + 1, 1, s));
      MB_CUDA(cudaMemcpyAsync(cnt.data(), dCnt + 1, 8 * (size_t)world, cudaMemcpyDeviceToHost, s));
      MB_CUDA(cudaStreamSynchronize(s));
      u64 others = 0;
      for (u32 r = 1; r < world; r++) others += cnt[r];
      std::vector<u64> mine(myWords);
      for (size_t i = 0; i < c->hist.values.size(); i++) { mine[2 * i] = c->hist.values[i]; mine[2 * i + 1] = c->hist.occ[i]; }
      u64 *dSend = (u64 *)hc.take(8 * (size_t)(myWords + 1));
      u64 *dRecv = (u64 *)hc.take(8 * (size_t)(others + 1));
      if (!dSend || !dRecv) return mb_fail(MB_ENOMEM, "internal: no room for the histogram exchange");
      if (myWords) MB_CUDA(cudaMemcpyAsync(dSend, mine.data(), 8 * (size_t)myWords, cudaMemcpyHostToDevice, s));
      MB_TRY(mb_comm_gather_u64(c->comm, dSend, myWords, dRecv, cnt.data(), s));
      MB_CUDA(cudaStreamSynchronize(s));
      if (me == 0 && others) {
        std::vector<u64> got(others);
        MB_CUDA(cudaMemcpy(got.data(), dRecv, 8 * (size_t)others, cudaMemcpyDeviceToHost));
        std::map<u64, u64> m;
        for (size_t i = 0; i < c->hist.values.size(); i++) m[c->hist.values[i]] += c->hist.occ[i];
        for (u64 i = 0; i < others; i += 2) {
          m[got[i]] += got[i + 1];
          if (got[i] == 1) c->hist.nUnique += got[i + 1];
          c->hist.nDistinct += got[i + 1];
          c->hist.nTotal    += got[i] * got[i + 1];
        }
        c->hist.values.clear();  c->hist.occ.clear();
        for (auto &kv : m) { c->hist.values.push_back(kv.first); c->hist.occ.push_back(kv.second); }
      }
    }
  return MB_OK;
}

static int runCount(mb_ctx *c) {
  const mb_config &cfg = c->cfg;
  const u32 k = cfg.k, wP = cfg.wPrefix, W = 8 * c->keyWords;
  const u32 binBits = c->binBits, nBins = 1u << binBits;
  const u32 blockBits = wP - 6;
  const u64 blocksPerBin  = (u64)1 << (wP - binBits);
  const u64 blocksPerFile = (u64)1 << blockBits;
  const u32 binsPerFile   = 1u << (binBits - 6);
  const u32 world = c->comm ? mb_comm_world(c->comm) : 1;
  const u32 me    = c->comm ? mb_comm_rank(c->comm)  : 0;
  cudaStream_t s = c->stream;

  if (cfg.worldSize > 1 && !c->comm)
    return mb_fail(MB_ESTATE, "mb_count: context has worldSize %u but mb_comm_init was not called", cfg.worldSize);

  MB_CUDA(cudaSetDevice(c->device));
  MB_TRY(ensureArena(c));

  EventTimer tAll, tStage;
  tAll.init(); tStage.init();
  tAll.start(s);

  //  ---- finish the ingest ------------------------------------------------------------------
  if (c->inRecord) { const char dot = '.'; MB_TRY(appendRaw(c, &dot, 1)); c->inRecord = false; }
  MB_TRY(flushStage(c));
  MB_TRY(storeReserve(c, 0));
  if (c->ingestStarted) {                              //  device span of the ingest: first chunk queued -> last chunk packed
    MB_CUDA(cudaEventRecord(c->evIngest[1], s));
    MB_CUDA(cudaEventSynchronize(c->evIngest[1]));
    float m = 0;  cudaEventElapsedTime(&m, c->evIngest[0], c->evIngest[1]);  c->st.msPack = m;
  }
  const size_t storeBytes = (size_t)c->nRecords * MB_STORE_REC_WORDS * 4;
  const size_t padBytes   = (size_t)MB_STORE_PAD_RECORDS * MB_STORE_REC_WORDS * 4;
  MB_CUDA(cudaMemsetAsync(c->arena + storeBytes, 0, padBytes, s));
  c->st.nBases = c->nRecords * 32;

  //  the staging buffers at the top of the arena are free from here on
  Carver top(c->arena + alignUp(storeBytes + padBytes), c->arenaBytes - alignUp(storeBytes + padBytes));

  Control *dCtl = (Control *)top.take(sizeof(Control));
  u64 *dHistSmall = (u64 *)top.take((size_t)MB_HIST_SMALL * sizeof(u64));
  if (!dCtl || !dHistSmall) return mb_fail(MB_ENOMEM, "device arena too small after the read store");
  MB_CUDA(cudaMemsetAsync(dCtl, 0, sizeof(Control), s));
  MB_CUDA(cudaMemsetAsync(dHistSmall, 0, (size_t)MB_HIST_SMALL * sizeof(u64), s));

  //  ---- K2a: k-mers per prefix bin ---------------------------------------------------------
  MbExtractParams xp;
  xp.store = (const u32 *)c->arena;  xp.nPositions = c->nRecords * 32;
  xp.k = k;  xp.mode = cfg.mode;  xp.binBits = binBits;  xp.binLo = 0;  xp.binHi = nBins;  xp.suffix32 = 0;
  xp.noPrefilter = getenv("MB_NO_PREFILTER") ? 1 : 0;
  xp.keptPerMille = 0;

  tStage.start(s);
  MB_CUDA(mb_launch_bin_histogram(xp, dCtl->binCount, s));
  tStage.stop(s);
  c->st.nLaunches++;
  std::vector<u64> binCount(nBins);
  MB_CUDA(cudaMemcpyAsync(binCount.data(), dCtl->binCount, nBins * sizeof(u64), cudaMemcpyDeviceToHost, s));
  MB_CUDA(cudaStreamSynchronize(s));
  c->st.msHistogram += tStage.ms();
  c->st.d2hBytes += nBins * sizeof(u64);

  u64 nLocalKmers = 0;
  for (u32 b = 0; b < nBins; b++) nLocalKmers += binCount[b];

  //  multi-GPU: everybody learns everybody's per-file counts
  std::vector<u64> C((size_t)world * 64, 0);            //  C[r][f]
  {
    for (u32 b = 0; b < nBins; b++) C[(size_t)me * 64 + b / binsPerFile] += binCount[b];
    if (world > 1) {
      //  65 words per rank: the 64 counts and the round's flags (bit 0: this rank counts the round as a batch -- it is closing
      //  one, or has closed one before; bit 1: this rank is in mb_count, no more reads follow)
      std::vector<u64> mine(65), all((size_t)world * 65);
      for (u32 f = 0; f < 64; f++) mine[f] = C[(size_t)me * 64 + f];
      mine[64] = ((c->closing || !c->batches.empty()) ? 1 : 0) | (c->finalRound ? 2 : 0);
      MB_CUDA(cudaMemcpyAsync(dCtl->fileCount, mine.data(), 65 * sizeof(u64), cudaMemcpyHostToDevice, s));
      MB_TRY(mb_comm_allgather_u64(c->comm, dCtl->fileCount, dCtl->allFileCount, 65, s));
      MB_CUDA(cudaMemcpyAsync(all.data(), dCtl->allFileCount, all.size() * sizeof(u64), cudaMemcpyDeviceToHost, s));
      MB_CUDA(cudaStreamSynchronize(s));
      c->roundAnyBatch = false;  c->roundAllFinal = true;
      for (u32 r = 0; r < world; r++) {
        for (u32 f = 0; f < 64; f++) C[(size_t)r * 64 + f] = all[(size_t)r * 65 + f];
        if (all[(size_t)r * 65 + 64] & 1)    c->roundAnyBatch = true;
        if (!(all[(size_t)r * 65 + 64] & 2)) c->roundAllFinal = false;
      }
    } else { c->roundAnyBatch = c->closing;  c->roundAllFinal = true; }
  }
  //  One rank's batch is everybody's batch: the images of this round stay in host memory and are merged at the end.
  const bool asBatch = c->closing || c->roundAnyBatch;
  uint8_t *const outBuf = asBatch ? nullptr : c->outBuf;
  std::vector<u64> fileTotal(64, 0);
  for (u32 r = 0; r < world; r++) for (u32 f = 0; f < 64; f++) fileTotal[f] += C[(size_t)r * 64 + f];

  //  On several GPUs, too, a k-mer travels and is sorted as the 32 bits below its prefix bin when they fit a word
  //  (k <= 21; MB_MULTI_SUFFIX32=0 turns it off: measured at 2 and 8 B200s).  Routing is then by bin,
  //  so everybody also learns everybody's per-bin counts.
  auto envOff = [](const char *name) { const char *e = getenv(name); return e && e[0] == '0'; };
  const bool want32 = world > 1 && !envOff("MB_MULTI_SUFFIX32") && 2 * k - binBits <= 32 && !c->stopAfterSort;

  //  Extraction fused with the exchange (the default since it ran on 2 and 8 B200s; MB_P2P_SCATTER=0: NCCL all-to-all).
  //  Every rank maps every other rank's arena (CUDA IPC) and the extraction kernel stores a kept k-mer straight into the
  //  receive segment of the GPU that owns its file -- the bin's "start" handed to the kernel is the peer address expressed
  //  as an index relative to the local key buffer, so the kernel itself does not change.  The k-mers cross NVLink tile by
  //  tile under the extraction; no send buffer is filled and no all-to-all follows.
  bool p2p = world > 1 && !envOff("MB_P2P_SCATTER");
  struct PeerMap {
    MbComm *comm = nullptr;  std::vector<uint8_t *> base;
    ~PeerMap() { if (comm) mb_comm_close_peer_arenas(comm, base.data()); }
  } peers;
  if (p2p) {
    //  Every rank tries to map every other rank's arena; the ranks then agree: if one of them could not (no peer access
    //  between two of the GPUs, a rank in another container), all of them route through the NCCL all-to-all instead.
    peers.base.assign(world, nullptr);
    const int orc = mb_comm_open_peer_arenas(c->comm, c->arena, dCtl->allFileCount, peers.base.data(), s);
    peers.comm = c->comm;
    const u64 mine = orc == MB_OK ? 1 : 0;
    std::vector<u64> all(world, 0);
    MB_CUDA(cudaMemcpyAsync(dCtl->fileCount, &mine, 8, cudaMemcpyHostToDevice, s));
    MB_TRY(mb_comm_allgather_u64(c->comm, dCtl->fileCount, dCtl->allFileCount, 1, s));
    MB_CUDA(cudaMemcpyAsync(all.data(), dCtl->allFileCount, (size_t)world * 8, cudaMemcpyDeviceToHost, s));
    MB_CUDA(cudaStreamSynchronize(s));
    for (u32 r = 0; r < world; r++) if (!all[r]) p2p = false;
    if (!p2p) {
      mb_comm_close_peer_arenas(c->comm, peers.base.data());
      peers.comm = nullptr;
      mb_clear_error();
      if (getenv("MB_TRACE")) fprintf(stderr, "[meryl-b200] rank %u: no peer mapping on every rank, routing through NCCL\n", me);
    }
  }
  //  With the peer scatter every key form is routed by BIN: an owned bin arrives as one sort segment, so the sort works on the
  //  2k - binBits bits below the bin instead of the 2k - 6 below the file (k = 31: nine digit passes instead of ten).
  const bool byBinFull = p2p;
  std::vector<u64> Cb;                                  //  Cb[r][b]
  std::vector<u64> binTotal;
  if (want32 || byBinFull) {
    u64 *dBinAll = (u64 *)top.take((size_t)(world + 1) * nBins * sizeof(u64));
    if (!dBinAll) return mb_fail(MB_ENOMEM, "device arena too small after the read store");
    Cb.assign((size_t)world * nBins, 0);
    MB_TRY(mb_comm_allgather_u64(c->comm, dCtl->binCount, dBinAll, nBins, s));
    MB_CUDA(cudaMemcpyAsync(Cb.data(), dBinAll, (size_t)world * nBins * sizeof(u64), cudaMemcpyDeviceToHost, s));
    MB_CUDA(cudaStreamSynchronize(s));
    binTotal.assign(nBins, 0);
    for (u32 r = 0; r < world; r++) for (u32 b = 0; b < nBins; b++) binTotal[b] += Cb[(size_t)r * nBins + b];
  }
  u64 nMine = 0;                                        //  k-mers this rank will count
  for (u32 f = 0; f < 64; f++) {
    c->db.owned[f] = (mb_file_owner(f, world) == me);
    if (c->db.owned[f]) nMine += fileTotal[f];
  }
  c->st.nKmers = nMine;

  const u64 histBigCap = nMine / MB_HIST_SMALL + 1024;
  u32 *dHistBig = (u32 *)top.take(histBigCap * sizeof(u32));
  if (!dHistBig) return mb_fail(MB_ENOMEM, "device arena too small after the read store");

  //  (Round 2 measured a second stream for the copy-back of a group's encoded blocks, alternating between two halves of
  //  the encode region: no gain end to end -- 7.2 s against 7.0 s at 100 Gbp -- and it is gone.)
  struct SortEvents { cudaEvent_t afterHist = nullptr; ~SortEvents() { if (afterHist) cudaEventDestroy(afterHist); } } side;
  //  ---- plan the passes --------------------------------------------------------------------
  const size_t R = top.left();
  const u64 nBlocksTotal = (u64)1 << wP;
  size_t encBytes = std::max<size_t>(std::max<size_t>((size_t)512 << 20, R / 6), (size_t)nBlocksTotal * 200 / world);
  if (encBytes > R / 2) encBytes = R / 2;
  u32 tileKeys = mb_sort_tile_keys(W);
  //  per key: two key buffers + tile status; fixed: ragged tiles and tables of every segment
  const u32 radix = mb_sort_radix();
  const double perKey = 2.0 * W + (double)radix * 8 / tileKeys + 16.0 / tileKeys;
  const size_t fixed  = (size_t)nBins * (radix * 8 + 16 * radix * 8) + ((size_t)8 << 20);
  if (R < encBytes + fixed + ((size_t)64 << 20)) return mb_fail(MB_ENOMEM, "device arena too small: %zu bytes left after the read store", R);
  u64 capKeys = (u64)((double)(R - encBytes - fixed) / perKey);
  capKeys = capKeys / 4096 * 4096;
  //  testing knob: cap the keys of a pass so that small inputs exercise the multi-pass machinery
  const u64 capLimit = getenv("MB_MAX_PASS_KEYS") ? strtoull(getenv("MB_MAX_PASS_KEYS"), nullptr, 10) : 0;
  if (capLimit && capKeys > capLimit) capKeys = capLimit;

  std::vector<Pass> passes;
  u32 sortBits;
  bool suffix32 = false;                                //  u32 suffix keys (k <= 21): half the bytes, twice the keys per pass
  bool byBin    = false;                                //  several GPUs: the sort segments are the owned bins (else the owned files)
  bool inPlace  = false;                                //  ... sorted in place group by group (one GPU)
  u64  gcapKeys = 0;                                    //  ... keys per group
  if (world == 1) {
    //  consecutive bins, as many as fit; segments = bins, sorted on the bits below the bin
    sortBits = 2 * k - binBits;
    //  A pass is cut at a data-file boundary (a multiple of binsPerFile bins) whenever every file fits a pass:
    //  the extraction kernel then tests pass membership on the first three bases alone (extract.cu).
    bool fileCuts = true;
    auto planBins = [&](u64 cap, std::vector<Pass> &out) -> int {
      out.clear();
      fileCuts = true;
      for (u32 f = 0; f < 64 && fileCuts; f++) {
        u64 inFile = 0;
        for (u32 b = f * binsPerFile; b < (f + 1) * binsPerFile; b++) inFile += binCount[b];
        if (inFile > cap) fileCuts = false;
      }
      Pass cur;
      for (u32 b = 0; b < nBins; b++) {
        if (binCount[b] > cap) return -1;
        bool cut = cur.nKeys + binCount[b] > cap;
        if (fileCuts && b % binsPerFile == 0) {            //  would the whole file still fit?
          u64 inFile = 0;
          for (u32 bb = b; bb < b + binsPerFile; bb++) inFile += binCount[bb];
          cut = cur.nKeys + inFile > cap;
        } else if (fileCuts) cut = false;
        if (cut && cur.nKeys > 0) {
          cur.binHi = b;  out.push_back(cur);
          cur = Pass();  cur.binLo = b;
        }
        cur.nKeys += binCount[b];
        cur.segCount.push_back(binCount[b]);
      }
      cur.binHi = nBins;
      out.push_back(cur);
      for (Pass &P : out) {
        P.nLocal = P.nKeys;
        P.encRanges.push_back(std::make_pair((u64)P.binLo * blocksPerBin, (u64)(P.binHi - P.binLo) * blocksPerBin));
      }
      return 0;
    };

    if (sortBits <= 32 && !c->stopAfterSort && !getenv("MB_NO_SUFFIX32")) {
      //  u32 keys, sorted IN PLACE group by group: a pass holds its keys once (4 B per key); a group of consecutive bins
      //  (<= gcapKeys keys) is sorted between its place in the pass and a scratch buffer T of 8 B per group key, which then
      //  takes the group's distinct k-mers expanded to u64 for the encoder.  So a pass holds about twice the keys of the
      //  ping-pong layout (100 Gbp / k=21 on one B200: 3 passes over the read store instead of 5).  MB_SORT_PINGPONG=1 keeps
      //  the round-1 layout (two whole-pass buffers, one sort launch per digit and pass) for A/B.
      const bool pingpong = getenv("MB_SORT_PINGPONG") != nullptr;
      const u32 tile32 = mb_sort_tile_keys(4);
      const double perKey32 = pingpong ? 2.0 * 4 + (double)radix * 8 / tile32 + 16.0 / tile32 : 4.0;
      u64 gKeys = std::min<u64>((u64)512 << 20, std::max<u64>(4096, R / 512));
      if (getenv("MB_GROUP_KEYS")) gKeys = std::max<u64>(64, strtoull(getenv("MB_GROUP_KEYS"), nullptr, 10));   //  testing knob: many small groups
      size_t enc32 = std::max<size_t>((size_t)4 << 30, (size_t)nBlocksTotal * 200);
      if (!pingpong) enc32 = std::max<size_t>(enc32, (size_t)24 * gKeys);
      if (enc32 > R / 3) { enc32 = R / 3; if (!pingpong) gKeys = std::min<u64>(gKeys, enc32 / 24); }
      const size_t scratch = pingpong ? 0 : alignUp(8 * (size_t)gKeys + 64) + alignUp(mb_sort_workspace_bytes(gKeys / tile32 + nBins + 1, nBins, 6));
      if (R > enc32 + fixed + scratch + ((size_t)64 << 20)) {
        u64 cap32 = (u64)((double)(R - enc32 - fixed - scratch) / perKey32);
        cap32 = cap32 / 8192 * 8192;
        if (capLimit && cap32 > capLimit) cap32 = capLimit;
        std::vector<Pass> p32;
        if (cap32 > 0 && planBins(cap32, p32) == 0) {
          bool ok = true;
          for (const Pass &P : p32) {
            const u64 gcap = pingpong ? std::min<u64>(P.nKeys / 2 + 8, (u64)(enc32 / 3 / 8)) : gKeys;
            for (u64 cnt : P.segCount) if (cnt > gcap || cnt >= ((u64)1 << 32)) ok = false;
          }
          if (ok) { passes.swap(p32); suffix32 = true; inPlace = !pingpong; gcapKeys = gKeys; }
        }
      }
    }
    if (!suffix32 && planBins(capKeys, passes) != 0) {
      u32 worst = 0;
      for (u32 b = 0; b < nBins; b++) if (binCount[b] > binCount[worst]) worst = b;
      return mb_fail(MB_ENOMEM, "prefix bin %u holds %llu k-mers, more than the %llu that fit the device at once "
                                "(spill-and-merge is not implemented; use more GPUs)", worst, (u64)binCount[worst], (u64)capKeys);
    }
  } else {
    //  slot j of every rank = files [j*world, (j+1)*world); a pass takes q consecutive slots.
    //  Every rank must fit what it sends and what it receives, in every pass: pick the largest q.
    //  The plan must come out the same on every rank, so the capacities are the smallest of all ranks'.
    const u32 tile32 = mb_sort_tile_keys(4);
    const double perKey32 = 2.0 * 4 + (double)radix * 8 / tile32 + 16.0 / tile32;
    size_t enc32 = std::max<size_t>((size_t)4 << 30, (size_t)nBlocksTotal * 200 / world);
    if (enc32 > R / 3) enc32 = R / 3;
    u64 cap32 = 0;
    if (want32 && R > enc32 + fixed + ((size_t)64 << 20)) {
      cap32 = (u64)((double)(R - enc32 - fixed) / perKey32) / 8192 * 8192;
      if (capLimit && cap32 > capLimit) cap32 = capLimit;
    }
    {
      u64 caps[3] = { capKeys, cap32, (u64)enc32 };
      std::vector<u64> all((size_t)world * 3);
      MB_CUDA(cudaMemcpyAsync(dCtl->fileCount, caps, sizeof(caps), cudaMemcpyHostToDevice, s));
      MB_TRY(mb_comm_allgather_u64(c->comm, dCtl->fileCount, dCtl->allFileCount, 3, s));
      MB_CUDA(cudaMemcpyAsync(all.data(), dCtl->allFileCount, all.size() * sizeof(u64), cudaMemcpyDeviceToHost, s));
      MB_CUDA(cudaStreamSynchronize(s));
      for (u32 r = 0; r < world; r++) {
        capKeys = std::min(capKeys, all[(size_t)r * 3]);
        cap32   = std::min(cap32,   all[(size_t)r * 3 + 1]);
        enc32   = (size_t)std::min<u64>(enc32, all[(size_t)r * 3 + 2]);
      }
    }
    const u32 slots = (64 + world - 1) / world;          //  (the last slot is short when the world size does not divide 64)
    //  Greedy: a pass takes consecutive slots while every rank still fits what it sends and what it receives.  (Canonical
    //  k-mers are skewed towards the low files, so the passes get longer towards the end.)  Every rank knows the whole count
    //  matrix and the agreed capacities, so every rank cuts at the same slots.
    typedef std::vector<std::pair<u32, u32>> SlotRanges;
    auto planSlots = [&](u64 cap, bool s32, SlotRanges &out) -> bool {
      out.clear();
      std::vector<u64> snd(world), rcv(world);
      auto add = [&](u32 slot, std::vector<u64> &sn, std::vector<u64> &rc) {
        for (u32 f = slot * world; f < std::min(64u, (slot + 1) * world); f++) {
          for (u32 r = 0; r < world; r++) sn[r] += C[(size_t)r * 64 + f];
          rc[mb_file_owner(f, world)] += fileTotal[f];
        }
      };
      auto fits = [&](const std::vector<u64> &sn, const std::vector<u64> &rc) {
        for (u32 r = 0; r < world; r++) if (sn[r] > cap || rc[r] > cap) return false;
        return true;
      };
      u32 p0 = 0;
      while (p0 < slots) {
        std::fill(snd.begin(), snd.end(), 0);  std::fill(rcv.begin(), rcv.end(), 0);
        add(p0, snd, rcv);
        if (!fits(snd, rcv)) return false;                 //  a single slot is too much
        u32 p1 = p0 + 1;
        while (p1 < slots) {
          std::vector<u64> sn(snd), rc(rcv);
          add(p1, sn, rc);
          if (!fits(sn, rc)) break;
          snd.swap(sn);  rcv.swap(rc);  p1++;
        }
        if (s32)                                            //  every owned bin must fit a count group (see the groups below)
          for (u32 f = p0 * world; f < std::min(64u, p1 * world); f++) {
            const u32 r = mb_file_owner(f, world);
            const u64 gcap = std::min<u64>(std::max(snd[r], rcv[r]) / 2 + 8, (u64)(enc32 / 3 / 8));
            for (u32 b = f * binsPerFile; b < (f + 1) * binsPerFile; b++)
              if (binTotal[b] > gcap || binTotal[b] >= ((u64)1 << 32)) return false;
          }
        out.push_back(std::make_pair(p0, p1));
        p0 = p1;
      }
      return true;
    };
    SlotRanges ranges;
    if (want32 && cap32 > 0 && planSlots(cap32, true, ranges)) suffix32 = true;
    else if (!planSlots(capKeys, false, ranges))
      return mb_fail(MB_ENOMEM, "a single data file's k-mers do not fit one of the %u devices (capacity %llu keys); use more GPUs", world, (u64)capKeys);
    byBin = suffix32 || byBinFull;
    sortBits = byBin ? 2 * k - binBits : 2 * k - 6;
    for (const auto &rg : ranges) {
      Pass P;
      const u32 fLo = rg.first * world, fHi = std::min(64u, rg.second * world);
      P.binLo = fLo * binsPerFile;  P.binHi = fHi * binsPerFile;
      for (u32 f = fLo; f < fHi; f++) {
        P.nLocal += C[(size_t)me * 64 + f];
        if (!c->db.owned[f]) continue;
        P.nKeys += fileTotal[f];
        if (byBin) {
          for (u32 b = f * binsPerFile; b < (f + 1) * binsPerFile; b++) { P.segCount.push_back(binTotal[b]); P.segBin.push_back(b); }
        } else {
          P.segCount.push_back(fileTotal[f]);
        }
        if (!suffix32) P.encRanges.push_back(std::make_pair((u64)f * blocksPerFile, blocksPerFile));
      }
      passes.push_back(P);
    }
  }
  const u32 keyBytes = suffix32 ? 4 : W;
  c->st.keyBytes = keyBytes;
  c->st.nPasses = (u32)passes.size();

  tileKeys = mb_sort_tile_keys(keyBytes);
  MbSortPlan plan;
  mb_sort_plan(sortBits, &plan);
  c->st.sortPasses = plan.nPasses;

  //  host images
  for (u32 f = 0; f < 64; f++) { c->db.dataVec[f].clear(); c->db.dataPtr[f] = nullptr; c->db.dataLen[f] = 0; c->db.index[f].clear(); c->db.streamed[f] = false; }
  c->outUsed = 0;

  //  mb_set_output_dir: the encoded data files do not stay in host memory, they go from the device to the directory
  //  while the count runs, in parallel across files (filesink.h).  Only for a count that is a single batch.
  struct Streamer {
    MbFileSink *sink = nullptr;  int fd[64];  cudaEvent_t ready = nullptr;
    Streamer() { for (int i = 0; i < 64; i++) fd[i] = -1; }
    ~Streamer() {
      if (sink) mb_filesink_drain(sink);
      for (int i = 0; i < 64; i++) if (fd[i] >= 0) close(fd[i]);
      if (ready) cudaEventDestroy(ready);
    }
  } stream;
  const bool streaming = !c->outDir.empty() && !asBatch && c->batches.empty() && !c->stopAfterSort &&
                         !(cfg.flags & (MB_FLAG_NO_IMAGES | MB_FLAG_KEEP_COUNTED));
  if (streaming) {
    if (mkdir(c->outDir.c_str(), 0777) != 0 && errno != EEXIST) return mb_fail(MB_EIO, "cannot create directory '%s': %s", c->outDir.c_str(), strerror(errno));
    char path[4096];
    for (u32 f = 0; f < 64; f++) {
      if (!c->db.owned[f]) continue;
      mb_file_name(path, sizeof(path), c->outDir.c_str(), f, false);
      stream.fd[f] = open(path, O_CREAT | O_TRUNC | O_WRONLY, 0666);
      if (stream.fd[f] < 0) return mb_fail(MB_EIO, "cannot create '%s': %s", path, strerror(errno));
      c->db.streamed[f] = true;
    }
    stream.sink = mb_filesink_get();
    MB_CUDA(cudaEventCreateWithFlags(&stream.ready, cudaEventDisableTiming));
  }
  std::vector<u64> fileBase(64, 0);

  std::vector<u64> segStart, segTile, binStart(MB_MAX_BINS + 1);
  EventTimer tSortHist;  tSortHist.init();
  cudaEvent_t &evAfterHist = side.afterHist;  MB_CUDA(cudaEventCreate(&evAfterHist));

  for (size_t pi = 0; pi < passes.size(); pi++) {
    const Pass &P = passes[pi];
    const u32 nSegs = (u32)P.segCount.size();
    const u64 n = P.nKeys;
    const u64 bufKeys = std::max(P.nLocal, P.nKeys);

    //  ---- carve this pass ------------------------------------------------------------------
    Carver pc((uint8_t *)top.base + alignUp(top.off), top.left());
    void *bufA = pc.take((size_t)keyBytes * bufKeys + 64);
    void *bufB = inPlace ? pc.take(8 * (size_t)gcapKeys + 64)          //  scratch of one group: sort partner, then its distinct k-mers as u64
                         : pc.take((size_t)keyBytes * bufKeys + 64);

    //  extraction layout: every bin of the pass is a contiguous run, bins ascending
    std::fill(binStart.begin(), binStart.end(), 0);
    {
      u64 at = 0;
      for (u32 b = 0; b < nBins; b++) {
        binStart[b] = at;
        if (b >= P.binLo && b < P.binHi) at += binCount[b];
      }
      binStart[nBins] = at;
    }
    //  sort layout
    segStart.assign(nSegs + 1, 0);  segTile.assign(nSegs + 1, 0);
    u64 nTiles = 0;
    {
      u64 at = 0;
      for (u32 i = 0; i < nSegs; i++) {
        segStart[i] = at;  segTile[i] = nTiles;
        at += P.segCount[i];
        nTiles += (P.segCount[i] + tileKeys - 1) / tileKeys;
      }
      segStart[nSegs] = at;  segTile[nSegs] = nTiles;
    }
    //  K5 + K6 work on one group of whole segments at a time.  Full keys (8/16 B): a single group, counted in place.
    //  u32 suffix keys: groups of consecutive bins whose distinct k-mers, expanded to u64, fit the scratch buffer
    //  (in-place form: the group is also the unit of the sort).
    struct Group { u32 segLo, segHi, binLo; u64 keyLo, keyHi, tileLo, tileHi; std::vector<std::pair<u64, u64>> encRanges; };
    std::vector<Group> groups;
    auto buildGroups = [&](u64 gcap) {
      groups.clear();
      if (!suffix32) {
        Group g;  g.segLo = 0;  g.segHi = nSegs;  g.binLo = P.binLo;  g.keyLo = 0;  g.keyHi = n;  g.tileLo = 0;  g.tileHi = nTiles;
        g.encRanges = P.encRanges;
        groups.push_back(g);
        return;
      }
      Group g;  g.segLo = 0;  g.binLo = P.binLo;  g.keyLo = 0;  g.tileLo = 0;
      u64 have = 0;
      auto segBin = [&](u32 i) -> u32 { return P.segBin.empty() ? P.binLo + i : P.segBin[i]; };
      for (u32 i = 0; i <= nSegs; i++) {
        //  a group is a run of consecutive bins (several GPUs: the owned files of a pass are not adjacent)
        if (i == nSegs || (have > 0 && have + P.segCount[i] > gcap) || (i > g.segLo && segBin(i) != segBin(i - 1) + 1)) {
          g.segHi = i;  g.keyHi = segStart[i];  g.tileHi = segTile[i];
          //  the count kernel adds the segment index to binLo: pass-wide indexes (wraps), or -- in place -- the group's own
          g.binLo = inPlace ? segBin(g.segLo) : segBin(g.segLo) - g.segLo;
          g.encRanges.clear();
          g.encRanges.push_back(std::make_pair((u64)segBin(g.segLo) * blocksPerBin, (u64)(g.segHi - g.segLo) * blocksPerBin));
          groups.push_back(g);
          g.segLo = i;  g.keyLo = segStart[i];  g.tileLo = segTile[i];  have = 0;
        }
        if (i < nSegs) have += P.segCount[i];
      }
    };
    size_t sortWs = std::max(mb_sort_workspace_bytes(nTiles, nSegs ? nSegs : 1, plan.nPasses), mb_count_workspace_bytes(n));
    if (inPlace) {                                        //  the sort's tables are per group
      buildGroups(gcapKeys);
      sortWs = 0;
      for (const Group &G : groups)
        sortWs = std::max(sortWs, std::max(mb_sort_workspace_bytes(G.tileHi - G.tileLo, std::max(1u, G.segHi - G.segLo), plan.nPasses), mb_count_workspace_bytes(G.keyHi - G.keyLo)));
    }
    void *ws = pc.take(sortWs);
    if (!bufA || !bufB || !ws) return mb_fail(MB_ENOMEM, "internal: pass %zu does not fit the arena", pi);

    if (p2p) {
      //  Everybody says where this pass's receive buffer lies in its arena.  The all-gather is also the barrier after
      //  which no rank still reads the buffers of the previous pass.
      std::vector<u64> recvOff(world);
      const u64 myOff = (u64)((uint8_t *)bufB - c->arena);
      tStage.start(s);
      MB_CUDA(cudaMemcpyAsync(dCtl->fileCount, &myOff, 8, cudaMemcpyHostToDevice, s));
      MB_TRY(mb_comm_allgather_u64(c->comm, dCtl->fileCount, dCtl->allFileCount, 1, s));
      MB_CUDA(cudaMemcpyAsync(recvOff.data(), dCtl->allFileCount, (size_t)world * 8, cudaMemcpyDeviceToHost, s));
      tStage.stop(s);
      MB_CUDA(cudaStreamSynchronize(s));
      c->st.msExchange += tStage.ms();
      //  bin b of file f goes to owner(f): into the segment of its unit (file; bin for suffix keys), behind the chunks of
      //  the ranks before me (and, inside my chunk of a file, behind my lower bins)
      std::vector<u64> at(world, 0);                      //  keys laid out so far in every destination's buffer
      bool aligned = true;
      auto target = [&](u32 d, u64 elem) -> u64 {
        const uint8_t *t = peers.base[d] + recvOff[d] + elem * keyBytes;
        const long long diff = (long long)((uintptr_t)t - (uintptr_t)bufA);
        if (diff % (long long)keyBytes) aligned = false;
        return (u64)(diff / (long long)keyBytes);          //  two's complement: bufA + index * keyBytes == t (mod 2^64)
      };
      for (u32 f = P.binLo / binsPerFile; f < P.binHi / binsPerFile; f++) {
        const u32 d = mb_file_owner(f, world);
        if (byBin) {
          for (u32 b = f * binsPerFile; b < (f + 1) * binsPerFile; b++) {
            u64 before = 0;
            for (u32 r = 0; r < me; r++) before += Cb[(size_t)r * nBins + b];
            binStart[b] = target(d, at[d] + before);
            at[d] += binTotal[b];
          }
        } else {
          u64 base = at[d];
          for (u32 r = 0; r < me; r++) base += C[(size_t)r * 64 + f];
          for (u32 b = f * binsPerFile; b < (f + 1) * binsPerFile; b++) { binStart[b] = target(d, base); base += binCount[b]; }
          at[d] += fileTotal[f];
        }
      }
      if (getenv("MB_TRACE")) fprintf(stderr, "[meryl-b200] rank %u pass %zu: peer scatter, %llu keys out, %llu keys in, %u-byte keys\n", me, pi, (u64)P.nLocal, (u64)n, keyBytes);
      if (!aligned) return mb_fail(MB_ESTATE, "internal: a peer arena is not aligned to the key size");
      if (at[me] != n) return mb_fail(MB_ESTATE, "internal: receive layout of pass %zu holds %llu keys, expected %llu", pi, (u64)at[me], (u64)n);
    }

    Carver ec((uint8_t *)pc.base + alignUp(pc.off), pc.left());       //  encode region: everything that is left

    MB_CUDA(cudaMemcpyAsync(dCtl->binStart, binStart.data(), (nBins + 1) * sizeof(u64), cudaMemcpyHostToDevice, s));
    MB_CUDA(cudaMemcpyAsync(dCtl->segStart, segStart.data(), (nSegs + 1) * sizeof(u64), cudaMemcpyHostToDevice, s));
    MB_CUDA(cudaMemcpyAsync(dCtl->segTile,  segTile.data(),  (nSegs + 1) * sizeof(u64), cudaMemcpyHostToDevice, s));
    MB_CUDA(cudaMemcpyAsync(dCtl->fileBase, fileBase.data(), 64 * sizeof(u64), cudaMemcpyHostToDevice, s));
    MB_CUDA(cudaMemsetAsync(dCtl->binCursor, 0, sizeof(dCtl->binCursor), s));
    c->st.h2dBytes += (nBins + 1 + 2 * (nSegs + 1) + 64) * sizeof(u64);

    //  ---- K2/K3 ------------------------------------------------------------------------------
    xp.binLo = P.binLo;  xp.binHi = P.binHi;  xp.suffix32 = suffix32 ? 1 : 0;
    xp.keptPerMille = nLocalKmers ? (u32)std::max<u64>(1, P.nLocal * 1000 / nLocalKmers) : 0;
    tStage.start(s);
    if (P.nLocal) { MB_CUDA(mb_launch_extract(xp, bufA, dCtl->binStart, dCtl->binCursor, s)); c->st.nLaunches++; }
    tStage.stop(s);
    MB_CUDA(cudaStreamSynchronize(s));
    c->st.msExtract += tStage.ms();

    //  ---- multi-GPU: route k-mers to the owners of their files ----------------------------------
    void *unsorted = bufA, *spare = bufB;
    if (p2p) {
      //  the k-mers are where they belong already; wait until every rank's extraction has finished
      tStage.start(s);
      MB_TRY(mb_comm_allgather_u64(c->comm, dCtl->fileCount, dCtl->allFileCount, 1, s));
      tStage.stop(s);
      MB_CUDA(cudaStreamSynchronize(s));
      c->st.msExchange += tStage.ms();
      unsorted = bufB;  spare = bufA;
    } else if (world > 1) {
      const u32 fLo = P.binLo / binsPerFile, fHi = P.binHi / binsPerFile;
      std::vector<u64> sendStart(fHi - fLo);
      for (u32 f = fLo; f < fHi; f++) sendStart[f - fLo] = binStart[f * binsPerFile];
      tStage.start(s);
      if (suffix32) MB_TRY(mb_comm_exchange_suffixes(c->comm, bufA, bufB, P.binLo, P.binHi, binsPerFile, &binStart[P.binLo], Cb.data(), nBins, s));
      else          MB_TRY(mb_comm_exchange_keys(c->comm, c->keyWords, bufA, bufB, fLo, fHi, sendStart.data(), C.data(), s));
      tStage.stop(s);
      MB_CUDA(cudaStreamSynchronize(s));
      c->st.msExchange += tStage.ms();
      unsorted = bufB;  spare = bufA;
    }

    //  ---- K4 ---------------------------------------------------------------------------------
    plan.maxSegKeys = 1;
    for (u64 cnt : P.segCount) if (cnt > plan.maxSegKeys) plan.maxSegKeys = cnt;
    //  one sort of [lo, hi) segments between `from` and `partner` (both addressed with the pass's key indexes); times and counts it
    auto sortSegments = [&](void *from, void *partner, u64 nk, u32 sLo, u32 sHi, u64 nt, int *inPartner) -> int {
      u32 nl = 0;
      tSortHist.start(s);
      MB_CUDA(mb_launch_sort(keyBytes, from, partner, nk, dCtl->segStart + sLo, dCtl->segTile + sLo, sHi - sLo, nt, plan, ws, sortWs, inPartner, &nl, s, evAfterHist));
      tStage.stop(s);                                 //  [evAfterHist, tStage.b] = the sweep passes
      MB_CUDA(cudaStreamSynchronize(s));
      { float m = 0;  cudaEventElapsedTime(&m, tSortHist.a, evAfterHist);  c->st.msSortHist += m;
        m = 0;        cudaEventElapsedTime(&m, evAfterHist, tStage.b);     c->st.msSort     += m; }
      c->st.nLaunches += nl;
      if (nk) c->st.sortLaunchKeys += nk * plan.nPasses;
      return MB_OK;
    };
    void *sorted = unsorted, *other = spare;
    if (!inPlace) {
      int inSpare = 0;
      MB_TRY(sortSegments(unsorted, spare, n, 0, nSegs, nTiles, &inSpare));
      sorted = inSpare ? spare : unsorted;
      other  = inSpare ? unsorted : spare;
    }

    if (c->stopAfterSort) {                           //  debug: hand back the sorted instances
      const size_t at = c->sortedLo.size();
      c->sortedLo.resize(at + n);
      if (c->keyWords == 2) c->sortedHi.resize(at + n);
      if (n) {
        if (c->keyWords == 1) {
          MB_CUDA(cudaMemcpy(c->sortedLo.data() + at, sorted, n * 8, cudaMemcpyDeviceToHost));
        } else {
          std::vector<u64> tmp(2 * n);
          MB_CUDA(cudaMemcpy(tmp.data(), sorted, n * 16, cudaMemcpyDeviceToHost));
          for (u64 i = 0; i < n; i++) { c->sortedLo[at + i] = tmp[2 * i]; c->sortedHi[at + i] = tmp[2 * i + 1]; }
        }
      }
      continue;
    }

    //  ---- K5 + K6, one group of whole segments at a time ----------------------------------------------
    //  Full keys (8/16 B): a single group, counted in place.  u32 suffix keys: the pass is cut into
    //  groups of consecutive bins whose distinct k-mers, expanded to u64, fit the spare key buffer.
    if (!inPlace) buildGroups(std::min<u64>(bufKeys / 2 + 8, (u64)(ec.left() / 3 / 8)));

    for (size_t gi = 0; gi < groups.size(); gi++) {
      const Group &G = groups[gi];
      const u64 gKeys = G.keyHi - G.keyLo;
      Carver gc((uint8_t *)ec.base + alignUp(ec.off), ec.left());                              //  per-group slice of the encode region

      if (inPlace && gKeys) {
        //  the group's segments between their place in the pass and the scratch buffer (addressed with the pass's key indexes:
        //  the scratch's virtual base lies gKeyLo keys below it); an odd number of digit passes leaves them in the scratch
        int inScratch = 0;
        uint8_t *virt = (uint8_t *)bufB - (size_t)G.keyLo * 4;
        MB_TRY(sortSegments(bufA, virt, gKeys, G.segLo, G.segHi, G.tileHi - G.tileLo, &inScratch));
        if (inScratch) MB_CUDA(cudaMemcpyAsync((uint8_t *)bufA + (size_t)G.keyLo * 4, bufB, (size_t)gKeys * 4, cudaMemcpyDeviceToDevice, s));
        sorted = bufA;  other = bufB;
      }

      void *distinctKeys;  u64 *runStart;  u32 encKeyWords;
      tStage.start(s);
      if (!suffix32) {
        distinctKeys = sorted;  runStart = (u64 *)other;  encKeyWords = c->keyWords;
        MB_CUDA(mb_launch_count(c->keyWords, sorted, n, runStart, &dCtl->numDistinct, ws, s));
      } else {
        distinctKeys = other;  encKeyWords = 1;
        if (gKeys > (inPlace ? gcapKeys : bufKeys / 2 + 8)) return mb_fail(MB_ENOMEM, "internal: group %zu of pass %zu (%llu k-mers) does not fit the spare key buffer", gi, pi, (u64)gKeys);
        runStart = (u64 *)gc.take(8 * (size_t)(gKeys + 1));
        void *cws = gc.take(8 * (size_t)(2 * (G.tileHi - G.tileLo) + 2));   //  (the sort workspace still holds the tile table)
        if (!runStart || !cws) return mb_fail(MB_ENOMEM, "internal: no room for the run table of pass %zu group %zu", pi, gi);
        if (inPlace)
          MB_CUDA(mb_launch_count_suffix32((const u32 *)sorted, mb_sort_tile_info(ws, G.tileHi - G.tileLo, std::max(1u, G.segHi - G.segLo), plan.nPasses),
                                           0, G.tileHi - G.tileLo, G.keyHi, G.binLo, 2 * k - binBits, (u64 *)other, runStart,
                                           &dCtl->numDistinct, cws, s));
        else
          MB_CUDA(mb_launch_count_suffix32((const u32 *)sorted, mb_sort_tile_info(ws, nTiles, nSegs ? nSegs : 1, plan.nPasses),
                                           G.tileLo, G.tileHi, G.keyHi, G.binLo, 2 * k - binBits, (u64 *)other, runStart,
                                           &dCtl->numDistinct, cws, s));
      }
      u64 D = 0;
      MB_CUDA(cudaMemcpyAsync(&D, &dCtl->numDistinct, sizeof(u64), cudaMemcpyDeviceToHost, s));
      MB_CUDA(mb_launch_value_histogram(runStart, &dCtl->numDistinct, gKeys, dHistSmall, dHistBig, &dCtl->histBigN, histBigCap, dCtl->histTotals, s));
      tStage.stop(s);
      MB_CUDA(cudaStreamSynchronize(s));
      c->st.msCount += tStage.ms();
      c->st.nLaunches += gKeys ? 2 : 0;
      c->st.d2hBytes += 8;

      if (cfg.flags & MB_FLAG_KEEP_COUNTED) {
        const size_t at = c->cLo.size();
        c->cLo.resize(at + D);  c->cCnt.resize(at + D);
        if (encKeyWords == 2) c->cHi.resize(at + D);
        std::vector<u64> rs(D + 1);
        if (D) {
          MB_CUDA(cudaMemcpy(rs.data(), runStart, (D + 1) * 8, cudaMemcpyDeviceToHost));
          if (encKeyWords == 1) {
            MB_CUDA(cudaMemcpy(c->cLo.data() + at, distinctKeys, D * 8, cudaMemcpyDeviceToHost));
          } else {
            std::vector<u64> tmp(2 * D);
            MB_CUDA(cudaMemcpy(tmp.data(), distinctKeys, D * 16, cudaMemcpyDeviceToHost));
            for (u64 i = 0; i < D; i++) { c->cLo[at + i] = tmp[2 * i]; c->cHi[at + i] = tmp[2 * i + 1]; }
          }
          for (u64 i = 0; i < D; i++) c->cCnt[at + i] = (u32)(rs[i + 1] - rs[i]);
        }
      }

      //  ---- K6: every block range of the group ---------------------------------------------------------
      tStage.start(s);
      MB_CUDA(cudaMemcpyAsync(dCtl->fileBase, fileBase.data(), 64 * sizeof(u64), cudaMemcpyHostToDevice, s));   //  bytes already emitted per file
      for (size_t ri = 0; ri < G.encRanges.size(); ri++) {
        MbEncodeParams ep;
        ep.keys = distinctKeys;  ep.runStart = runStart;  ep.nDistinct = D;  ep.keyWords = encKeyWords;
        ep.k = k;  ep.wPrefix = wP;  ep.counts = nullptr;  ep.streamMode = 0;
        ep.blockLo = G.encRanges[ri].first;
        ep.nBlocks = G.encRanges[ri].second;

        Carver rc((uint8_t *)gc.base + alignUp(gc.off), gc.left());
        MbBlockMeta *dMeta = (MbBlockMeta *)rc.take(ep.nBlocks * sizeof(MbBlockMeta));
        void *dScanTmp     = rc.take(mb_encode_scan_tmp_bytes(ep.nBlocks));
        if (!dMeta || !dScanTmp) return mb_fail(MB_ENOMEM, "device arena too small for the block table of pass %zu (%llu blocks)", pi, (u64)ep.nBlocks);

        MB_CUDA(mb_launch_encode_layout(ep, dMeta, dCtl->encTotals, dScanTmp, s));
        c->st.nLaunches += 4;

        //  chunk boundaries: first block of every file touched by this range
        const u32 fileLo = (u32)(ep.blockLo >> blockBits), fileHi = (u32)((ep.blockLo + ep.nBlocks - 1) >> blockBits);
        const u32 nf = fileHi - fileLo + 1;
        std::vector<MbBlockMeta> fm(nf + 1);
        std::vector<u64> fblk(nf + 1);
        for (u32 i = 0; i < nf; i++) {
          const u64 fb = (u64)(fileLo + i) << blockBits;
          fblk[i] = fb > ep.blockLo ? fb - ep.blockLo : 0;
        }
        fblk[nf] = ep.nBlocks;
        for (u32 i = 0; i < nf; i++)
          MB_CUDA(cudaMemcpyAsync(&fm[i], dMeta + fblk[i], sizeof(MbBlockMeta), cudaMemcpyDeviceToHost, s));
        u64 encTotals[4];
        MB_CUDA(cudaMemcpyAsync(encTotals, dCtl->encTotals, sizeof(encTotals), cudaMemcpyDeviceToHost, s));
        MB_CUDA(cudaStreamSynchronize(s));
        fm[nf].outByte = encTotals[0];  fm[nf].subFirst = (u32)encTotals[1];
        c->st.d2hBytes += nf * sizeof(MbBlockMeta) + sizeof(encTotals);

        //  greedy groups of whole files that fit what is left of the arena
        const size_t encLeft = rc.left();
        u32 fi = 0;
        while (fi < nf) {
          u32 fj = fi;
          while (fj < nf) {
            const size_t bytes = fm[fj + 1].outByte - fm[fi].outByte;
            const size_t subs  = (size_t)(fm[fj + 1].subFirst - fm[fi].subFirst);
            const size_t idx   = (size_t)(fblk[fj + 1] - fblk[fi]) * 24;
            const size_t tot   = alignUp(bytes) + alignUp(subs * 4) + alignUp(idx) + 3 * kAlign;
            if (tot > encLeft) break;
            fj++;
          }
          if (fj == fi)
            return mb_fail(MB_ENOMEM, "encoded data file %u of pass %zu (%llu bytes) does not fit the %zu bytes left on the device",
                           fileLo + fi, pi, (u64)(fm[fi + 1].outByte - fm[fi].outByte), encLeft);

          MbEncodeChunk ch;
          ch.chunkFirst = fblk[fi];  ch.chunkCount = fblk[fj] - fblk[fi];
          ch.chunkByte0 = fm[fi].outByte;  ch.chunkSub0 = fm[fi].subFirst;
          const size_t bytes = fm[fj].outByte - fm[fi].outByte;
          const size_t subs  = (size_t)(fm[fj].subFirst - fm[fi].subFirst);

          Carver cc((uint8_t *)rc.base + alignUp(rc.off), rc.left());
          uint8_t *dOut  = (uint8_t *)cc.take(bytes);
          u32     *dSub  = (u32 *)cc.take(subs * 4);
          u64     *dIdx  = (u64 *)cc.take(ch.chunkCount * 24);
          if (!dOut || !dSub || !dIdx) return mb_fail(MB_ENOMEM, "internal: encode chunk does not fit");

          MB_CUDA(cudaMemsetAsync(dOut, 0, bytes, s));
          MB_CUDA(mb_launch_encode_write(ep, ch, dMeta, dCtl->encTotals, dSub, dOut, dIdx, dCtl->fileBase, s));
          c->st.nLaunches += D ? 2 : 1;

          if (streaming) {
            MB_CUDA(cudaEventRecord(stream.ready, s));
            for (u32 f = fi; f < fj; f++) {
              const u32 file = fileLo + f;
              const size_t fb = fm[f + 1].outByte - fm[f].outByte;
              const size_t ib = (size_t)(fblk[f + 1] - fblk[f]) * 24;
              std::vector<uint8_t> &iv = c->db.index[file];
              mb_filesink_submit(stream.sink, c->device, stream.fd[file], c->db.dataLen[file],
                                 dOut + (fm[f].outByte - ch.chunkByte0), fb, stream.ready);
              c->db.dataLen[file] += fb;
              const size_t i0 = iv.size();
              iv.resize(i0 + ib);
              MB_CUDA(cudaMemcpyAsync(iv.data() + i0, (uint8_t *)dIdx + (fblk[f] - ch.chunkFirst) * 24, ib, cudaMemcpyDeviceToHost, s));
              c->st.d2hBytes += fb + ib;
            }
            MB_CUDA(cudaStreamSynchronize(s));
            MB_TRY(mb_filesink_drain(stream.sink));     //  (the encode region is reused by the next chunk)
          } else if (!(cfg.flags & MB_FLAG_NO_IMAGES)) {
            for (u32 f = fi; f < fj; f++) {
              const u32 file = fileLo + f;
              const size_t fb = fm[f + 1].outByte - fm[f].outByte;
              const size_t ib = (size_t)(fblk[f + 1] - fblk[f]) * 24;
              std::vector<uint8_t> &iv = c->db.index[file];
              uint8_t *dst;
              if (outBuf) {                             //  append to the caller's buffer: a file's pieces arrive back to back
                if (c->outUsed + fb > c->outCap)
                  return mb_fail(MB_ENOMEM, "output buffer too small: data file %u needs %zu more bytes, %zu of %zu used", file, fb, c->outUsed, c->outCap);
                if (c->db.dataLen[file] == 0) c->db.dataPtr[file] = outBuf + c->outUsed;
                else if (c->db.dataPtr[file] + c->db.dataLen[file] != outBuf + c->outUsed)
                  return mb_fail(MB_ESTATE, "internal: data file %u is not contiguous in the output buffer", file);
                dst = outBuf + c->outUsed;
                c->outUsed += fb;
              } else {
                std::vector<uint8_t> &dv = c->db.dataVec[file];
                const size_t d0 = dv.size();
                dv.resize(d0 + fb);
                dst = dv.data() + d0;
              }
              c->db.dataLen[file] += fb;
              const size_t i0 = iv.size();
              iv.resize(i0 + ib);
              MB_CUDA(cudaMemcpyAsync(dst, dOut + (fm[f].outByte - ch.chunkByte0), fb, cudaMemcpyDeviceToHost, s));
              MB_CUDA(cudaMemcpyAsync(iv.data() + i0, (uint8_t *)dIdx + (fblk[f] - ch.chunkFirst) * 24, ib, cudaMemcpyDeviceToHost, s));
              c->st.d2hBytes += fb + ib;
            }
          }
          MB_CUDA(cudaStreamSynchronize(s));
          for (u32 f = fi; f < fj; f++) fileBase[fileLo + f] += fm[f + 1].outByte - fm[f].outByte;
          fi = fj;
        }
      }
      tStage.stop(s);
      MB_CUDA(cudaStreamSynchronize(s));
      c->st.msEncode += tStage.ms();
    }
  }

  if (p2p) {
    //  Unmap the peers' arenas, then a barrier: nobody returns (and later frees its arena) while somebody still maps it.
    mb_comm_close_peer_arenas(c->comm, peers.base.data());
    peers.comm = nullptr;
    MB_TRY(mb_comm_allgather_u64(c->comm, dCtl->fileCount, dCtl->allFileCount, 1, s));
    MB_CUDA(cudaStreamSynchronize(s));
  }

  if (!c->stopAfterSort) {
    //  ---- value histogram -> host ------------------------------------------------------------
    std::vector<u64> hs(MB_HIST_SMALL);
    u64 totals[3], bigN = 0;
    MB_CUDA(cudaMemcpyAsync(hs.data(), dHistSmall, (size_t)MB_HIST_SMALL * 8, cudaMemcpyDeviceToHost, s));
    MB_CUDA(cudaMemcpyAsync(totals, dCtl->histTotals, sizeof(totals), cudaMemcpyDeviceToHost, s));
    MB_CUDA(cudaMemcpyAsync(&bigN, &dCtl->histBigN, 8, cudaMemcpyDeviceToHost, s));
    MB_CUDA(cudaStreamSynchronize(s));
    if (bigN > histBigCap) return mb_fail(MB_ENOMEM, "internal: value histogram overflow list too small");
    std::vector<u32> big(bigN);
    if (bigN) MB_CUDA(cudaMemcpy(big.data(), dHistBig, bigN * 4, cudaMemcpyDeviceToHost));
    c->st.d2hBytes += (size_t)MB_HIST_SMALL * 8 + bigN * 4;

    std::map<u64, u64> bigMap;
    for (u64 i = 0; i < bigN; i++) bigMap[big[i]]++;
    c->hist.values.clear();  c->hist.occ.clear();
    for (u64 v = 1; v < MB_HIST_SMALL; v++) if (hs[v]) { c->hist.values.push_back(v); c->hist.occ.push_back(hs[v]); }
    for (auto &kv : bigMap) { c->hist.values.push_back(kv.first); c->hist.occ.push_back(kv.second); }
    c->hist.nUnique = totals[0];  c->hist.nDistinct = totals[1];  c->hist.nTotal = totals[2];
    c->st.nUnique = totals[0];  c->st.nDistinct = totals[1];  c->st.nTotal = totals[2];

    //  multi-GPU: rank 0 collects the other ranks' histograms for the master index (after the merge when this round is a batch)
    if (world > 1 && !asBatch) MB_TRY(gatherHistograms(c, (uint8_t *)top.base + alignUp(top.off), top.left()));

    mb_build_master_index(k, wP, c->hist, c->db.master);
    c->st.dbBytes = (me == 0) ? c->db.master.size() : 0;
    for (u32 f = 0; f < 64; f++) {
      if (!outBuf) c->db.dataPtr[f] = c->db.dataVec[f].data();          //  vectors may have moved while growing
      c->st.dbBytes += c->db.dataLen[f] + c->db.index[f].size();
    }
  }

  tAll.stop(s);
  c->st.msTotal = tAll.ms();
  c->counted = true;
  return MB_OK;
}


extern "C" int mb_set_input_compression(mb_ctx *c, int on) {
  mb_clear_error();
  if (!c)          return mb_fail(MB_EINVAL, "mb_set_input_compression: NULL context");
  if (c->counted)  return mb_fail(MB_ESTATE, "mb_set_input_compression: already counted");
  if (c->inRecord) return mb_fail(MB_ESTATE, "mb_set_input_compression: inside a record (end it with endOfSequence first)");
  if (on) c->cfg.flags |= MB_FLAG_COMPRESS; else c->cfg.flags &= ~(uint32_t)MB_FLAG_COMPRESS;
  c->hpLast = 0;
  return MB_OK;
}

extern "C" int mb_set_output_dir(mb_ctx *c, const char *outDir) {
  mb_clear_error();
  if (!c || !outDir || !outDir[0])  return mb_fail(MB_EINVAL, "mb_set_output_dir: NULL argument");
  if (c->counted)                   return mb_fail(MB_ESTATE, "mb_set_output_dir: already counted");
  c->outDir = outDir;
  return MB_OK;
}

extern "C" int mb_set_output_buffer(mb_ctx *c, void *hostBuf, uint64_t bytes) {
  mb_clear_error();
  if (!c || (!hostBuf && bytes))  return mb_fail(MB_EINVAL, "mb_set_output_buffer: NULL argument");
  if (c->counted)                 return mb_fail(MB_ESTATE, "mb_set_output_buffer: already counted");
  c->outBuf = (uint8_t *)hostBuf;  c->outCap = bytes;  c->outUsed = 0;
  return MB_OK;
}

//  The images of the round that has just been counted become a batch; the store starts again.
static void stashBatch(mb_ctx *c) {
  std::unique_ptr<MbDbImages> b(new MbDbImages());
  for (u32 f = 0; f < 64; f++) {
    b->dataVec[f] = std::move(c->db.dataVec[f]);  b->dataLen[f] = c->db.dataLen[f];  b->dataPtr[f] = b->dataVec[f].data();
    b->index[f]   = std::move(c->db.index[f]);    b->owned[f] = c->db.owned[f];
    c->db.dataVec[f].clear();  c->db.dataLen[f] = 0;  c->db.dataPtr[f] = nullptr;  c->db.index[f].clear();
  }
  c->batches.push_back(std::move(b));
  c->nRecords = 0;  c->ingestStarted = false;  c->counted = false;
}

//  Count what the store holds as one batch, set its images aside, start the store again (finishBatch).
//  Several GPUs: the call is collective -- a rank whose store fills during the ingest meets the others here, each of them
//  in its own closeBatch or, when its reads are all in, in mb_count's rounds.
static int closeBatch(mb_ctx *c) {
  const mb_stats prev = c->st;
  const bool inRecord = c->inRecord;
  c->inRecord = false;  c->closing = true;
  c->st.nBases = 0;  c->st.nKmers = 0;  c->st.nPasses = 0;  c->st.msPack = 0;  c->st.msTotal = 0;
  const int rc = runCount(c);
  c->inRecord = inRecord;  c->closing = false;
  if (rc != MB_OK) return rc;
  //  (the other statistics are sums already)
  c->st.nBases += prev.nBases;  c->st.nKmers += prev.nKmers;  c->st.nPasses += prev.nPasses;
  c->st.msPack += prev.msPack;  c->st.msTotal += prev.msTotal;
  c->st.nBatches = prev.nBatches + 1;
  stashBatch(c);
  if (getenv("MB_TRACE")) fprintf(stderr, "[meryl-b200] rank %u: batch %u closed (%llu k-mers counted by this rank so far)\n", c->cfg.rank, c->st.nBatches, (u64)c->st.nKmers);
  return MB_OK;
}

//  mergeBatches: the batches' blocks are decoded, merged with saturating sums and encoded again with the block writer's
//  rules, file group by file group, inside this context's arena (the read store is not needed any more).  Several GPUs:
//  every rank merges the files it owns, then rank 0 collects the histograms.
static int mergeBatches(mb_ctx *c) {
  std::vector<const MbDbImages *> in;
  for (auto &b : c->batches) in.push_back(b.get());
  const bool multi = c->comm && mb_comm_world(c->comm) > 1;
  MbMergeEndpoints E;
  E.memInputs = in.data();  E.nMem = (u32)in.size();  E.k = c->cfg.k;  E.wPrefix = c->cfg.wPrefix;
  E.outImages = &c->db;  E.outBuf = c->outBuf;  E.outCap = c->outCap;  E.outUsed = &c->outUsed;  E.outHist = &c->hist;
  E.arena = c->arena;  E.arenaBytes = c->arenaBytes;
  E.outPrefix = c->cfg.wPrefix;  E.streamMode = 0;
  if (multi) E.onlyFiles = c->db.owned;
  c->outUsed = 0;
  mb_union_stats us;
  MB_TRY(mb_merge_pipeline(MB_OP_MERGE_BATCHES, 0, nullptr, 0, nullptr, c->device, 0, &us, &E));
  c->batches.clear();
  c->st.nUnique = c->hist.nUnique;  c->st.nDistinct = c->hist.nDistinct;  c->st.nTotal = c->hist.nTotal;     //  this rank's files
  if (multi) MB_TRY(gatherHistograms(c, c->arena, c->arenaBytes));
  mb_build_master_index(c->cfg.k, c->cfg.wPrefix, c->hist, c->db.master);
  c->st.dbBytes = (!multi || mb_comm_rank(c->comm) == 0) ? c->db.master.size() : 0;
  for (u32 f = 0; f < 64; f++) c->st.dbBytes += c->db.dataLen[f] + c->db.index[f].size();
  c->st.nLaunches += us.nLaunches;  c->st.h2dBytes += us.h2dBytes;  c->st.d2hBytes += us.d2hBytes;
  c->st.msMerge = us.msTotal;  c->st.msTotal += us.msTotal;
  c->counted = true;
  return MB_OK;
}

extern "C" int mb_count(mb_ctx *c) {
  mb_clear_error();
  if (!c)          return mb_fail(MB_EINVAL, "mb_count: NULL context");
  if (c->counted)  return mb_fail(MB_ESTATE, "mb_count: already counted");
  if (c->cfg.worldSize <= 1) {
    if (c->batches.empty()) { MB_TRY(runCount(c)); c->st.nBatches = 1; return MB_OK; }
    if (c->nRecords || c->fill || c->inRecord) MB_TRY(closeBatch(c));        //  the last batch
    return mergeBatches(c);
  }
  //  Several GPUs: rounds of the collective count until every rank has said that nothing follows on its side.  A rank that has
  //  closed batches (or meets one that has) counts the round as a batch; a rank whose reads are all in keeps taking part with
  //  an empty store while another rank still closes batches (its files receive that rank's k-mers).
  c->finalRound = true;
  for (;;) {
    if (!c->batches.empty()) MB_TRY(closeBatch(c));
    else {
      const mb_stats prev = c->st;
      MB_TRY(runCount(c));
      if (c->roundAnyBatch) { c->st.nBatches = prev.nBatches + 1; stashBatch(c); }
    }
    if (c->roundAllFinal) break;
  }
  c->finalRound = false;
  if (c->batches.empty()) { c->st.nBatches = 1; return MB_OK; }
  return mergeBatches(c);
}


//  ------------------------------------------------------------------------------------------
//  results
//  ------------------------------------------------------------------------------------------

extern "C" int mb_get_image(mb_ctx *c, uint32_t which, const uint8_t **ptr, uint64_t *len) {
  mb_clear_error();
  if (!c || !ptr || !len || which > 128)  return mb_fail(MB_EINVAL, "mb_get_image: bad argument");
  if (!c->counted)                        return mb_fail(MB_ESTATE, "mb_get_image: call mb_count first");
  if (which < 64 && c->db.streamed[which]) return mb_fail(MB_ESTATE, "mb_get_image: data file %u was streamed to '%s' (mb_set_output_dir)", which, c->outDir.c_str());
  if (which < 64) { *ptr = c->db.dataPtr[which]; *len = c->db.dataLen[which]; return MB_OK; }
  const std::vector<uint8_t> &v = which < 128 ? c->db.index[which - 64] : c->db.master;
  *ptr = v.data();
  *len = v.size();
  return MB_OK;
}

extern "C" int mb_write_db(mb_ctx *c, const char *outDir) {
  mb_clear_error();
  if (!c || !outDir)  return mb_fail(MB_EINVAL, "mb_write_db: NULL argument");
  if (!c->counted)    return mb_fail(MB_ESTATE, "mb_write_db: call mb_count first");
  if (c->cfg.flags & MB_FLAG_NO_IMAGES) return mb_fail(MB_ESTATE, "mb_write_db: context was created with MB_FLAG_NO_IMAGES");
  for (u32 f = 0; f < 64; f++)
    if (c->db.streamed[f] && c->outDir != outDir)
      return mb_fail(MB_ESTATE, "mb_write_db: the data files were streamed to '%s' while counting (mb_set_output_dir); write the rest there", c->outDir.c_str());
  return mb_write_db_dir(c->db, outDir, c->cfg.rank == 0);
}

extern "C" int mb_get_stats(mb_ctx *c, mb_stats *st) {
  mb_clear_error();
  if (!c || !st) return mb_fail(MB_EINVAL, "mb_get_stats: NULL argument");
  *st = c->st;
  return MB_OK;
}

extern "C" int mb_get_counted(mb_ctx *c, const uint64_t **lo, const uint64_t **hi, const uint32_t **cnt, uint64_t *n) {
  mb_clear_error();
  if (!c || !lo || !cnt || !n)  return mb_fail(MB_EINVAL, "mb_get_counted: NULL argument");
  if (!c->counted || !(c->cfg.flags & MB_FLAG_KEEP_COUNTED)) return mb_fail(MB_ESTATE, "mb_get_counted: count with MB_FLAG_KEEP_COUNTED first");
  *lo = (const uint64_t *)c->cLo.data();
  if (hi) *hi = c->keyWords == 2 ? (const uint64_t *)c->cHi.data() : nullptr;
  *cnt = c->cCnt.data();
  *n = c->cLo.size();
  return MB_OK;
}

extern "C" int mb_get_histogram(mb_ctx *c, const uint64_t **values, const uint64_t **occ, uint64_t *n) {
  mb_clear_error();
  if (!c || !values || !occ || !n)  return mb_fail(MB_EINVAL, "mb_get_histogram: NULL argument");
  if (!c->counted)                  return mb_fail(MB_ESTATE, "mb_get_histogram: call mb_count first");
  *values = (const uint64_t *)c->hist.values.data();
  *occ    = (const uint64_t *)c->hist.occ.data();
  *n      = c->hist.values.size();
  return MB_OK;
}

extern "C" int mb_merge_histogram(mb_ctx *c, const uint64_t *values, const uint64_t *occ, uint64_t n) {
  mb_clear_error();
  if (!c || (n && (!values || !occ)))  return mb_fail(MB_EINVAL, "mb_merge_histogram: NULL argument");
  if (!c->counted)                     return mb_fail(MB_ESTATE, "mb_merge_histogram: call mb_count first");
  std::map<u64, u64> m;
  for (size_t i = 0; i < c->hist.values.size(); i++) m[c->hist.values[i]] += c->hist.occ[i];
  for (u64 i = 0; i < n; i++) {
    m[values[i]] += occ[i];
    if (values[i] == 1) c->hist.nUnique += occ[i];
    c->hist.nDistinct += occ[i];
    c->hist.nTotal    += values[i] * occ[i];
  }
  c->hist.values.clear();  c->hist.occ.clear();
  for (auto &kv : m) { c->hist.values.push_back(kv.first); c->hist.occ.push_back(kv.second); }
  mb_build_master_index(c->cfg.k, c->cfg.wPrefix, c->hist, c->db.master);
  return MB_OK;
}


//  engine-internal accessors for the debug entry points and the exchange layer
void mb_ctx_set_stop_after_sort(mb_ctx *c)                               { c->stopAfterSort = true; }
const std::vector<u64> &mb_ctx_sorted_lo(mb_ctx *c)                      { return c->sortedLo; }
const std::vector<u64> &mb_ctx_sorted_hi(mb_ctx *c)                      { return c->sortedHi; }
MbDbImages &mb_ctx_images(mb_ctx *c)                                     { return c->db; }
void mb_ctx_set_comm(mb_ctx *c, MbComm *m)                               { c->comm = m; }
MbComm *mb_ctx_comm(mb_ctx *c)                                           { return c->comm; }
cudaStream_t mb_ctx_stream(mb_ctx *c)                                    { return c->stream; }
int mb_ctx_device(mb_ctx *c)                                             { return c->device; }
const mb_config &mb_ctx_config(mb_ctx *c)                                { return c->cfg; }
