//  Shared host-side declarations for the B200 `meryl count` engine.
#pragma once

#include <stdint.h>
#include <stddef.h>
#include <stdarg.h>
#include <string>
#include <vector>

#include "../../include/meryl_b200.h"

typedef unsigned long long u64;
typedef unsigned int       u32;

//  Thread-local error string behind mb_last_error().
int  mb_fail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
void mb_clear_error(void);

#define MB_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess)                                                                   \
      return mb_fail(MB_ENODEVICE, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

#define MB_TRY(call)  do { int r_ = (call); if (r_ != MB_OK) return r_; } while (0)

//  Bin (pass-planning / partition) granularity: the top binBits bits of the k-mer.
//  Always block aligned because binBits <= wPrefix.
static inline u32 mb_bin_bits(u32 wPrefix) { return wPrefix < 10 ? wPrefix : 10; }

//  ---- planner (planner.cpp) ---------------------------------------------------------------
struct MbPlan {
  u32  wPrefix;
  int  simple;
  u64  memUsed;
  u32  nBatches;
};
int mb_plan_impl(u32 k, u64 nEstimate, u64 memBytes, MbPlan *out);

//  ---- host FASTA/FASTQ parser (parser.cpp) --------------------------------------------------
//  Byte-level state machine equivalent to bufSeqFile::open + loadBases + the '.' record breaks of
//  the reference's work-buffer builder.  Emits the base stream through a sink callback.
class SeqParser {
public:
  typedef int (*Sink)(void *arg, const char *bases, size_t n);   //  appends to the logical base stream
  SeqParser(Sink sink, void *arg) : _sink(sink), _arg(arg) {}

  //  Feed consecutive pieces of ONE file; call finish() at its end.
  int  feed(const uint8_t *d, size_t n);
  int  finish(void);

  //  Format check on the first bytes of a file (needs >= 1001 bytes or the whole file).
  static bool sniff(const uint8_t *d, size_t n);

  //  The machine's state, for parsing a file in pieces on several threads (engine.cu: parseParallel).
  enum State { A, P1, P1CR, P2, P2CR, Q, H, HCR, L, W, G };
  State state(void) const   { return _st; }
  void  setState(State s)   { _st = s; }
  int   breakRecord(void)   { return brk(); }
  std::vector<std::string> &pool(void) { return _pool; }     //  output buffers of the pieces, reused from block to block

private:
  int  emit(const uint8_t *p, size_t n) { return n ? _sink(_arg, (const char *)p, n) : 0; }
  int  brk(void)                        { static const char dot = '.'; return _sink(_arg, &dot, 1); }

  State  _st = A;
  std::vector<std::string> _pool;
  Sink   _sink;
  void  *_arg;
};

//  One block of a file parsed on several host threads, exactly as one sequential pass would (engine.cu).
int mb_parse_parallel(SeqParser::Sink out, void *outArg, SeqParser &ps, const uint8_t *d, size_t n);

//  ---- db assembly on the host (dbwriter.cpp) ------------------------------------------------
struct MbHistogram {                      //  value -> occurrences, plus the master-index totals
  std::vector<u64> values, occ;           //  ascending by value
  u64 nUnique = 0, nDistinct = 0, nTotal = 0;
};

struct MbDbImages {
  //  A data file image lives either in its own vector or, when the caller supplied an output
  //  buffer (mb_set_output_buffer), as a contiguous slice of that buffer.
  std::vector<uint8_t> dataVec[64];
  const uint8_t       *dataPtr[64] = { nullptr };
  size_t               dataLen[64] = { 0 };
  std::vector<uint8_t> index[64];
  std::vector<uint8_t> master;
  bool owned[64];                         //  files this rank owns
  bool streamed[64] = { false };          //  the data file is in the output directory already (mb_set_output_dir): only its length is known here
};

void mb_build_master_index(u32 k, u32 wPrefix, const MbHistogram &h, std::vector<uint8_t> &out);
int  mb_write_db_dir(const MbDbImages &db, const char *dir, bool writeMaster);
void mb_file_name(char *out, size_t cap, const char *dir, u32 file, bool isIndex);

//  ---- decode / merge / encode pipeline (union.cu) with explicit endpoints ----------------------
//  The set operations read directories and write a directory (mb_set_operation); the merge of count batches
//  (merylBlockWriter::mergeBatches, src/utility/src/kmers-v1/kmers-writer-block.C:221-386) reads the batches' images from
//  host memory, writes images, keeps the batches' prefix width and the block writer's rules (a block for every prefix),
//  sums with saturation, and works inside the counting context's device arena.
#define MB_OP_MERGE_BATCHES 100u          //  internal operator: saturating uint32 sum (kmers-writer-block.C:305-309)
#define MB_OP_KEEP_SUFFIX   101u          //  internal one-input operator: keep the k-mers whose low bits equal a constant (count-suffix=)
struct MbMergeEndpoints {
  const MbDbImages *const *memInputs = nullptr;      //  inputs in host memory instead of directories ...
  u32          nMem = 0, k = 0, wPrefix = 0;         //  ... all of this k and prefix width
  MbDbImages  *outImages = nullptr;                  //  output images instead of a directory (data, index; master by the caller)
  uint8_t     *outBuf = nullptr;                     //  optional caller-owned memory for the data images, appended in file order
  size_t       outCap = 0, *outUsed = nullptr;
  MbHistogram *outHist = nullptr;
  uint8_t     *arena = nullptr;                      //  device arena to work in (NULL: allocate one)
  size_t       arenaBytes = 0;
  u64          param2 = 0;                           //  MB_OP_KEEP_SUFFIX: the mask (param is the value)
  u32          outPrefix = 12;                       //  kmers-writer.C:52-58: every non-count operation writes prefix 12
  u32          streamMode = 1;                       //  1: stream writer (blocks of present prefixes), 0: block writer (all prefixes)
  const bool  *onlyFiles = nullptr;                  //  64 flags: the data files to produce (several GPUs: the files this rank owns); NULL = all
};
int mb_merge_pipeline(uint32_t op, uint64_t param, const char *const *inDirs, uint32_t nInputs, const char *outDir,
                      int32_t device, uint64_t workspaceBytes, mb_union_stats *st, const MbMergeEndpoints *ends);

//  engine.cu: free the device arena a finished count context left behind (device < 0: on every device)
void mb_arena_cache_release(int device);
