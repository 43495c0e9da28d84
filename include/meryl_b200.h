/*
 *  meryl_b200.h -- C ABI of the B200-native `meryl count` path.
 *
 *  The reference (marbl/meryl) has no plugin/FFI interface; the count path sits
 *  behind two seams and this header mirrors both:
 *
 *   process seam     argv word grammar + db directory
 *                      src/meryl/merylCommandBuilder.C:188-326 (options), :333-435 (operations),
 *                      :441-563 (output / inputs)          ->  mb_count_files(), and the
 *                      `meryl-b200` CLI built on it (csrc/cli.cpp)
 *
 *   in-process seam  merylOperation::configureCounting()       src/meryl/merylOp-count.C:300-403
 *                                                          ->  mb_plan()
 *                    merylOperation::countThreads(wPrefix, nPrefix, wData, wDataMask)
 *                                                              src/meryl/merylOp.H:210-213,
 *                                                              src/meryl/merylOp-countThreads.C:385-474
 *                                                          ->  mb_create() / mb_count() / mb_destroy()
 *                    merylInput::loadBases(char*, max, &len, &endOfSequence)
 *                                                              src/meryl/merylInput.H:67-70, merylInput.C:246-271
 *                                                          ->  mb_add_bases()  (push form of the same pull)
 *                    bufSeqFile::open / loadBases               src/utility/src/sequence/bufSeqFile-v1.C:54-104, :554-595
 *                                                          ->  mb_add_file(), mb_add_file_bytes()
 *                    merylBlockWriter::addBlock / finish, ~merylFileWriter
 *                                                              src/utility/src/kmers-v1/kmers-writer-block.C:101-218,
 *                                                              src/utility/src/kmers-v1/kmers-writer.C:118-291
 *                                                          ->  mb_write_db(), mb_get_image()
 *
 *  Conventions: plain pointers and sizes only; every function returns 0 on
 *  success or a negative MB_E* code and never calls exit(); mb_last_error()
 *  gives the thread-local message.  A context is bound to one CUDA device and
 *  must be driven from one host thread at a time.  There is NO CPU fallback:
 *  without a usable CUDA device mb_create() fails with MB_ENODEVICE.
 */
#ifndef MERYL_B200_H
#define MERYL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_OK          0
#define MB_EINVAL     -1     /*  bad argument (k out of range, NULL, ...)                                  */
#define MB_ENODEVICE  -2     /*  no usable CUDA device / CUDA runtime error                                */
#define MB_ENOMEM     -3     /*  input does not fit the device workspace                                   */
#define MB_EIO        -4     /*  cannot read input / write output                                          */
#define MB_EFORMAT    -5     /*  input is not FASTA/FASTQ (bufSeqFile::open would reject it)               */
#define MB_ESTATE     -6     /*  call out of order (e.g. mb_write_db before mb_count)                      */
#define MB_ECOMM      -7     /*  NCCL error                                                                */

/*  merylOp: opCount / opCountForward / opCountReverse  (src/meryl/merylOp.H)  */
#define MB_MODE_CANONICAL 0
#define MB_MODE_FORWARD   1
#define MB_MODE_REVERSE   2

typedef struct mb_ctx mb_ctx;

typedef struct mb_config {
  uint32_t k;               /*  mer size, 3..64                     (k=)                                    */
  uint32_t mode;            /*  MB_MODE_*                           (count / count-forward / count-reverse) */
  uint32_t wPrefix;         /*  prefix bits of the db, from mb_plan (wPrefix of countThreads)               */
  int32_t  device;          /*  CUDA device ordinal; -1 = current device                                    */
  uint64_t workspaceBytes;  /*  device arena size; 0 = all free memory minus a margin                       */
  uint32_t rank;            /*  multi-GPU: this process' rank   (0 when worldSize <= 1)                     */
  uint32_t worldSize;       /*  multi-GPU: number of ranks (one per GPU); 0 or 1 = single GPU               */
  uint32_t flags;           /*  MB_FLAG_*                                                                   */
  uint32_t reserved;
} mb_config;

#define MB_FLAG_KEEP_COUNTED   1u   /*  keep (distinct k-mer, count) arrays on the host for mb_get_counted (tests)  */
#define MB_FLAG_NO_IMAGES      2u   /*  do not copy the encoded db to the host (kernel-only benchmarking)           */
#define MB_FLAG_COMPRESS       4u   /*  `compress`: homopolymer-compress the bases of file inputs (mb_add_file / mb_add_file_bytes) before
                                     *  counting -- every run of one letter (case-insensitive, any letter) becomes its first letter, as
                                     *  merylInput::loadBases does with homopolyCompress (src/meryl/merylInput.C:258-262,
                                     *  src/utility/src/sequence/sequence-v1.C:204-260).  Bases pushed with mb_add_bases are taken as they
                                     *  are: at that seam the reference has already compressed them.                                     */

typedef struct mb_stats {
  uint64_t nBases;          /*  positions in the packed store (bases + record breaks)     */
  uint64_t nKmers;          /*  k-mer instances counted (this rank's share when sharded)  */
  uint64_t nDistinct;
  uint64_t nUnique;
  uint64_t nTotal;          /*  sum of stored counts (== nKmers unless a count wrapped)   */
  uint64_t dbBytes;         /*  bytes of the 129 images owned by this rank                */
  uint32_t nPasses;         /*  prefix-range passes used                                  */
  uint32_t nLaunches;       /*  CUDA kernels launched by mb_count (+ ingest)              */
  uint64_t h2dBytes;        /*  host->device bytes moved                                  */
  uint64_t d2hBytes;        /*  device->host bytes moved                                  */
  /*  device time (CUDA events on the launching stream), milliseconds, summed over passes  */
  float    msPack;          /*  ingest span on the device: first chunk queued -> last chunk packed (H2D + K1)  */
  float    msHistogram;     /*  K2a prefix-bin histogram pre-pass                          */
  float    msExtract;       /*  K2/K3 canonical extraction + prefix scatter                */
  float    msSortHist;      /*  K4a digit histograms                                       */
  float    msSort;          /*  K4  onesweep scatter passes                                */
  float    msCount;         /*  K5  run-length count + value histogram                     */
  float    msEncode;        /*  K6  block layout + bit scatter                             */
  float    msExchange;      /*  multi-GPU all-to-all                                       */
  float    msTotal;         /*  whole mb_count()                                           */
  uint32_t sortPasses;      /*  onesweep digit passes per key                              */
  uint32_t keyBytes;        /*  8 (k<=32) or 16                                            */
  uint64_t sortLaunchKeys;  /*  sum over onesweep launches of keys moved                   */
  uint32_t nBatches;        /*  count batches (1 unless the packed reads outgrew the device: merylBlockWriter::finishBatch /
                             *  mergeBatches, src/utility/src/kmers-v1/kmers-writer-block.C:172-386)  */
  float    msMerge;         /*  decode + merge + encode of the batches (0 for one batch)   */
} mb_stats;

/*  Thread-local message of the last failing call on this thread.  */
const char *mb_last_error(void);

/*  == configureCounting: simple-vs-complex decision and findBestPrefixSize
 *  (src/meryl/merylOp-count.C:119-227, :300-403; simple-mode prefix: merylOp-countSimple.C:196-199).
 *  nEstimate is what the reference calls _expNumKmers (n= or the sum of input file sizes);
 *  memBytes is memory= in bytes.  Outputs may be NULL.  */
int mb_plan(uint32_t k, uint64_t nEstimate, uint64_t memBytes,
            uint32_t *wPrefix, int *simpleMode, uint64_t *memUsed, uint32_t *nBatches);

/*  _expNumKmers guess for one input file name: src/meryl/merylOp-count.C:410-433.  */
uint64_t mb_estimate_kmers_in_file(const char *path);

int  mb_create(const mb_config *cfg, mb_ctx **ctx);
void mb_destroy(mb_ctx *ctx);

/*  A process keeps one device arena per GPU between contexts (a service counting job after job allocates its memory
 *  once; with several GPUs the peers' mappings of an arena stay valid from job to job).  The reference frees its
 *  buckets when countThreads returns (src/meryl/merylOp-countThreads.C:466-473); this is the explicit form of that for
 *  a caller that wants the device memory back before the process ends.  MB_NO_ARENA_CACHE=1 disables the cache.  */
void mb_release_device_memory(void);

/*  Input.  All forms append to the same logical base stream ("record . record . ...",
 *  src/meryl/merylOp-countThreads.C:139-231); they may be mixed and called repeatedly.  */
int mb_add_bases(mb_ctx *ctx, const char *bases, uint64_t len, int endOfSequence);  /*  host memory   */
int mb_add_bases_device(mb_ctx *ctx, const char *dBases, uint64_t len);             /*  device memory, work-buffer format ('.' = break)  */
int mb_add_file(mb_ctx *ctx, const char *path);                                     /*  FASTA/FASTQ, optionally .gz/.bz2/.xz via popen like the reference  */
int mb_add_file_bytes(mb_ctx *ctx, const uint8_t *data, uint64_t len);              /*  a FASTA/FASTQ file image in host memory  */
/*  Shard `shard` of `nShards` of a plain FASTA/FASTQ file, cut at record starts (multi-GPU: rank r adds shard r of every
 *  input).  All shards together are exactly the file if every call succeeds; a shard that does not end at a record boundary
 *  (the cut was guessed wrong) fails with MB_EFORMAT.  Compressed files and stdin: MB_EINVAL.  */
int mb_add_file_shard(mb_ctx *ctx, const char *path, uint32_t shard, uint32_t nShards);

/*  Multi-GPU (one process per GPU).  Rank 0 calls mb_comm_unique_id, the caller broadcasts the
 *  128 bytes (torch.distributed / MPI / files), every rank calls mb_comm_init before mb_count.  */
int mb_comm_unique_id(uint8_t id[128]);
int mb_comm_init(mb_ctx *ctx, const uint8_t id[128]);
/*  Which rank owns data file `file` (0..63): min(file, 63-file) mod worldSize -- pairs the heavy low
 *  files with the light high ones (canonical k-mers are skewed), SURVEY.md section 8e.  */
uint32_t mb_file_owner_of(uint32_t file, uint32_t worldSize);

/*  Optional: host memory (ideally page-locked) that receives the 64 data-file images, appended in file
 *  order; mb_get_image(0..63) then points into it.  Without it the library allocates the images itself.
 *  MB_ENOMEM from mb_count if it is too small (the db of N k-mers of which D are distinct needs about
 *  D * (2k - wPrefix + 36) / 8 + 2^wPrefix * 104 bytes).  */
int mb_set_output_buffer(mb_ctx *ctx, void *hostBuf, uint64_t bytes);

/*  Homopolymer compression for the inputs added from now on (the reference's `compress` word applies to the inputs that
 *  follow it: merylCommandBuilder.C:237-240, :546; merylInput.C:258-262).  MB_FLAG_COMPRESS in mb_config = on from the start.
 *  Only between records.  */
int mb_set_input_compression(mb_ctx *ctx, int on);

/*  Optional, before mb_count: the encoded data files go from the device straight into outDir while the count runs, one
 *  host thread per file (the reference's writer has one OpenMP thread per file, merylOp-countThreads.C:452-459), instead
 *  of staying in host memory; mb_write_db(ctx, outDir) then adds the 64 block indexes and merylIndex.  mb_get_image
 *  refuses the streamed data files.  (A count that closed batches keeps its images in memory as before.)  */
int mb_set_output_dir(mb_ctx *ctx, const char *outDir);

/*  Count everything added so far.  After this the encoded db (this rank's files) is in host memory.  */
int mb_count(mb_ctx *ctx);

/*  Images: which = 0..63 -> 0xFFFFFF.merylData, 64..127 -> 0xFFFFFF.merylIndex, 128 -> merylIndex.
 *  Pointers stay valid until mb_destroy.  Files not owned by this rank have len 0.  */
int mb_get_image(mb_ctx *ctx, uint32_t which, const uint8_t **ptr, uint64_t *len);

/*  Writes the files owned by this rank into outDir (created if needed); rank 0 also writes merylIndex.  */
int mb_write_db(mb_ctx *ctx, const char *outDir);

int mb_get_stats(mb_ctx *ctx, mb_stats *st);

/*  With MB_FLAG_KEEP_COUNTED: the sorted distinct k-mers (hi word only for k > 32) and their counts.  */
int mb_get_counted(mb_ctx *ctx, const uint64_t **kmerLo, const uint64_t **kmerHi, const uint32_t **count, uint64_t *n);

/*  Value histogram of the master index: ascending values and their occurrences (this rank's share).  */
int mb_get_histogram(mb_ctx *ctx, const uint64_t **values, const uint64_t **occurrences, uint64_t *n);
/*  Multi-GPU: rank 0 receives the other ranks' histograms/statistics (gathered by the caller) before writing merylIndex.  */
int mb_merge_histogram(mb_ctx *ctx, const uint64_t *values, const uint64_t *occurrences, uint64_t n);

/*  One-shot: `meryl k=K [n=N] [memory=GB] count* <paths> output <outDir>` on one GPU.  */
int mb_count_files(const mb_config *cfg, const char **paths, uint32_t nPaths, const char *outDir, mb_stats *st);

/* ---- `meryl union-sum A B ... output C` (csrc/union.cu) ------------------------
 *
 *  Replaces, for this operator, merylOperation::nextMer + findSumCount (src/meryl/merylOp-nextMer.C:419-683, :44-48)
 *  between merylFileReader (src/utility/src/kmers-v1/kmers-reader.C:456-531, kmers-files.C:25-128) and
 *  merylStreamWriter (src/utility/src/kmers-v1/kmers-writer-stream.C:89-157): every k-mer of any input is written
 *  once with the sum of its values (kmvalu = uint32, wraps; a zero sum is dropped).  The output is what the
 *  reference writes: prefix width 12, no blocks for absent prefixes, 64 data + 64 index files + merylIndex.
 *  Inputs must be meryl v.03 databases of the same k (>= 7), not multisets.  Decode, merge and encode run on the
 *  device, one group of consecutive data files at a time.                                                      */
typedef struct mb_union_stats {
  uint64_t nInputKmers;     /*  (k-mer, value) pairs read from all inputs                 */
  uint64_t nOutputKmers;    /*  distinct k-mers written                                   */
  uint64_t nUnique, nDistinct, nTotal;   /*  master-index statistics of the output       */
  uint64_t inBytes;         /*  .merylData bytes read                                     */
  uint64_t outBytes;        /*  bytes of the 129 files written                            */
  uint64_t h2dBytes, d2hBytes;
  uint32_t nInputs;
  uint32_t nGroups;         /*  groups of data files processed one after the other        */
  uint32_t nLaunches;       /*  CUDA kernels launched                                     */
  uint32_t k;
  float    msDecode;        /*  device time incl. the host reads feeding it (CUDA events) */
  float    msMerge;
  float    msEncode;        /*  value histogram + encode + copy back                      */
  float    msTotal;
} mb_union_stats;

/*  device: CUDA ordinal or -1 (current); workspaceBytes: device arena, 0 = all free memory minus a margin.  */
int mb_union_sum(const char *const *inDirs, uint32_t nInputs, const char *outDir, int32_t device,
                 uint64_t workspaceBytes, mb_union_stats *st);

/*  The whole merge family of merylOperation::nextMer (src/meryl/merylOp-nextMer.C:559-612; operation words of
 *  merylCommandBuilder.C:364-377) through the same decode / merge / encode pipeline; mb_union_sum == MB_OP_UNION_SUM.
 *  "Active" = the inputs that hold a k-mer:
 *    union: value = number of active inputs;  union-min/-max/-sum: min / max / uint32 sum of their values;
 *    intersect*: only k-mers of ALL inputs, value = the first input's / min / max / sum;
 *    subtract: k-mers of the first input, its value minus the others' in turn, dropped as soon as it would not stay positive;
 *    difference: k-mers of the first input that no other input holds;
 *    symmetric-difference: k-mers of exactly one input, with its value (more than two inputs: composed of three folds). */
#define MB_OP_UNION                 0
#define MB_OP_UNION_MIN             1
#define MB_OP_UNION_MAX             2
#define MB_OP_UNION_SUM             3
#define MB_OP_INTERSECT             4
#define MB_OP_INTERSECT_MIN         5
#define MB_OP_INTERSECT_MAX         6
#define MB_OP_INTERSECT_SUM         7
#define MB_OP_SUBTRACT              8
#define MB_OP_DIFFERENCE            9
#define MB_OP_SYMMETRIC_DIFFERENCE 10
int mb_set_operation(uint32_t op, const char *const *inDirs, uint32_t nInputs, const char *outDir, int32_t device,
                     uint64_t workspaceBytes, mb_union_stats *st);

/*  The one-input operators with a number (merylOp-nextMer.C:497-557; words of merylCommandBuilder.C:350-362):
 *  filters keep a k-mer iff its value compares so with `param`; arithmetic changes the value (computed in 64 bits,
 *  stored as uint32; increase / multiply become 2^32-1 only on 64-bit overflow; decrease below the constant, divide or
 *  modulo by 0 give 0, i.e. the k-mer is not written).  Same pipeline: decode, one pass of the merge kernel, encode.  */
#define MB_OP_LESS_THAN            11
#define MB_OP_GREATER_THAN         12
#define MB_OP_AT_LEAST             13
#define MB_OP_AT_MOST              14
#define MB_OP_EQUAL_TO             15
#define MB_OP_NOT_EQUAL_TO         16
#define MB_OP_INCREASE             17
#define MB_OP_DECREASE             18
#define MB_OP_MULTIPLY             19
#define MB_OP_DIVIDE               20
#define MB_OP_DIVIDE_ROUND         21
#define MB_OP_MODULO               22
int mb_value_operation(uint32_t op, uint64_t param, const char *inDir, const char *outDir, int32_t device,
                       uint64_t workspaceBytes, mb_union_stats *st);
/*  What `meryl histogram` / `meryl statistics` print (merylOp-histogram.C:20-100): k, the totals and the value histogram of
 *  a database, read from its master index (host only).  values / occurrences may be NULL to ask for the length (*n) only.  */
int mb_db_histogram(const char *dbDir, uint32_t *k, uint64_t *nUnique, uint64_t *nDistinct, uint64_t *nTotal,
                    uint64_t *values, uint64_t *occurrences, uint64_t cap, uint64_t *n);
/*  `count-suffix=<bases>` (merylCommandBuilder.C:271-274, merylOp-countSimple.C:50-93, :172-199): keeps the k-mers of a count
 *  database that end with the given bases and writes them with the geometry of the reference's simple counter (prefix width
 *  6 + max(0, 2k - 2s - 26), a block for every prefix).  `meryl-b200 count count-suffix=ACG ...` = count, then this.  */
int mb_keep_suffix(const char *inDir, const char *suffix, const char *outDir, int32_t device, uint64_t workspaceBytes,
                   mb_union_stats *st);
/*  The filters' threshold given as a share of the input (merylOperation::initializeThreshold, merylOp-nextMer.C:65-125;
 *  words distinct= / d=, word-frequency= / f= of merylCommandBuilder.C:277-290): fracDistinct -> the smallest value v such
 *  that at least fracDistinct * numDistinct k-mers have a value <= v; wordFrequency -> wordFrequency * numTotal (wins if both
 *  are given); negative = not given.  Host only: reads the histogram in dbDir/merylIndex.  */
int mb_db_threshold(const char *dbDir, double fracDistinct, double wordFrequency, uint64_t *threshold);

/*  merylExactLookup::value for many k-mers at once (src/utility/src/kmers-v1/kmers-lookup.H:139-146, kmers-exact.C;
 *  SURVEY.md 8(f).4): values[i] = the value of k-mer i in the database, 0 if it is absent or its value lies outside
 *  [minValue, maxValue] (the lookup table's load-time filter).  A k-mer is 2 bits per base, first base in the top bits,
 *  A=0 C=1 T=2 G=3, in kmerLo[i] (and kmerHi[i] for k > 32; NULL otherwise); it is looked up as given -- pass the canonical
 *  form for databases made by `count`, as meryl-lookup does.  The blocks are decoded on the device and every query is a
 *  binary search in the list of its data file.  */
int mb_lookup(const char *dbDir, uint32_t minValue, uint32_t maxValue, const uint64_t *kmerLo, const uint64_t *kmerHi,
              uint64_t nQueries, uint32_t *values, int32_t device, uint64_t workspaceBytes);

/*  Every k-mer of a database with its value, in ascending order (merylFileReader::nextMer, kmers-reader.C:456-531 -- the
 *  loop behind `meryl print`): blocks decoded on the device, lists handed to the callback in pieces (kmerHi NULL for k <= 32).
 *  A non-zero return of the callback stops the walk.  */
typedef int (*mb_kmer_callback)(void *arg, uint32_t k, const uint64_t *kmerLo, const uint64_t *kmerHi, const uint32_t *values, uint64_t n);
int mb_for_each_kmer(const char *dbDir, mb_kmer_callback cb, void *arg, int32_t device, uint64_t workspaceBytes);

/*  meryl-lookup's per-sequence numbers (src/meryl-lookup/existence.C:59-86 processSequence; include-exclude.C:62-97): sequence i is
 *  bases[seqStart[i] .. seqStart[i+1]) (seqStart[0] = 0, nSeqs + 1 entries; letters as in the file, any case; a k-mer never spans
 *  two sequences).  nTotal[i] = its k-mers (k consecutive ACGT letters); nFound[i] = those whose forward or reverse-complement form
 *  is in the database with a value in [minValue, maxValue] (merylExactLookup::value > 0 after load(minV, maxV)); the k-mers that start
 *  in the first skipFirst positions of a sequence are not looked up (-10x).  *nDbKmers (may be NULL) = merylExactLookup::nKmers.  */
int mb_existence(const char *dbDir, uint32_t minValue, uint32_t maxValue, const char *bases, const uint64_t *seqStart,
                 uint64_t nSeqs, uint32_t skipFirst, uint64_t *nTotal, uint64_t *nFound, uint64_t *nDbKmers,
                 int32_t device, uint64_t workspaceBytes);

/*  The per-position form behind meryl-lookup -bed / -wig-count / -wig-depth (src/meryl-lookup/dump.C:86-157 processSequence): for
 *  position p of the letters (same layout as mb_existence), flags[p] bit 0 = a k-mer starts at p, bit 1 = the database holds its
 *  forward or reverse-complement form (exists(c) || exists(n)); values[p] (may be NULL) = value(fmer) + value(rmer), counted once
 *  when the two forms are equal (dump.C:150-153).  As in the reference, value() is what the table stores: the k-mer's value less
 *  (minValue - 1) when minValue > 1 (kmers-exact.C:57-62; mb_lookup, in contrast, returns the k-mer's own value).  */
int mb_position_lookup(const char *dbDir, uint32_t minValue, uint32_t maxValue, const char *bases, const uint64_t *seqStart,
                       uint64_t nSeqs, uint8_t *flags, uint32_t *values, int32_t device, uint64_t workspaceBytes);

/*  The lookup table as an object (merylExactLookup: load() once, value() / exists() many times; kmers-lookup.H:83-146): the whole
 *  database is decoded into device memory and stays there; the queries above run against it.  mb_table_open fails with MB_ENOMEM
 *  when the decoded database does not fit the device (the one-shot calls, which work group of files by group, remain).
 *  mb_table_query = mb_existence + mb_position_lookup in one (any output may be NULL); mb_table_values = mb_lookup;
 *  mb_table_kmers = merylExactLookup::nKmers (the k-mers with a value in [minValue, maxValue]).  */
typedef struct mb_table mb_table;
int      mb_table_open(const char *dbDir, uint32_t minValue, uint32_t maxValue, int32_t device, uint64_t workspaceBytes, mb_table **out);
void     mb_table_close(mb_table *t);
uint32_t mb_table_k(const mb_table *t);
uint64_t mb_table_kmers(const mb_table *t);
int      mb_table_query(mb_table *t, const char *bases, const uint64_t *seqStart, uint64_t nSeqs, uint32_t skipFirst,
                        uint64_t *nTotal, uint64_t *nFound, uint8_t *flags, uint32_t *values);
int      mb_table_values(mb_table *t, const uint64_t *kmerLo, const uint64_t *kmerHi, uint64_t nQueries, uint32_t *values);

/* ---- stage-level entry points used by the parity tests (tests/test_gpu_stages.py) ------------- */

/*  K1+K2: host ASCII work-buffer -> all k-mers in stream order is not defined on the device (the
 *  scatter is unordered), so this returns them sorted ascending.  kmerHi may be NULL for k <= 32.  */
int mb_dbg_extract_sorted(uint32_t k, uint32_t mode, const char *bases, uint64_t len,
                          uint64_t *kmerLo, uint64_t *kmerHi, uint64_t cap, uint64_t *n);
/*  K4 alone: sort `n` keys of `bits` significant bits (keys[] in/out; hi may be NULL).  */
int mb_dbg_sort(uint64_t *lo, uint64_t *hi, uint64_t n, uint32_t bits);
/*  K4 on u32 keys (the suffix-key form a pass uses when 2k - binBits <= 32).  */
int mb_dbg_sort32(uint32_t *keys, uint64_t n, uint32_t bits);
/*  K6 alone: encode one block like merylFileWriter::writeBlockToFile.  out must hold cap bytes.  */
int mb_dbg_encode_block(uint32_t k, uint32_t wPrefix, uint64_t prefix, uint64_t n,
                        const uint64_t *sufLo, const uint64_t *sufHi, const uint32_t *values,
                        uint8_t *out, uint64_t cap, uint64_t *outLen);
/*  P1/P2 alone (host parser): file image -> base stream in work-buffer format.  */
int mb_dbg_parse(const uint8_t *data, uint64_t len, char *out, uint64_t cap, uint64_t *outLen);
/*  The same through the multi-threaded block parser the file inputs use (cuts at guessed record starts, verified).  */
int mb_dbg_parse_parallel(const uint8_t *data, uint64_t len, uint64_t block, char *out, uint64_t cap, uint64_t *outLen);
/*  Select the sort sweep kernel (1: 6-bit digits, 2: 8-bit digits) for contexts created afterwards; returns the previous one.  */
int mb_dbg_sort_version(int version);
/*  The count / merge kernels' warp-wide decoupled look-back on a prepared status array: *out = exclusive prefix of `tile`.  */
int mb_dbg_lookback(const uint64_t *status, uint64_t n, uint64_t tile, uint64_t *out);

/* ---- synthetic input for the benchmark and the tests (csrc/synth.cu) --------------------------- */

/*  Seeded reads from a virtual uniform-random genome: start uniform, strand 50/50,
 *  substitution rate err.  format 0: "<bases>." per read (work-buffer form); format 1: 4-line FASTQ
 *  "@r%012llu\n<bases>\n+\n<I...>\n".  Writes nReads * mb_synth_record_bytes() bytes of device memory.  */
uint64_t mb_synth_record_bytes(uint32_t readLen, uint32_t format);
int      mb_synth_reads_device(void *dOut, uint64_t firstRead, uint64_t nReads, uint32_t readLen,
                               uint64_t genomeLen, double err, uint64_t seed, uint32_t format, void *cudaStream);

#ifdef __cplusplus
}
#endif
#endif
