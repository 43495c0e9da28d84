//  P7: prefix-width planner.  The db bytes depend on wPrefix, so this must agree with the
//  reference for every (k, n, memory):
//    simple-mode size        src/meryl/merylOp-count.C:119-165  (findExpectedSimpleSize)
//    complex-mode search     src/meryl/merylOp-count.C:174-227  (findBestPrefixSize)
//    batching loop/decision  src/meryl/merylOp-count.C:346-382  (configureCounting)
//    simple-mode prefix      src/meryl/merylOp-countSimple.C:196-199
//    input size guess        src/meryl/merylOp-count.C:410-433
//  The GPU engine itself needs none of the reference's memory model (its passes are planned
//  from the measured prefix histogram); the model is evaluated only to land on the same wPrefix.

#include <string.h>
#include <sys/stat.h>

#include "common.h"

namespace {

const u64 kBucketStruct = 3232;     //  sizeof(merylCountArray) in the reference build (instrumented)
const u64 kSegBytes     = 4096;     //  one page per segment
const u64 kSegBits      = kSegBytes * 8;

inline u32 bitsNeeded(u64 x) { return x ? 64 - __builtin_clzll(x) : 0; }

//  Smallest-memory prefix width for n k-mers; 0 / UINT64_MAX if none is acceptable.
void searchPrefix(u32 k, u64 n, u64 memAllowed, u32 &best, u64 &bestMem) {
  best    = 0;
  bestMem = UINT64_MAX;

  for (u32 wp = 1; wp < 2 * k - 1; wp++) {
    const u64 buckets       = (u64)1 << wp;
    const u64 perBucket     = n / buckets + 1;
    const u64 perSegment    = kSegBits / (2 * k - wp);
    const u64 segsPerBucket = perBucket / perSegment + 1;

    if (wp + bitsNeeded(segsPerBucket) + bitsNeeded(kSegBytes) >= 64)
      break;

    const u64 fixedMem = kBucketStruct * buckets + sizeof(u64 *) * buckets * segsPerBucket;
    const u64 minData  = buckets * kSegBytes;
    const u64 total    = fixedMem + buckets * segsPerBucket * kSegBytes;

    if (fixedMem + minData > memAllowed)
      break;

    if (wp > 9 && total + (u64)16 * wp * 1024 * 1024 < bestMem) {
      bestMem = total;
      best    = wp;
    }

    if (total > 16 * bestMem)
      break;
  }
}

}  //  namespace


int mb_plan_impl(u32 k, u64 nEstimate, u64 memBytes, MbPlan *out) {
  if (k < 3 || k > 64)
    return mb_fail(MB_EINVAL, "mb_plan: k=%u outside 3..64", k);

  u64 simpleMem = UINT64_MAX;
  if (2 * k <= 42) {
    const u64 entries  = (u64)1 << (2 * k);
    const u64 maxCount = (u64)(0.004 * nEstimate);
    const u64 cntBits  = bitsNeeded(maxCount) + 1;
    const u64 extra    = cntBits < 16 ? 0 : cntBits - 16;
    simpleMem = (entries * 16 + entries * extra) / 8;
  }

  u32 best = 0, batches = 1;
  u64 complexMem = UINT64_MAX;
  if (k > 5) {
    for (batches = 1; complexMem > memBytes; batches++) {
      searchPrefix(k, nEstimate / batches, memBytes, best, complexMem);
      if (batches > 1000000)
        return mb_fail(MB_EINVAL, "mb_plan: no prefix width fits memory=%llu bytes", (u64)memBytes);
    }
  }

  out->simple = 0;
  if (simpleMem < complexMem && simpleMem < memBytes) {
    const u32 ps = 2 * k - 6;
    const u32 ws = ps > 20 ? 20 : ps;
    best        = 6 + ps - ws;
    complexMem  = simpleMem;
    batches     = 1;
    out->simple = 1;
  }

  if (best == 0)
    return mb_fail(MB_EINVAL, "mb_plan: no counting mode fits k=%u memory=%llu bytes", k, (u64)memBytes);

  out->wPrefix  = best;
  out->memUsed  = complexMem;
  out->nBatches = batches;
  return MB_OK;
}


extern "C" int mb_plan(uint32_t k, uint64_t nEstimate, uint64_t memBytes,
                       uint32_t *wPrefix, int *simpleMode, uint64_t *memUsed, uint32_t *nBatches) {
  mb_clear_error();
  MbPlan p;
  MB_TRY(mb_plan_impl(k, nEstimate, memBytes, &p));
  if (wPrefix)    *wPrefix    = p.wPrefix;
  if (simpleMode) *simpleMode = p.simple;
  if (memUsed)    *memUsed    = p.memUsed;
  if (nBatches)   *nBatches   = p.nBatches;
  return MB_OK;
}


extern "C" uint64_t mb_estimate_kmers_in_file(const char *path) {
  if (path == NULL || strcmp(path, "-") == 0)
    return 0;
  struct stat st;
  if (stat(path, &st) != 0)
    return 0;
  const u64    size = (u64)st.st_size;
  const size_t len  = strlen(path);
  if (len > 3 && strcmp(path + len - 3, ".gz")  == 0)  return size * 3;
  if (len > 4 && strcmp(path + len - 4, ".bz2") == 0)  return (u64)(size * 3.5);
  if (len > 3 && strcmp(path + len - 3, ".xz")  == 0)  return (u64)(size * 4.0);
  return size;
}
