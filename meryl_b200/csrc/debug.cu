//  Stage-level entry points for the parity tests: each runs ONE device stage on caller data so a
//  failure points at a kernel, not at the pipeline.  (tests/test_gpu_stages.py)
#include <string.h>
#include <vector>

#include "common.h"
#include "kernels.cuh"
#include "engine.h"

extern "C" int mb_dbg_parse(const uint8_t *data, uint64_t len, char *out, uint64_t cap, uint64_t *outLen) {
  mb_clear_error();
  if (!outLen) return mb_fail(MB_EINVAL, "mb_dbg_parse: NULL argument");
  if (!SeqParser::sniff(data, len < 65536 ? len : 65536)) return mb_fail(MB_EFORMAT, "input is not FASTA or FASTQ");
  struct Acc { std::string s; } acc;
  SeqParser ps([](void *a, const char *b, size_t n) { ((Acc *)a)->s.append(b, n); return 0; }, &acc);
  //  feed in awkward pieces to exercise the restartable state machine
  size_t at = 0, step = 1;
  while (at < len) {
    const size_t n = (len - at < step) ? (len - at) : step;
    MB_TRY(ps.feed(data + at, n));
    at += n;
    step = step * 3 + 1;
    if (step > 100000) step = 7;
  }
  MB_TRY(ps.finish());
  *outLen = acc.s.size();
  if (out) {
    if (acc.s.size() > cap) return mb_fail(MB_EINVAL, "mb_dbg_parse: output needs %zu bytes", acc.s.size());
    memcpy(out, acc.s.data(), acc.s.size());
  }
  return MB_OK;
}

extern "C" int mb_dbg_extract_sorted(uint32_t k, uint32_t mode, const char *bases, uint64_t len,
                                     uint64_t *kmerLo, uint64_t *kmerHi, uint64_t cap, uint64_t *n) {
  mb_clear_error();
  if (!n) return mb_fail(MB_EINVAL, "mb_dbg_extract_sorted: NULL argument");
  mb_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.k = k;  cfg.mode = mode;  cfg.device = -1;
  cfg.wPrefix = 2 * k < 10 ? 2 * k : 10;
  cfg.workspaceBytes = (size_t)2 << 30;
  mb_ctx *c = nullptr;
  MB_TRY(mb_create(&cfg, &c));
  mb_ctx_set_stop_after_sort(c);
  int rc = mb_add_bases(c, bases, len, 1);
  if (rc == MB_OK) rc = mb_count(c);
  if (rc == MB_OK) {
    const std::vector<u64> &lo = mb_ctx_sorted_lo(c), &hi = mb_ctx_sorted_hi(c);
    *n = lo.size();
    if (kmerLo) {
      if (lo.size() > cap) rc = mb_fail(MB_EINVAL, "mb_dbg_extract_sorted: %zu k-mers, room for %llu", lo.size(), (u64)cap);
      else {
        memcpy(kmerLo, lo.data(), lo.size() * 8);
        if (kmerHi) { if (k > 32) memcpy(kmerHi, hi.data(), hi.size() * 8); else memset(kmerHi, 0, lo.size() * 8); }
      }
    }
  }
  char keep[1024];
  strncpy(keep, mb_last_error(), sizeof(keep) - 1);  keep[sizeof(keep) - 1] = 0;
  mb_destroy(c);
  if (rc != MB_OK) mb_fail(rc, "%s", keep);
  return rc;
}

extern "C" int mb_dbg_sort(uint64_t *lo, uint64_t *hi, uint64_t n, uint32_t bits) {
  mb_clear_error();
  if (!lo && n) return mb_fail(MB_EINVAL, "mb_dbg_sort: NULL argument");
  const u32 kw = hi ? 2 : 1;
  const u32 W  = 8 * kw;
  MbSortPlan plan;
  mb_sort_plan(bits, &plan);
  const u32 tileKeys = mb_sort_tile_keys(W);
  const u64 nTiles   = (n + tileKeys - 1) / tileKeys;
  const size_t wsBytes = mb_sort_workspace_bytes(nTiles, 1, plan.nPasses);

  void *a = nullptr, *b = nullptr, *ws = nullptr;
  u64  *seg = nullptr;
  MB_CUDA(cudaMalloc(&a, W * n + 64));
  MB_CUDA(cudaMalloc(&b, W * n + 64));
  MB_CUDA(cudaMalloc(&ws, wsBytes + 64));
  MB_CUDA(cudaMalloc((void **)&seg, 4 * sizeof(u64)));
  std::vector<u64> host(kw * n);
  for (u64 i = 0; i < n; i++) { host[kw * i] = lo[i]; if (hi) host[kw * i + 1] = hi[i]; }
  const u64 segs[4] = { 0, n, 0, nTiles };
  MB_CUDA(cudaMemcpy(a, host.data(), W * n, cudaMemcpyHostToDevice));
  MB_CUDA(cudaMemcpy(seg, segs, sizeof(segs), cudaMemcpyHostToDevice));
  int inB = 0;
  plan.maxSegKeys = (n & 1) ? n : 0;                    //  exercise both the 32-bit and the 64-bit output offsets
  MB_CUDA(mb_launch_sort(W, a, b, n, seg, seg + 2, 1, nTiles, plan, ws, wsBytes, &inB, nullptr, 0, nullptr));
  MB_CUDA(cudaDeviceSynchronize());
  MB_CUDA(cudaMemcpy(host.data(), inB ? b : a, W * n, cudaMemcpyDeviceToHost));
  for (u64 i = 0; i < n; i++) { lo[i] = host[kw * i]; if (hi) hi[i] = host[kw * i + 1]; }
  cudaFree(a); cudaFree(b); cudaFree(ws); cudaFree(seg);
  return MB_OK;
}

//  K4 on u32 keys (the suffix-key path used for k <= 21).
extern "C" int mb_dbg_sort32(uint32_t *keys, uint64_t n, uint32_t bits) {
  mb_clear_error();
  if (!keys && n) return mb_fail(MB_EINVAL, "mb_dbg_sort32: NULL argument");
  if (bits > 32)  return mb_fail(MB_EINVAL, "mb_dbg_sort32: bits > 32");
  MbSortPlan plan;
  mb_sort_plan(bits, &plan);
  const u32 tileKeys = mb_sort_tile_keys(4);
  const u64 nTiles   = (n + tileKeys - 1) / tileKeys;
  const size_t wsBytes = mb_sort_workspace_bytes(nTiles, 1, plan.nPasses);
  void *a = nullptr, *b = nullptr, *ws = nullptr;
  u64  *seg = nullptr;
  MB_CUDA(cudaMalloc(&a, 4 * n + 64));
  MB_CUDA(cudaMalloc(&b, 4 * n + 64));
  MB_CUDA(cudaMalloc(&ws, wsBytes + 64));
  MB_CUDA(cudaMalloc((void **)&seg, 4 * sizeof(u64)));
  const u64 segs[4] = { 0, n, 0, nTiles };
  MB_CUDA(cudaMemcpy(a, keys, 4 * n, cudaMemcpyHostToDevice));
  MB_CUDA(cudaMemcpy(seg, segs, sizeof(segs), cudaMemcpyHostToDevice));
  int inB = 0;
  plan.maxSegKeys = (n & 1) ? n : 0;
  MB_CUDA(mb_launch_sort(4, a, b, n, seg, seg + 2, 1, nTiles, plan, ws, wsBytes, &inB, nullptr, 0, nullptr));
  MB_CUDA(cudaDeviceSynchronize());
  MB_CUDA(cudaMemcpy(keys, inB ? b : a, 4 * n, cudaMemcpyDeviceToHost));
  cudaFree(a); cudaFree(b); cudaFree(ws); cudaFree(seg);
  return MB_OK;
}

extern "C" int mb_dbg_encode_block(uint32_t k, uint32_t wPrefix, uint64_t prefix, uint64_t n,
                                   const uint64_t *sufLo, const uint64_t *sufHi, const uint32_t *values,
                                   uint8_t *out, uint64_t cap, uint64_t *outLen) {
  mb_clear_error();
  if (!outLen || (n && (!sufLo || !values))) return mb_fail(MB_EINVAL, "mb_dbg_encode_block: NULL argument");
  const u32 kw = k <= 32 ? 1 : 2;
  const u32 wS = 2 * k - wPrefix;

  //  keys = (prefix << wS) | suffix;  runStart = running sum of the values
  std::vector<u64> keys(kw * (n ? n : 1)), rs(n + 1);
  typedef unsigned __int128 u128;
  for (u64 i = 0; i < n; i++) {
    u128 s = ((u128)(sufHi ? sufHi[i] : 0) << 64) | sufLo[i];
    u128 v = (wS >= 128 ? 0 : ((u128)prefix << wS)) | s;
    keys[kw * i] = (u64)v;
    if (kw == 2) keys[kw * i + 1] = (u64)(v >> 64);
  }
  rs[0] = 0;
  for (u64 i = 0; i < n; i++) rs[i + 1] = rs[i] + values[i];

  void *dKeys = nullptr;  u64 *dRs = nullptr, *dTot = nullptr, *dFileBase = nullptr;  MbBlockMeta *dMeta = nullptr;  void *dTmp = nullptr;
  MB_CUDA(cudaMalloc(&dKeys, keys.size() * 8 + 64));
  MB_CUDA(cudaMalloc((void **)&dRs, (n + 1) * 8));
  MB_CUDA(cudaMalloc((void **)&dTot, 4 * 8));
  MB_CUDA(cudaMalloc((void **)&dFileBase, 64 * 8));
  MB_CUDA(cudaMalloc((void **)&dMeta, sizeof(MbBlockMeta)));
  MB_CUDA(cudaMalloc(&dTmp, mb_encode_scan_tmp_bytes(1) + 64));
  MB_CUDA(cudaMemcpy(dKeys, keys.data(), keys.size() * 8, cudaMemcpyHostToDevice));
  MB_CUDA(cudaMemcpy(dRs, rs.data(), (n + 1) * 8, cudaMemcpyHostToDevice));
  MB_CUDA(cudaMemset(dFileBase, 0, 64 * 8));
  MB_CUDA(cudaMemset(dTot, 0, 4 * 8));

  MbEncodeParams ep;
  ep.keys = dKeys;  ep.runStart = dRs;  ep.nDistinct = n;  ep.keyWords = kw;  ep.k = k;  ep.wPrefix = wPrefix;  ep.counts = nullptr;  ep.streamMode = 0;
  ep.blockLo = prefix;  ep.nBlocks = 1;
  MB_CUDA(mb_launch_encode_layout(ep, dMeta, dTot, dTmp, 0));
  u64 tot[4];
  MB_CUDA(cudaMemcpy(tot, dTot, sizeof(tot), cudaMemcpyDeviceToHost));

  uint8_t *dOut = nullptr;  u32 *dSub = nullptr;  u64 *dIdx = nullptr;
  MB_CUDA(cudaMalloc((void **)&dOut, tot[0] + 64));
  MB_CUDA(cudaMalloc((void **)&dSub, tot[1] * 4 + 64));
  MB_CUDA(cudaMalloc((void **)&dIdx, 24));
  MB_CUDA(cudaMemset(dOut, 0, tot[0]));
  MbEncodeChunk ch = { 0, 1, 0, 0 };
  MB_CUDA(mb_launch_encode_write(ep, ch, dMeta, dTot, dSub, dOut, dIdx, dFileBase, 0));
  MB_CUDA(cudaDeviceSynchronize());

  *outLen = tot[0];
  int rc = MB_OK;
  if (out) {
    if (tot[0] > cap) rc = mb_fail(MB_EINVAL, "mb_dbg_encode_block: block needs %llu bytes", (u64)tot[0]);
    else MB_CUDA(cudaMemcpy(out, dOut, tot[0], cudaMemcpyDeviceToHost));
  }
  cudaFree(dKeys); cudaFree(dRs); cudaFree(dTot); cudaFree(dFileBase); cudaFree(dMeta); cudaFree(dTmp);
  cudaFree(dOut); cudaFree(dSub); cudaFree(dIdx);
  return rc;
}

//  Select the sweep shape (2 or 3, see sort.cu) for the contexts created afterwards; returns the previous one.
extern "C" int mb_dbg_sort_version(int v) {
  const int old = mb_sort_version();
  if (v == 2 || v == 3) mb_sort_set_version(v);
  return old;
}

//  The count kernels' warp-wide decoupled look-back on a caller-made status array (words: flag << 62 | value; flag 1 =
//  aggregate, 2 = inclusive prefix; every word below `tile` must be published).  *out = exclusive prefix of `tile`.
extern "C" int mb_dbg_lookback(const uint64_t *status, uint64_t n, uint64_t tile, uint64_t *out) {
  mb_clear_error();
  if (!status || !out || tile > n) return mb_fail(MB_EINVAL, "mb_dbg_lookback: bad argument");
  u64 *dSt = nullptr, *dOut = nullptr;
  MB_CUDA(cudaMalloc((void **)&dSt, (n + 1) * 8));
  MB_CUDA(cudaMalloc((void **)&dOut, 8));
  MB_CUDA(cudaMemcpy(dSt, status, n * 8, cudaMemcpyHostToDevice));
  MB_CUDA(mb_launch_lookback_test(dSt, tile, dOut, 0));
  MB_CUDA(cudaDeviceSynchronize());
  MB_CUDA(cudaMemcpy(out, dOut, 8, cudaMemcpyDeviceToHost));
  cudaFree(dSt);  cudaFree(dOut);
  return MB_OK;
}

//  P1/P2 through the multi-threaded block parser (mb_parse_parallel): the file image in blocks of `block` bytes.
extern "C" int mb_dbg_parse_parallel(const uint8_t *data, uint64_t len, uint64_t block, char *out, uint64_t cap, uint64_t *outLen) {
  mb_clear_error();
  if (!outLen || block == 0) return mb_fail(MB_EINVAL, "mb_dbg_parse_parallel: bad argument");
  if (!SeqParser::sniff(data, len < 65536 ? len : 65536)) return mb_fail(MB_EFORMAT, "input is not FASTA or FASTQ");
  std::string acc;
  auto sink = [](void *a, const char *b, size_t n) -> int { ((std::string *)a)->append(b, n); return 0; };
  SeqParser ps(sink, &acc);
  for (uint64_t at = 0; at < len; at += block) MB_TRY(mb_parse_parallel(sink, &acc, ps, data + at, (size_t)(len - at < block ? len - at : block)));
  MB_TRY(ps.finish());
  *outLen = acc.size();
  if (out) {
    if (acc.size() > cap) return mb_fail(MB_EINVAL, "mb_dbg_parse_parallel: output needs %zu bytes", acc.size());
    memcpy(out, acc.data(), acc.size());
  }
  return MB_OK;
}

#ifdef MB_CUSIM
//  Emulator only (tests/cusim): `world` ranks as fibers of this process, each with its own context and its share of the
//  input (file images), the emulated collectives of exchange.cu between them; every rank writes its files into outDir.
extern "C" int mb_sim_comm_attach(mb_ctx *ctx);
static uint32_t g_simKeyBytes[32];
extern "C" uint32_t mb_sim_last_key_bytes(uint32_t rank) { return rank < 32 ? g_simKeyBytes[rank] : 0; }   //  of the last mb_sim_multi_count
extern "C" int mb_sim_multi_count(uint32_t world, uint32_t k, uint32_t mode, uint32_t wPrefix, const uint8_t *const *data, const uint64_t *lens,
                                  const char *outDir, uint64_t *nKmersPerRank, uint32_t *nPassesPerRank) {
  mb_clear_error();
  std::vector<int> rcs(world, MB_OK);
  std::vector<std::string> errs(world);
  cusim::run_ranks(world, [&](unsigned r) {
    mb_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.k = k;  cfg.mode = mode;  cfg.wPrefix = wPrefix;  cfg.device = -1;  cfg.rank = r;  cfg.worldSize = world;
    cfg.workspaceBytes = (size_t)384 << 20;
    mb_ctx *c = nullptr;
    int rc = mb_create(&cfg, &c);
    if (rc == MB_OK) rc = mb_sim_comm_attach(c);
    if (rc == MB_OK && (r & 1)) rc = mb_set_output_dir(c, outDir);          //  odd ranks stream their data files, even ranks keep images
    if (rc == MB_OK && lens[r]) rc = mb_add_file_bytes(c, data[r], lens[r]);
    if (rc == MB_OK) rc = mb_count(c);
    if (rc == MB_OK) rc = mb_write_db(c, outDir);
    mb_stats st;
    if (rc == MB_OK && mb_get_stats(c, &st) == MB_OK) { if (nKmersPerRank) nKmersPerRank[r] = st.nKmers; if (nPassesPerRank) nPassesPerRank[r] = st.nPasses; if (r < 32) g_simKeyBytes[r] = st.keyBytes; }
    if (rc != MB_OK) errs[r] = mb_last_error();
    rcs[r] = rc;
    if (c) mb_destroy(c);
  });
  for (uint32_t r = 0; r < world; r++) if (rcs[r] != MB_OK) return mb_fail(rcs[r], "rank %u: %s", r, errs[r].c_str());
  return MB_OK;
}
#endif
