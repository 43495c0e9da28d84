//  Multi-GPU routing of extracted k-mers to the GPU that owns their data file (NCCL all-to-all).
//  Placeholder until the single-GPU path is parity-green on hardware; see DESIGN.md (e).
#include "common.h"
#include "engine.h"

struct MbComm { int unused; };

void mb_comm_destroy(MbComm *m) { delete m; }

extern "C" int mb_comm_unique_id(uint8_t id[128]) {
  mb_clear_error();
  (void)id;
  return mb_fail(MB_ECOMM, "multi-GPU exchange is not built yet");
}

extern "C" int mb_comm_init(mb_ctx *ctx, const uint8_t id[128]) {
  mb_clear_error();
  (void)ctx; (void)id;
  return mb_fail(MB_ECOMM, "multi-GPU exchange is not built yet");
}
