//  Engine-internal interfaces shared by engine.cu, exchange.cu and debug.cu.
#pragma once

#include <vector>
#include <cuda_runtime.h>

#include "common.h"

struct MbComm;                                        //  exchange.cu: NCCL communicator + ownership map
void mb_comm_destroy(MbComm *m);

void                     mb_ctx_set_stop_after_sort(mb_ctx *c);
const std::vector<u64>  &mb_ctx_sorted_lo(mb_ctx *c);
const std::vector<u64>  &mb_ctx_sorted_hi(mb_ctx *c);
MbDbImages              &mb_ctx_images(mb_ctx *c);
void                     mb_ctx_set_comm(mb_ctx *c, MbComm *m);
MbComm                  *mb_ctx_comm(mb_ctx *c);
cudaStream_t             mb_ctx_stream(mb_ctx *c);
int                      mb_ctx_device(mb_ctx *c);
const mb_config         &mb_ctx_config(mb_ctx *c);
