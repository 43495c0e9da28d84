//  Multi-GPU exchange layer (exchange.cu).
#pragma once
#include <cuda_runtime.h>
#include "common.h"

struct MbComm;
u32 mb_comm_rank(const MbComm *m);
u32 mb_comm_world(const MbComm *m);
u32 mb_file_owner(u32 file, u32 world);
int mb_comm_allgather_u64(MbComm *m, const u64 *dMine, u64 *dAll, size_t n, cudaStream_t s);
int mb_comm_exchange_keys(MbComm *m, u32 keyWords, const void *sendBuf, void *recvBuf,
                          u32 fileLo, u32 fileHi, const u64 *sendStart, const u64 *C, cudaStream_t s);
int mb_comm_exchange_suffixes(MbComm *m, const void *sendBuf, void *recvBuf, u32 binLo, u32 binHi, u32 binsPerFile,
                              const u64 *sendStart, const u64 *Cb, u32 nBins, cudaStream_t s);
int mb_comm_gather_u64(MbComm *m, const u64 *dSend, u64 n, u64 *dRecv, const u64 *counts, cudaStream_t s);
int mb_comm_open_peer_arenas(MbComm *m, void *myArena, u64 *dScratch, uint8_t **peerBase, cudaStream_t s);
int mb_comm_close_peer_arenas(MbComm *m, uint8_t **peerBase);
