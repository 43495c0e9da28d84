//  Device memory -> files, in parallel ACROSS files (filesink.cpp).
//
//  Replaces, for the 64 data files, the fwrite of merylFileWriter::writeBlockToFile / merylBlockWriter::closeFileDumpIndex
//  (src/utility/src/kmers-v1/kmers-writer.C:183-291, kmers-writer-block.C:134-156), where the reference gets its
//  parallelism from one OpenMP thread per file.  Measured here (tmpfs, 8 cores): one file takes 2.1 GB/s from one thread
//  and no more from eight (buffered writes hold the inode lock), eight files take 13 GB/s from eight threads -- so a
//  job is one (file, byte range) and a worker streams it alone: device -> its two page-locked slots (its own CUDA
//  stream, the next piece in flight while one is written) -> pwrite.  Two jobs of one file never run at the same time.
#pragma once

#include <stdint.h>
#include <stddef.h>
#include <cuda_runtime.h>

struct MbFileSink;
MbFileSink *mb_filesink_get(void);                           //  the process-wide pool (created on first use)
//  Copy [dSrc, dSrc + bytes) of device `device` to file descriptor fd at fileOff.  `ready` (may be null): event after which
//  dSrc holds the bytes.  dSrc must stay untouched until mb_filesink_drain returns.
void mb_filesink_submit(MbFileSink *S, int device, int fd, uint64_t fileOff, const uint8_t *dSrc, size_t bytes, cudaEvent_t ready);
int  mb_filesink_drain(MbFileSink *S);                       //  all submitted jobs done; MB_OK or the first error (mb_fail)
