//  Device memory -> files, in parallel across files.  See filesink.h.
#include <errno.h>
#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "common.h"
#include "filesink.h"

namespace {
const size_t kSlotBytes = (size_t)8 << 20;

struct Job { int device, fd; uint64_t off; const uint8_t *src; size_t bytes; cudaEvent_t ready; };
}  //  namespace

struct MbFileSink {
  std::mutex mu;
  std::condition_variable cvWork, cvIdle;
  std::deque<Job> queue;
  std::set<int>   busyFd;
  size_t inFlight = 0;                                       //  queued + running
  int    errCode = MB_OK;
  std::string errText;
  std::vector<std::thread> workers;
  bool   stop = false;

  void fail(int code, const std::string &text) {
    std::lock_guard<std::mutex> lk(mu);
    if (errCode == MB_OK) { errCode = code; errText = text; }
  }

  void run(void) {
    uint8_t *slot[2] = { nullptr, nullptr };
    cudaEvent_t ev[2] = { nullptr, nullptr };
    cudaStream_t st = nullptr;
    int device = -1;
    for (;;) {
      Job j;
      {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
          if (stop) goto out;
          bool got = false;
          for (auto it = queue.begin(); it != queue.end(); ++it)
            if (!busyFd.count(it->fd)) { j = *it; queue.erase(it); busyFd.insert(j.fd); got = true; break; }
          if (got) break;
          cvWork.wait(lk);
        }
      }
      {
        bool ok = true;
        std::string why;
        auto cu = [&](cudaError_t e, const char *what) { if (ok && e != cudaSuccess) { ok = false; why = std::string(what) + ": " + cudaGetErrorString(e); } };
        if (device != j.device) {
          cu(cudaSetDevice(j.device), "cudaSetDevice");
          if (st) { cudaStreamDestroy(st); st = nullptr; }
          device = j.device;
        }
        if (ok && !st) cu(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), "cudaStreamCreate");
        for (int i = 0; i < 2 && ok; i++) {
          if (!slot[i]) cu(cudaMallocHost((void **)&slot[i], kSlotBytes), "cudaMallocHost");
          if (ok && !ev[i]) cu(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming), "cudaEventCreate");
        }
        if (ok && j.ready) cu(cudaStreamWaitEvent(st, j.ready, 0), "cudaStreamWaitEvent");
        const size_t nPieces = (j.bytes + kSlotBytes - 1) / kSlotBytes;
        auto issue = [&](size_t p) {
          const size_t o = p * kSlotBytes, len = std::min(kSlotBytes, j.bytes - o);
          cu(cudaMemcpyAsync(slot[p & 1], j.src + o, len, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync");
          cu(cudaEventRecord(ev[p & 1], st), "cudaEventRecord");
        };
        if (ok && nPieces) issue(0);
        int wcode = MB_OK;
        for (size_t p = 0; p < nPieces && ok; p++) {
          if (p + 1 < nPieces) issue(p + 1);
          cu(cudaEventSynchronize(ev[p & 1]), "cudaEventSynchronize");
          if (!ok) break;
          const size_t o = p * kSlotBytes, len = std::min(kSlotBytes, j.bytes - o);
          size_t put = 0;
          while (put < len) {
            const ssize_t r = pwrite(j.fd, slot[p & 1] + put, len - put, (off_t)(j.off + o + put));
            if (r <= 0) { ok = false; wcode = MB_EIO; why = std::string("pwrite: ") + strerror(errno); break; }
            put += (size_t)r;
          }
        }
        if (!ok) { if (st) cudaStreamSynchronize(st); fail(wcode != MB_OK ? wcode : MB_ENODEVICE, why); }
      }
      {
        std::lock_guard<std::mutex> lk(mu);
        busyFd.erase(j.fd);
        inFlight--;
      }
      cvWork.notify_all();                                   //  (a job of the same file may have been waiting)
      cvIdle.notify_all();
    }
  out:
    for (int i = 0; i < 2; i++) { if (slot[i]) cudaFreeHost(slot[i]); if (ev[i]) cudaEventDestroy(ev[i]); }
    if (st) cudaStreamDestroy(st);
  }
};

MbFileSink *mb_filesink_get(void) {
  static MbFileSink *S = nullptr;
  static std::mutex m;
  std::lock_guard<std::mutex> lk(m);
  if (!S) {
    S = new MbFileSink();                                    //  lives until the process ends (detached workers)
    unsigned n = std::thread::hardware_concurrency();
    if (getenv("MB_WRITE_THREADS")) n = (unsigned)atoi(getenv("MB_WRITE_THREADS"));
    if (n > 16) n = 16;
    if (n < 2)  n = 2;
    MbFileSink *self = S;
    for (unsigned i = 0; i < n; i++) { S->workers.emplace_back([self]() { self->run(); }); S->workers.back().detach(); }
  }
  return S;
}

void mb_filesink_submit(MbFileSink *S, int device, int fd, uint64_t fileOff, const uint8_t *dSrc, size_t bytes, cudaEvent_t ready) {
  if (bytes == 0) return;
  {
    std::lock_guard<std::mutex> lk(S->mu);
    Job j;
    j.device = device;  j.fd = fd;  j.off = fileOff;  j.src = dSrc;  j.bytes = bytes;  j.ready = ready;
    S->queue.push_back(j);
    S->inFlight++;
  }
  S->cvWork.notify_one();
}

int mb_filesink_drain(MbFileSink *S) {
  std::unique_lock<std::mutex> lk(S->mu);
  S->cvIdle.wait(lk, [&]() { return S->inFlight == 0; });
  if (S->errCode == MB_OK) return MB_OK;
  const int code = S->errCode;
  const std::string text = S->errText;
  S->errCode = MB_OK;  S->errText.clear();
  lk.unlock();
  return mb_fail(code, "writing the database: %s", text.c_str());
}
