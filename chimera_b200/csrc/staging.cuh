// staging.cuh -- host<->device copies of PAGEABLE host buffers at PCIe speed.
//
// The drop-in boundary (include/chimera_b200.h) receives plain numpy-owned pointers: pageable memory.  A
// cudaMemcpyAsync from/to pageable memory is staged by the driver through one internal buffer on one thread
// (~10 GB/s measured here, profiles/r01g bench `e2e_per_call`), far below the 55 GB/s the link gives from
// page-locked memory.  The Stager splits a copy into (cache-sized) chunks, moves each chunk between the user buffer and a
// ring of page-locked slots with a small pool of host threads, and lets the DMA engine work on the previous
// slot meanwhile (events guard slot reuse).  Already page-locked buffers (cudaHostRegister / cudaHostAlloc) and
// small copies bypass it.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#include "common.cuh"

namespace chb {

class Stager {
 public:
  static constexpr size_t kChunkDefault = size_t(4) << 20;  // bytes per slot (CHIMERA_STAGE_CHUNK_MB overrides); measured
                                                            // 2.4 GB up + 1.2 GB down: 4 MB 51.7, 16 MB 40.7, 64 MB 37.9 GB/s
  static constexpr int kSlots = 6;
  static constexpr size_t kMinBytes = size_t(4) << 20;  // below this the plain copy is as fast

  // both calls are ordered on `st` like a cudaMemcpyAsync; h2d returns once the last chunk is enqueued,
  // d2h returns once the data is in `dst` (it waits for earlier work on `st`)
  int h2d(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t st);
  int d2h(void* dst_host, const void* src_dev, size_t bytes, cudaStream_t st);
  static bool pageable(const void* host_ptr);
  ~Stager();

 private:
  int init();
  void parallel_copy(char* dst, const char* src, size_t bytes);
  void worker(int id);

  bool ready_ = false;
  size_t chunk_ = kChunkDefault;
  char* slot_[kSlots] = {};
  cudaEvent_t ev_[kSlots] = {};
  bool busy_[kSlots] = {};  // an event has been recorded for the slot and not yet waited for
  // thread pool: one job at a time = copy [src, src+bytes) -> dst split evenly over the workers + the caller
  std::vector<std::thread> pool_;
  std::mutex mu_;
  std::condition_variable cv_job_, cv_done_;
  unsigned long long job_id_ = 0;
  int pending_ = 0;
  bool stop_ = false;
  char* j_dst_ = nullptr;
  const char* j_src_ = nullptr;
  size_t j_bytes_ = 0;
  int nparts_ = 1;
};

}  // namespace chb
