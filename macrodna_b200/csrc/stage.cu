// Host-side staging of PAGEABLE input matrices (what a caller holds after pd.read_csv / df.to_numpy()).
// cudaMemcpyAsync from pageable memory goes through the driver's own single-threaded bounce buffer (measured:
// ~11 GB/s on the B200 box, 0.87 s for the 9.6 GB of the large config -- more than the whole device path).  Here a
// small pool of host threads copies row blocks into two pinned buffers while the previous block is on the wire, so
// the transfer runs at min(host memcpy bandwidth, PCIe) and stays asynchronous to the compute stream.
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>

#include "mcd_internal.cuh"

struct mcd_stager {
  static constexpr size_t kStageBytes = (size_t)64 << 20;
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  uint64_t generation = 0;
  int pending = 0;
  bool quit = false;
  // current job: copy `rows` rows of `width_bytes` from src (pitch src_pitch) to dst (dense)
  const char* src = nullptr;
  char* dst = nullptr;
  size_t width_bytes = 0, src_pitch = 0;
  int64_t rows = 0;

  explicit mcd_stager(int nthreads) {
    for (int t = 0; t < nthreads; ++t) workers.emplace_back([this, t, nthreads] { run(t, nthreads); });
  }
  ~mcd_stager() {
    {
      std::lock_guard<std::mutex> lk(mu);
      quit = true;
      ++generation;
    }
    cv_job.notify_all();
    for (auto& w : workers) w.join();
  }
  void slice(int t, int nthreads) const {
    const int64_t per = (rows + nthreads - 1) / nthreads;
    const int64_t lo = t * per < rows ? t * per : rows, hi = lo + per < rows ? lo + per : rows;
    if (src_pitch == width_bytes) {
      if (hi > lo) memcpy(dst + lo * width_bytes, src + lo * src_pitch, (size_t)(hi - lo) * width_bytes);
    } else {
      for (int64_t r = lo; r < hi; ++r) memcpy(dst + r * width_bytes, src + r * src_pitch, width_bytes);
    }
  }
  void run(int t, int nthreads) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_job.wait(lk, [&] { return generation != seen; });
        seen = generation;
        if (quit) return;
      }
      slice(t, nthreads);
      {
        std::lock_guard<std::mutex> lk(mu);
        if (--pending == 0) cv_done.notify_one();
      }
    }
  }
  void copy(char* d, const char* s, size_t wbytes, size_t pitch, int64_t nrows) {
    std::unique_lock<std::mutex> lk(mu);
    src = s;
    dst = d;
    width_bytes = wbytes;
    src_pitch = pitch;
    rows = nrows;
    pending = (int)workers.size();
    ++generation;
    cv_job.notify_all();
    cv_done.wait(lk, [&] { return pending == 0; });
  }
};

void mcd_stager_destroy(mcd_stager* s) { delete s; }

bool mcd_is_pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

int mcd_staged_h2d(mcd_context* h, double* dst, int64_t dst_ld, const double* src, int64_t src_ld, int64_t width,
                   int64_t rows, cudaStream_t stream) {
  if (rows <= 0 || width <= 0) return MCD_OK;
  if (h->stager == nullptr) {
    int nt = (int)std::thread::hardware_concurrency();
    nt = nt < 2 ? 2 : (nt > 8 ? 8 : nt);
    h->stager = new (std::nothrow) mcd_stager(nt);
    if (h->stager == nullptr) return mcd_fail(h, MCD_ERR_NOMEM, "staging threads");
  }
  for (int b = 0; b < 2; ++b) {
    if (h->h_stage[b] == nullptr) {
      MCD_CUDA(h, cudaMallocHost(&h->h_stage[b], mcd_stager::kStageBytes));
      MCD_CUDA(h, cudaEventCreateWithFlags(&h->stage_ev[b], cudaEventDisableTiming));
    }
  }
  const size_t wbytes = (size_t)width * 8;
  int64_t rows_per = (int64_t)(mcd_stager::kStageBytes / wbytes);
  if (rows_per < 1) return mcd_fail(h, MCD_ERR_UNSUPPORTED, "a single row exceeds the staging buffer");
  int b = h->stage_next;
  for (int64_t r0 = 0; r0 < rows; r0 += rows_per, b ^= 1) {
    const int64_t nr = rows - r0 < rows_per ? rows - r0 : rows_per;
    MCD_CUDA(h, cudaEventSynchronize(h->stage_ev[b]));  // the previous transfer out of this buffer has finished
    h->stager->copy(static_cast<char*>(h->h_stage[b]), reinterpret_cast<const char*>(src + r0 * src_ld), wbytes,
                    (size_t)src_ld * 8, nr);
    MCD_CUDA(h, cudaMemcpy2DAsync(dst + r0 * dst_ld, (size_t)dst_ld * 8, h->h_stage[b], wbytes, wbytes, (size_t)nr,
                                  cudaMemcpyHostToDevice, stream));
    MCD_CUDA(h, cudaEventRecord(h->stage_ev[b], stream));
  }
  h->stage_next = b;
  return MCD_OK;
}
