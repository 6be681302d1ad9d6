// host_stage.h -- uploads from PAGEABLE host memory (what an unmodified front end holds: std::vector<Eigen::Vector3d>,
// ref: include/tloam/open3d/PointCloud2.hpp:396) at close to PCIe speed.
//
// cudaMemcpyAsync from pageable memory goes through the driver's own staging at ~10 GB/s (13 MB of map per frame: 1.2 ms of
// the 1.6 ms frame).  Here the copy is cut into 2 MB chunks that a small pool of threads copies into a ring of pinned slots
// (several threads: one core's memcpy is ~12 GB/s, PCIe 5 x16 moves 50+) while the DMA engine drains the slots filled
// before (the workers poll for the next chunk while a burst lasts and sleep otherwise); a slot is reused once the event recorded behind its DMA has fired.  The caller's buffer has been read completely
// when the call returns.  Pinned (or registered) caller memory never comes here: it is DMA'd directly.
#pragma once
#include <cuda_runtime.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace tloam {

class HostStage {
 public:
  static constexpr size_t kChunk = 2u << 20;     // bytes per pinned slot
  static constexpr int kSlots = 8;
  static constexpr int kWorkers = 7;             // + the calling thread: one core's memcpy is ~10 GB/s, PCIe 5 x16 moves 50+
  static constexpr long long kSpinNs = 400000;   // a worker keeps polling this long after its last task before it sleeps

  ~HostStage() { shutdown(); }

  // true if `p` is ordinary pageable memory (not pinned / registered / managed / device)
  static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();                                   // be kind to the sibling hardware thread while polling
#endif
  }

  static bool pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
  }

  // dst (device) <- src (pageable host), enqueued on `stream`; returns when src has been read completely
  cudaError_t upload(void* dst, const void* src, size_t bytes, cudaStream_t stream) {
    cudaError_t e = init();
    if (e != cudaSuccess) return e;
    const char* s = static_cast<const char*>(src);
    char* d = static_cast<char*>(dst);
    for (size_t off = 0; off < bytes; off += kChunk) {
      const size_t len = bytes - off < kChunk ? bytes - off : kChunk;
      const int slot = next_;
      next_ = (next_ + 1) % kSlots;
      if (used_[slot]) { e = cudaEventSynchronize(ev_[slot]); if (e != cudaSuccess) return e; }
      parallel_copy(pin_ + (size_t)slot * kChunk, s + off, len);
      e = cudaMemcpyAsync(d + off, pin_ + (size_t)slot * kChunk, len, cudaMemcpyHostToDevice, stream);
      if (e != cudaSuccess) return e;
      e = cudaEventRecord(ev_[slot], stream);
      if (e != cudaSuccess) return e;
      used_[slot] = true;
    }
    return cudaSuccess;
  }

  void shutdown() {
    stop_.store(true, std::memory_order_release);
    {
      std::unique_lock<std::mutex> lk(m_);
      cv_work_.notify_all();
    }
    for (std::thread& t : workers_) if (t.joinable()) t.join();
    workers_.clear();
    if (pin_) {
      for (int i = 0; i < kSlots; ++i) { if (used_[i]) cudaEventSynchronize(ev_[i]); if (ev_[i]) cudaEventDestroy(ev_[i]); ev_[i] = nullptr; used_[i] = false; }
      cudaFreeHost(pin_);
      pin_ = nullptr;
    }
    stop_.store(false, std::memory_order_release);
  }

 private:
  struct Task { char* dst; const char* src; size_t len; };

  cudaError_t init() {
    if (pin_) return cudaSuccess;
    cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&pin_), kChunk * kSlots, cudaHostAllocDefault);
    if (e != cudaSuccess) { pin_ = nullptr; return e; }
    for (int i = 0; i < kSlots; ++i) {
      e = cudaEventCreateWithFlags(&ev_[i], cudaEventDisableTiming);
      if (e != cudaSuccess) return e;
      used_[i] = false;
    }
    try {
      for (int w = 0; w < kWorkers; ++w) workers_.emplace_back([this, w] { run(w); });
    } catch (...) {                                                         // no exception crosses the C ABI: fewer workers is an error
      shutdown();
      return cudaErrorUnknown;
    }
    return cudaSuccess;
  }

  // workers poll the generation counter for kSpinNs after their last task (a frame's uploads arrive as a burst of chunks:
  // a condition-variable wake-up per chunk costs more than copying it), then sleep
  void run(int w) {
    unsigned seen = 0;
    auto last = std::chrono::steady_clock::now();
    while (true) {
      unsigned g = generation_.load(std::memory_order_acquire);
      if (g == seen) {
        cpu_relax();
        if (stop_.load(std::memory_order_acquire)) return;
        if (std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - last).count() > kSpinNs) {
          std::unique_lock<std::mutex> lk(m_);
          sleepers_.fetch_add(1, std::memory_order_acq_rel);
          cv_work_.wait(lk, [&] { return stop_.load(std::memory_order_acquire) || generation_.load(std::memory_order_acquire) != seen; });
          sleepers_.fetch_sub(1, std::memory_order_acq_rel);
          if (stop_.load(std::memory_order_acquire)) return;
          last = std::chrono::steady_clock::now();
        }
        continue;
      }
      seen = g;
      const Task t = tasks_[w];
      if (t.len) memcpy(t.dst, t.src, t.len);
      done_.fetch_add(1, std::memory_order_acq_rel);
      last = std::chrono::steady_clock::now();
    }
  }

  void parallel_copy(char* dst, const char* src, size_t len) {
    if (len < (128u << 10)) { memcpy(dst, src, len); return; }            // small: not worth a hand-off
    const size_t part = ((len / (kWorkers + 1)) + 4095) & ~(size_t)4095;
    for (int w = 0; w < kWorkers; ++w) {
      const size_t o = part * (size_t)(w + 1);
      const size_t l = o < len ? (len - o < part ? len - o : part) : 0;
      tasks_[w] = Task{dst + o, src + o, l};
    }
    done_.store(0, std::memory_order_release);
    generation_.fetch_add(1, std::memory_order_acq_rel);                   // publishes the tasks
    if (sleepers_.load(std::memory_order_acquire) > 0) {
      std::unique_lock<std::mutex> lk(m_);
      cv_work_.notify_all();
    }
    memcpy(dst, src, part < len ? part : len);                            // the caller's share
    while (done_.load(std::memory_order_acquire) < kWorkers) cpu_relax();   // the others finish within microseconds
  }

  char* pin_ = nullptr;
  cudaEvent_t ev_[kSlots] = {};
  bool used_[kSlots] = {};
  int next_ = 0;
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_work_;
  Task tasks_[kWorkers] = {};
  std::atomic<unsigned> generation_{0};
  std::atomic<int> done_{0}, sleepers_{0};
  std::atomic<bool> stop_{false};
};

}  // namespace tloam
