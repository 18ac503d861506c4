// phx_hostpool.h -- a small persistent host thread pool for the *_host entry points (the part of
// the end-to-end path that runs on the CPU: expanding compact device results into the caller's
// float32 planes while later chunks are still crossing PCIe).  parallel_for blocks the caller
// (which takes a share of the work itself) until every slice is done.
#pragma once
#include <condition_variable>
#include <cstddef>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace phx {

class HostPool {
 public:
  // n_threads total workers including the calling thread (>= 1)
  explicit HostPool(int n_threads);
  ~HostPool();
  HostPool(const HostPool&) = delete;
  HostPool& operator=(const HostPool&) = delete;

  int size() const { return n_; }
  // fn(begin, end) over [0, count) split into size() contiguous slices whose boundaries are
  // multiples of `align` (the last slice takes the remainder)
  void parallel_for(size_t count, size_t align, const std::function<void(size_t, size_t)>& fn);

  // PHX_HOST_THREADS, else min(16, CPUs this process may run on)
  static int default_threads();

 private:
  void worker(int index);
  void run_slice(int index);

  int n_;
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_start_, cv_done_;
  unsigned long generation_ = 0;
  int pending_ = 0;
  bool stop_ = false;
  size_t count_ = 0, align_ = 1;
  const std::function<void(size_t, size_t)>* fn_ = nullptr;
};

}  // namespace phx
