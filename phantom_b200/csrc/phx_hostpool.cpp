// phx_hostpool.cpp -- see phx_hostpool.h
#include "phx_hostpool.h"

#include <sched.h>

#include <algorithm>
#include <cstdlib>

namespace phx {

int HostPool::default_threads() {
  if (const char* s = std::getenv("PHX_HOST_THREADS")) {
    const int n = std::atoi(s);
    if (n >= 1) return std::min(n, 256);
  }
  cpu_set_t set;
  int cpus = 1;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) cpus = std::max(1, CPU_COUNT(&set));
  return std::min(16, cpus);
}

HostPool::HostPool(int n_threads) : n_(std::max(1, n_threads)) {
  for (int i = 1; i < n_; ++i) threads_.emplace_back([this, i] { worker(i); });
}

HostPool::~HostPool() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
    ++generation_;
  }
  cv_start_.notify_all();
  for (auto& t : threads_) t.join();
}

void HostPool::run_slice(int index) {
  const size_t units = (count_ + align_ - 1) / align_;
  const size_t per = (units + n_ - 1) / n_;
  const size_t b = std::min(count_, (size_t)index * per * align_);
  const size_t e = std::min(count_, ((size_t)index + 1) * per * align_);
  if (b < e) (*fn_)(b, e);
}

void HostPool::worker(int index) {
  unsigned long seen = 0;
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_start_.wait(lk, [&] { return generation_ != seen; });
      seen = generation_;
      if (stop_) return;
    }
    run_slice(index);
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (--pending_ == 0) cv_done_.notify_one();
    }
  }
}

void HostPool::parallel_for(size_t count, size_t align,
                            const std::function<void(size_t, size_t)>& fn) {
  if (count == 0) return;
  {
    std::lock_guard<std::mutex> lk(mu_);
    count_ = count;
    align_ = std::max<size_t>(1, align);
    fn_ = &fn;
    pending_ = n_ - 1;
    ++generation_;
  }
  cv_start_.notify_all();
  run_slice(0);
  std::unique_lock<std::mutex> lk(mu_);
  cv_done_.wait(lk, [&] { return pending_ == 0; });
  fn_ = nullptr;
}

}  // namespace phx
