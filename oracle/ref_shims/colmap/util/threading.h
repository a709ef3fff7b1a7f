// colmap/util/threading.h — SHIM (test infrastructure written for this repo; not COLMAP).
// colmap::ThreadPool as solve.cc:617-635 uses it: a fixed number of workers draining a FIFO
// task queue, AddTask(f, args...) and Wait().
#ifndef LFR_SHIM_COLMAP_THREADING_H_
#define LFR_SHIM_COLMAP_THREADING_H_
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <queue>
#include <thread>
#include <vector>

namespace colmap {

class ThreadPool {
 public:
  explicit ThreadPool(int num_threads) : stop_(false), active_(0), t_created_(std::chrono::steady_clock::now()), wait_ms_(0.0) {
    if (num_threads <= 0) num_threads = (int)std::thread::hardware_concurrency();
    if (num_threads <= 0) num_threads = 1;
    for (int i = 0; i < num_threads; ++i) workers_.emplace_back([this]() { Run(); });
  }
  ~ThreadPool() {
    {
      std::unique_lock<std::mutex> lock(mutex_);
      stop_ = true;
    }
    task_cv_.notify_all();
    for (auto& w : workers_) w.join();
    // bench.py's reference arm: solve.cc prints its "Solver time" (pool construction -> Wait(),
    // solve.cc:615-638) in whole milliseconds; the same scope in microseconds goes to the file
    // named by LFR_POOL_TIMING_FILE
    if (const char* path = std::getenv("LFR_POOL_TIMING_FILE")) {
      if (std::FILE* f = std::fopen(path, "a")) {
        std::fprintf(f, "%.3f %zu\n", wait_ms_, workers_.size());
        std::fclose(f);
      }
    }
  }
  template <class F, class... Args>
  void AddTask(F&& f, Args&&... args) {
    {
      std::unique_lock<std::mutex> lock(mutex_);
      tasks_.push(std::bind(std::forward<F>(f), std::forward<Args>(args)...));
    }
    task_cv_.notify_one();
  }
  void Wait() {
    std::unique_lock<std::mutex> lock(mutex_);
    done_cv_.wait(lock, [this]() { return tasks_.empty() && active_ == 0; });
    wait_ms_ = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_created_).count();
  }

 private:
  void Run() {
    for (;;) {
      std::function<void()> task;
      {
        std::unique_lock<std::mutex> lock(mutex_);
        task_cv_.wait(lock, [this]() { return stop_ || !tasks_.empty(); });
        if (stop_ && tasks_.empty()) return;
        task = std::move(tasks_.front());
        tasks_.pop();
        ++active_;
      }
      task();
      {
        std::unique_lock<std::mutex> lock(mutex_);
        --active_;
      }
      done_cv_.notify_all();
    }
  }
  std::vector<std::thread> workers_;
  std::queue<std::function<void()>> tasks_;
  std::mutex mutex_;
  std::condition_variable task_cv_, done_cv_;
  bool stop_;
  int active_;
  std::chrono::steady_clock::time_point t_created_;
  double wait_ms_;
};

}  // namespace colmap
#endif
