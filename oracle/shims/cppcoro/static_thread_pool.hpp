// Shim (see task.hpp): `co_await pool.schedule()` continues inline -- the search runs on the calling thread.  The
// reference only uses the pool to parallelise independent sub-problems of a memoised recursion; the result does not
// depend on the execution order of cache misses (every key is computed from its own sub-keys only).
#pragma once
#include <coroutine>

namespace cppcoro {

class static_thread_pool {
 public:
  struct schedule_operation {
    bool await_ready() const noexcept { return true; }
    void await_suspend(std::coroutine_handle<>) const noexcept {}
    void await_resume() const noexcept {}
  };
  schedule_operation schedule() noexcept { return {}; }
};

}  // namespace cppcoro
