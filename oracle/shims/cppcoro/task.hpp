// Test-infrastructure shim (oracle/): just enough of cppcoro::task<T> for the reference's planner sources
// (/root/reference/oobleck/csrc/planning/pipeline_template.{h,cpp}) to compile UNMODIFIED in an image that has no cppcoro.
// A lazily started coroutine task, single-threaded (see static_thread_pool.hpp), driven by a trampoline: awaiting a
// child parks the parent and queues the child, a finished child queues its parent, and sync_wait resumes whatever is
// queued until the root is done.  The stack therefore stays flat however many sub-problems the search awaits (with
// symmetric transfer it would only stay flat if the compiler turned every transfer into a tail call, which GCC does not
// do at the -O1 the reference builds with; real cppcoro gets a fresh stack at every thread-pool hop instead).
// Not part of the product; only oracle/Makefile uses it to build oracle/_ref/.
#pragma once
#include <coroutine>
#include <exception>
#include <optional>
#include <utility>
#include <vector>

namespace cppcoro {

namespace detail {
inline std::vector<std::coroutine_handle<>>& ready_queue() {
  static thread_local std::vector<std::coroutine_handle<>> q;
  return q;
}
}  // namespace detail

template <typename T>
class task {
 public:
  struct promise_type {
    std::optional<T> value;
    std::exception_ptr error;
    std::coroutine_handle<> continuation;

    task get_return_object() { return task{std::coroutine_handle<promise_type>::from_promise(*this)}; }
    std::suspend_always initial_suspend() noexcept { return {}; }
    struct final_awaiter {
      bool await_ready() noexcept { return false; }
      void await_suspend(std::coroutine_handle<promise_type> h) noexcept {
        if (auto c = h.promise().continuation) detail::ready_queue().push_back(c);   // back to the trampoline
      }
      void await_resume() noexcept {}
    };
    final_awaiter final_suspend() noexcept { return {}; }
    template <typename U>
    void return_value(U&& v) { value.emplace(std::forward<U>(v)); }
    void unhandled_exception() { error = std::current_exception(); }
  };

  task() noexcept = default;
  explicit task(std::coroutine_handle<promise_type> h) noexcept : h_(h) {}
  task(task&& o) noexcept : h_(std::exchange(o.h_, {})) {}
  task& operator=(task&& o) noexcept {
    if (this != &o) {
      if (h_) h_.destroy();
      h_ = std::exchange(o.h_, {});
    }
    return *this;
  }
  task(const task&) = delete;
  task& operator=(const task&) = delete;
  ~task() { if (h_) h_.destroy(); }

  struct awaiter {
    std::coroutine_handle<promise_type> h;
    bool await_ready() const noexcept { return !h || h.done(); }
    void await_suspend(std::coroutine_handle<> cont) {
      h.promise().continuation = cont;
      detail::ready_queue().push_back(h);   // the trampoline starts the child; it queues `cont` when it finishes
    }
    T await_resume() {
      if (h.promise().error) std::rethrow_exception(h.promise().error);
      return std::move(*h.promise().value);
    }
  };
  awaiter operator co_await() & noexcept { return awaiter{h_}; }
  awaiter operator co_await() && noexcept { return awaiter{h_}; }

  // used by sync_wait below: the trampoline -- run this task and everything it awaits on the calling thread
  T run_to_completion() {
    auto& q = detail::ready_queue();
    const std::size_t base = q.size();          // re-entrancy: leave an outer trampoline's entries alone
    q.push_back(h_);
    while (!h_.done() && q.size() > base) {
      auto next = q.back();
      q.pop_back();
      next.resume();
    }
    if (h_.promise().error) std::rethrow_exception(h_.promise().error);
    return std::move(*h_.promise().value);
  }

 private:
  std::coroutine_handle<promise_type> h_;
};

}  // namespace cppcoro
