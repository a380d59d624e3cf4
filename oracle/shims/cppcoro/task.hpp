// Test-infrastructure shim (oracle/): just enough of cppcoro::task<T> for the reference's planner sources
// (/root/reference/oobleck/csrc/planning/pipeline_template.{h,cpp}) to compile UNMODIFIED in an image that has no cppcoro.
// A lazily started coroutine task with symmetric transfer, single-threaded (see static_thread_pool.hpp).
// Not part of the product; only oracle/Makefile uses it to build oracle/_ref/.
#pragma once
#include <coroutine>
#include <exception>
#include <optional>
#include <utility>

namespace cppcoro {

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
      std::coroutine_handle<> await_suspend(std::coroutine_handle<promise_type> h) noexcept {
        auto c = h.promise().continuation;
        return c ? c : std::noop_coroutine();
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
    std::coroutine_handle<> await_suspend(std::coroutine_handle<> cont) noexcept {
      h.promise().continuation = cont;
      return h;   // symmetric transfer: start the child, it resumes `cont` from its final suspend
    }
    T await_resume() {
      if (h.promise().error) std::rethrow_exception(h.promise().error);
      return std::move(*h.promise().value);
    }
  };
  awaiter operator co_await() & noexcept { return awaiter{h_}; }
  awaiter operator co_await() && noexcept { return awaiter{h_}; }

  // used by sync_wait below: run to completion on the calling thread
  T run_to_completion() {
    if (!h_.done()) h_.resume();
    if (h_.promise().error) std::rethrow_exception(h_.promise().error);
    return std::move(*h_.promise().value);
  }

 private:
  std::coroutine_handle<promise_type> h_;
};

}  // namespace cppcoro
