// Shim (see task.hpp): awaits the tasks one after the other and returns their results in order.
#pragma once
#include <vector>
#include "task.hpp"

namespace cppcoro {

template <typename T>
task<std::vector<T>> when_all(std::vector<task<T>> tasks) {
  std::vector<T> out;
  out.reserve(tasks.size());
  for (auto& t : tasks) out.push_back(co_await t);
  co_return out;
}

}  // namespace cppcoro
