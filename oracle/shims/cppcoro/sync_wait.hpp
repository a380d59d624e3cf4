// Shim (see task.hpp).
#pragma once
#include "task.hpp"

namespace cppcoro {

template <typename T>
T sync_wait(task<T>&& t) { return t.run_to_completion(); }
template <typename T>
T sync_wait(task<T>& t) { return t.run_to_completion(); }

}  // namespace cppcoro
