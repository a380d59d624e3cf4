// Test-infrastructure shim (oracle/): the reference's planner memoises in a oneTBB concurrent map; single-threaded here
// (cppcoro shim), so std::unordered_map has the interface it uses (find / end / insert).
#pragma once
#include <functional>
#include <unordered_map>

namespace oneapi::tbb {
template <typename K, typename V, typename H = std::hash<K>, typename E = std::equal_to<K>>
using concurrent_unordered_map = std::unordered_map<K, V, H, E>;
}  // namespace oneapi::tbb
