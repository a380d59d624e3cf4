// Test-infrastructure shim (oracle/): every heap allocation of the reference planner module starts zeroed.
// StageExecutionResult::forward_ / backward_ / mem_required_ are accumulated with += without ever being initialised
// (/root/reference/oobleck/csrc/planning/execution_result.h:78-112); on a fresh heap they happen to be zero, on a
// recycled block they are whatever was there.  The module is linked with -Bsymbolic, so these replacements serve the
// module's own allocations only.
#include <cstdlib>
#include <new>

void* operator new(std::size_t n) {
  if (void* p = std::calloc(1, n ? n : 1)) return p;
  throw std::bad_alloc();
}
void* operator new[](std::size_t n) {
  if (void* p = std::calloc(1, n ? n : 1)) return p;
  throw std::bad_alloc();
}
void operator delete(void* p) noexcept { std::free(p); }
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete(void* p, std::size_t) noexcept { std::free(p); }
void operator delete[](void* p, std::size_t) noexcept { std::free(p); }
