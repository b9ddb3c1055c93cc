// stand-in for oneTBB: the declarations the reference's headers name; the golden-vector driver never runs a parallel loop.
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include <algorithm>
#include <functional>
#include <ranges>
#include <type_traits>
namespace tbb {
template <typename T> struct blocked_range {
  T b_, e_;
  blocked_range(T b, T e) : b_(b), e_(e) {}
  T begin() const { return b_; }
  T end() const { return e_; }
};
template <typename Range, typename Body> void parallel_for(const Range& r, const Body& body) { body(r); }
template <typename T> struct combinable {
  T v_{};
  combinable() = default;
  // oneTBB: combinable(finit) with a callable, or an exemplar value to copy from
  template <typename F> explicit combinable(F f) { if constexpr (std::is_invocable_v<F>) v_ = f(); else v_ = f; }
  T& local() { return v_; }
  template <typename F> T combine(F) { return v_; }
  template <typename F> void combine_each(F f) { f(v_); }
};
struct global_control { enum parameter { max_allowed_parallelism }; global_control(parameter, std::size_t) {} };
namespace this_task_arena { inline int max_concurrency() { return 1; } }
}  // namespace tbb
namespace oneapi { namespace tbb = ::tbb; }
