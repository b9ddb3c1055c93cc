// stand-in for SYCL: the reference only constructs a device object to print its name. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <string>
namespace sycl {
struct cpu_selector_t {}; struct gpu_selector_t {};
inline constexpr cpu_selector_t cpu_selector_v{}; inline constexpr gpu_selector_t gpu_selector_v{};
namespace info { namespace device { struct name {}; } }
struct device {
  device() = default;
  explicit device(cpu_selector_t) {}
  explicit device(gpu_selector_t) {}
  template <typename T> std::string get_info() const { return "none"; }
};
}  // namespace sycl
