// stand-in for magic_enum: enum_integer, enum_count (enumerators with consecutive values from 0, found by GCC's __PRETTY_FUNCTION__
// like magic_enum itself does), enum_name, bitwise operators.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
#include <string_view>
#include <type_traits>
#include <utility>
namespace magic_enum {
template <typename E> constexpr auto enum_integer(E e) noexcept { return static_cast<std::underlying_type_t<E>>(e); }
namespace detail {
template <typename E, E V> constexpr bool is_named() {
  const std::string_view s = __PRETTY_FUNCTION__;          // "... [with E = X; E V = X::Name]" or "... V = (X)7]"
  const std::size_t p = s.rfind("V = ");
  return p != std::string_view::npos && s[p + 4] != '(';
}
template <typename E, E V> constexpr std::string_view name_of() {
  const std::string_view s = __PRETTY_FUNCTION__;
  const std::size_t p = s.rfind("::"), q = s.rfind(';') == std::string_view::npos ? s.rfind(']') : s.rfind(']');
  return s.substr(p + 2, q - p - 2);
}
template <typename E, std::size_t... I> constexpr std::size_t count(std::index_sequence<I...>) { return (std::size_t(is_named<E, static_cast<E>(I)>()) + ... + 0); }
}  // namespace detail
template <typename E> constexpr std::size_t enum_count() noexcept { return detail::count<E>(std::make_index_sequence<64>{}); }
template <typename E> constexpr std::string_view enum_name(E) noexcept { return "enum"; }
namespace bitwise_operators {
template <typename E, typename = std::enable_if_t<std::is_enum_v<E>>> constexpr E operator|(E a, E b) noexcept { return static_cast<E>(enum_integer(a) | enum_integer(b)); }
template <typename E, typename = std::enable_if_t<std::is_enum_v<E>>> constexpr E operator&(E a, E b) noexcept { return static_cast<E>(enum_integer(a) & enum_integer(b)); }
template <typename E, typename = std::enable_if_t<std::is_enum_v<E>>> constexpr E operator^(E a, E b) noexcept { return static_cast<E>(enum_integer(a) ^ enum_integer(b)); }
template <typename E, typename = std::enable_if_t<std::is_enum_v<E>>> constexpr E operator~(E a) noexcept { return static_cast<E>(~enum_integer(a)); }
}  // namespace bitwise_operators
}  // namespace magic_enum
