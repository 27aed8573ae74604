// TEST INFRASTRUCTURE ONLY (oracle build).  Stand-in for gcem v1.13.1, which the
// reference fetches from GitHub at configure time (cmake/gcem.cmake:4-8) and
// which is not vendored under /root/reference.  The reference only uses
// gcem::sqrt / gcem::log for constexpr constants (RandomGen.hh:51-54,
// NEST.hh:205-207, LArParameters.hh:23-25, RandomGen.cpp:67); GCC evaluates the
// builtins at compile time with correct rounding.
#pragma once
namespace gcem {
constexpr double sqrt(double x) { return __builtin_sqrt(x); }
constexpr double log(double x) { return __builtin_log(x); }
}  // namespace gcem
