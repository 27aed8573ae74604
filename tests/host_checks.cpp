// Host-side checks of nest_b200/csrc/nb_rng.cuh (compiled and run by tests/test_host_rng_cpu.py; g++, no GPU):
// the Philox4x32-10 block function against the Random123 known-answer vectors, the round-key and
// per-event-prepared variants against it, the sequential Stream's word order, and the two arithmetic shortcuts of
// nb_yields.cuh (nearly_equal without its division, div_z) against the plain expressions they replace.
#include <cstdio>
#include <cstdint>
#include <cfloat>
#include <cmath>
#include <cstring>
#include "../nest_b200/csrc/nb_rng.cuh"
#include "../nest_b200/csrc/nb_yields.cuh"

static int fails = 0;
#define CHECK(c)                                          \
  do {                                                    \
    if (!(c)) {                                           \
      std::printf("FAIL line %d: %s\n", __LINE__, #c);    \
      ++fails;                                            \
    }                                                     \
  } while (0)

int main() {
  using namespace nb;
  {  // Random123 kat_vectors: philox4x32 10 rounds
    u128 r = philox4x32_10(0u, 0u, 0u, 0u, 0u, 0u);
    CHECK(r.x == 0x6627e8d5u && r.y == 0xe169c58du && r.z == 0xbc57ac4cu && r.w == 0x9b00dbd8u);
    r = philox4x32_10(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    CHECK(r.x == 0x408f276du && r.y == 0x41c83b0eu && r.z == 0xa20bc7c6u && r.w == 0x6d5451fdu);
    r = philox4x32_10(0x243f6a88u, 0x85a308d3u, 0x13198a2eu, 0x03707344u, 0xa4093822u, 0x299f31d0u);
    CHECK(r.x == 0xd16cfe09u && r.y == 0x94fdccebu && r.z == 0x5001e420u && r.w == 0x24126ea1u);
  }
  {  // precomputed round keys and the per-event split of round 1 give the same blocks
    const uint64_t seed = 0x123456789abcdef0ull;
    uint32_t rk[20];
    philox_round_keys(seed, rk);
    uint64_t s = 88172645463325252ull;
    for (int i = 0; i < 200000; ++i) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      const uint32_t c0 = uint32_t(s), c1 = uint32_t(s >> 32);
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      const uint32_t c2 = uint32_t(s), c3 = uint32_t(s >> 40) & 0xffu;
      const u128 a = philox4x32_10(c0, c1, c2, c3, uint32_t(seed), uint32_t(seed >> 32));
      const u128 b = philox4x32_10_rk(c0, c1, c2, c3, rk);
      const PhiloxEvent pe = philox_event_prep(c0, c1, c3, rk);
      const u128 c = philox4x32_10_rk_event(pe, c2, rk);
      CHECK(a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w);
      CHECK(a.x == c.x && a.y == c.y && a.z == c.z && a.w == c.w);
      if (fails > 5) break;
    }
  }
  {  // Stream: blocks 0, 1, 2, ... of (seed; event, block, stream), words (x,y) then (z,w)
    uint32_t keys[20];
    philox_round_keys(42ull, keys);
    Stream g;
    g.init(keys, 0x100000007ull, STREAM_POST);
    for (uint32_t blk = 0; blk < 50; ++blk) {
      const u128 r = philox4x32_10(7u, 1u, blk, STREAM_POST, 42u, 0u);
      uint32_t a, b;
      g.next2(a, b);
      CHECK(a == r.x && b == r.y);
      g.next2(a, b);
      CHECK(a == r.z && b == r.w);
    }
  }
  {  // uniforms live on the open interval
    CHECK(u53(0u, 0u) > 0. && u53(0xffffffffu, 0xffffffffu) < 1.);  // (2^53 - 1 + 1/2) 2^-53 rounds to 1 unclamped
    CHECK(u53(0xffffffffu, 0xfffff000u) < 1. && u53(0xffffffffu, 0xfffff000u) > 0.9999999999999997);
    CHECK(u32d(0u) > 0. && u32d(0xffffffffu) < 1.);
    CHECK(u53(0x80000000u, 0u) == 0.5 + 1.1102230246251565404e-16 * 0.5);
  }
  {  // nearly_equal decides diff / sum < eps without the division unless the quotient is within a factor 2 of eps:
     // the same answer as ValidityTests::nearlyEqual (ValidityTests.cpp:6-21) written out, on 4e6 pairs that crowd
     // the threshold, plus the special cases
    auto plain = [](double a, double b, double eps) {
      if (a == b) return true;
      const double absA = std::fabs(a), absB = std::fabs(b), diff = std::fabs(a - b);
      if (a == 0 || b == 0 || (absA + absB < DBL_MIN)) return diff < (eps * DBL_MIN);
      return diff / std::fmin(absA + absB, DBL_MAX) < eps;
    };
    uint32_t keys[20];
    philox_round_keys(7ull, keys);
    Stream g;
    g.init(keys, 1ull, STREAM_HOST);
    int bad = 0;
    for (int i = 0; i < 4000000 && bad < 5; ++i) {
      const double a = std::ldexp(g.uniform() - 0.5, int(g.uniform() * 600.) - 300);
      // relative distance log-uniform around 2e-9 (the quotient's threshold for eps = 1e-9), every few draws exact
      const double rel = 2e-9 * std::exp((g.uniform() - 0.5) * ((i & 3) ? 2.0 : 40.0));
      const double b = (i % 7 == 0) ? a : a * (1. + ((i & 1) ? rel : -rel));
      if (nearly_equal(a, b) != plain(a, b, 1e-9)) ++bad;
      if (nearly_equal(a, 1.) != plain(a, 1., 1e-9)) ++bad;
    }
    CHECK(bad == 0);
    const double sp[] = {0., -0., 1., -1., DBL_MIN, 4.9e-324, 1e-320, DBL_MAX, -DBL_MAX, 1. + 2e-9, 1. + 1.9999e-9, 1. + 2.0001e-9};
    for (double a : sp)
      for (double b : sp) CHECK(nearly_equal(a, b) == plain(a, b, 1e-9));
  }
  {  // div_z(a, b) has the bits of a / b, whatever the operands
    const double v[] = {0., -0., 1., -1., 3.5, 1e-310, DBL_MIN, DBL_MAX, 1e300, INFINITY, -INFINITY, NAN, 0.37, -2e-9};
    for (double a : v)
      for (double b : v) {
        const double q = a / b, z = div_z(a, b);
        CHECK(std::memcmp(&q, &z, 8) == 0 || (std::isnan(q) && std::isnan(z)));
      }
  }
  std::printf(fails ? "host checks: %d failure(s)\n" : "host checks ok\n", fails);
  return fails != 0;
}
