// nb_rng.cuh -- counter-based random streams and the distribution samplers of the chain.
//
// Replaces the reference's process-global sequential generator (RandomGen singleton over
// xoroshiro128+, src/RandomGen.cpp:8-42) with Philox4x32-10 keyed by the run seed and
// counted by (event id, draw block, stream id): every event owns its streams, so the
// result for event e depends only on (seed, e, parameters).  Only the *distributions* of
// the reference samplers are kept (src/RandomGen.cpp:39-182); where a cheaper exact
// construction of the same distribution exists it is used and said so at the sampler.
#pragma once
#include <cstdint>
#include <cfloat>
#include <cmath>

#ifdef __CUDACC__
#define NB_HD __host__ __device__ __forceinline__
#define NB_D __device__ __forceinline__
#else
#define NB_HD inline
#define NB_D inline
#endif

// layers of the hit loop's ziggurat normal: 2^NB_ZIG_BITS
#ifndef NB_ZIG_BITS
#define NB_ZIG_BITS 11
#endif
// alias tables of the fixed-probability binomial sampler (binomial_fixed_p below)
#define NB_ALIAS_BITS 12
#define NB_ALIAS_ENTRIES ((2 << NB_ALIAS_BITS) - 1 + NB_ALIAS_BITS + 1)  // 8204
#define NB_ALIAS_MAXN (16 << NB_ALIAS_BITS)  // 65536; beyond: BTRS

namespace nb {

// stream ids (counter word 3)
enum : uint32_t {
  STREAM_MAIN = 0,   // per-event scalar draws, sequential blocks
  STREAM_HIT = 1,    // S1 PMT hits: block = hit index (SPE part)
  STREAM_HIT2 = 2,   // S1 PMT hits: second photo-electron of a double-phe hit
  STREAM_HITX = 3,   // S1 PMT hits: rare redraws of the zero-truncated Gaussian
  STREAM_HIT3 = 4,   // S1 PMT hits: single-phe efficiency decisions (sPEeff < 1 only)
  STREAM_POST = 5,   // per-event scalar draws after the S1 hit loop (k_post)
  STREAM_HIT4 = 6,   // S1 PMT hits: truncation check of a second photo-electron
  STREAM_ZIG = 7,    // S1 PMT hits: ziggurat wedge / tail completion of a hit's normal (| attempt << 8)
  STREAM_ZIG2 = 10,  // ... of its second photo-electron's normal
  STREAM_ZIGX = 11,  // ... of a joint (S, T) redraw (11: S, 12: T)
  STREAM_LAR = 8,
  STREAM_SE = 9,     // CalculateG2's single-electron Monte Carlo
  // each binomial draw owns a stream: candidate k of a draw is Philox block k of it, whatever else the event consumes
  STREAM_BIN_NI = 32, STREAM_BIN_S1 = 33, STREAM_BIN_NEE = 34, STREAM_BIN_S2HIT = 35,
  STREAM_BIN_S2DPE = 36, STREAM_BIN_PAR0 = 37, STREAM_BIN_PAR1 = 38, STREAM_BIN_PAR2 = 39,
  STREAM_BIN_LAR = 40, STREAM_BIN_S1DPE = 41, /* 42: its BTRS fall-back */
  STREAM_BIN_SE_HIT = 48, STREAM_BIN_SE_DPE = 49,  // CalculateG2's single-electron Monte Carlo (not the chain's events)
  STREAM_HOST = 16   // host-side draws (Poisson event count, WIMP table)
};

struct u128 {
  uint32_t x, y, z, w;
};

NB_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
  uint64_t p = uint64_t(a) * uint64_t(b);  // one IMAD.WIDE.U32
  lo = uint32_t(p);
  hi = uint32_t(p >> 32);
}

// Philox4x32-10 (Salmon et al., SC'11); constants of the Random123 reference implementation.
NB_HD u128 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                         uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    mulhilo(0xD2511F53u, c0, hi0, lo0);
    mulhilo(0xCD9E8D57u, c2, hi1, lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return u128{c0, c1, c2, c3};
}

#ifdef __CUDACC__
// Out-of-line copy for the rare completions of the hit loop (zig_finish, pe_redraw): those are subroutines of
// k_hits and its finisher, and what they need in registers is taken from the callers' budget -- with the block
// function in line they cost the hit loop its spill-free main loop (152 bytes of spills, k_hits + 12 %).
__device__ __noinline__ u128 philox_block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                          uint32_t k1) {
  return philox4x32_10(c0, c1, c2, c3, k0, k1);
}
#endif

// Same function with the 10 round keys taken from a precomputed table (rk[2r], rk[2r+1]).  The kernels
// pass a pointer into their own parameter block (RunParams::rk, SweepArgs::rk, ...): once everything is
// inlined the table is a constant-bank operand, so each round is two IMAD.WIDE and two three-input LOP3
// with a constant operand -- no key-schedule adds (2 of the 6 instructions of a round; the scalar stages
// k_pre / k_post issue-bind in their Philox phases, all warps of a block being in step).
NB_HD void philox_round_keys(uint64_t seed, uint32_t* rk) {
  uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    rk[2 * r] = k0;
    rk[2 * r + 1] = k1;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
NB_HD u128 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t* rk) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    mulhilo(0xD2511F53u, c0, hi0, lo0);
    mulhilo(0xCD9E8D57u, c2, hi1, lo1);
    uint32_t n0 = hi1 ^ c1 ^ rk[2 * r];
    uint32_t n2 = hi0 ^ c3 ^ rk[2 * r + 1];
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
  }
  return u128{c0, c1, c2, c3};
}

// The same block with the first round split off for a loop that walks blocks (c0, c1, *, c3) of one event:
// of round 1 only the product with c2 changes from block to block; the c0 product, n2 and the key/counter
// mix of n0 are per-event constants (philox_event_prep).  Bit-identical to philox4x32_10_rk.
struct PhiloxEvent {
  uint32_t x;    // c1 ^ rk[0]
  uint32_t n2;   // hi(M0 c0) ^ c3 ^ rk[1]
  uint32_t lo0;  // lo(M0 c0)
};
NB_HD PhiloxEvent philox_event_prep(uint32_t c0, uint32_t c1, uint32_t c3, const uint32_t* rk) {
  uint32_t hi0, lo0;
  mulhilo(0xD2511F53u, c0, hi0, lo0);
  return PhiloxEvent{c1 ^ rk[0], hi0 ^ c3 ^ rk[1], lo0};
}
NB_HD u128 philox4x32_10_rk_event(const PhiloxEvent& ev, uint32_t blk, const uint32_t* rk) {
  uint32_t hi1, lo1;
  mulhilo(0xCD9E8D57u, blk, hi1, lo1);
  uint32_t c0 = hi1 ^ ev.x, c1 = lo1, c2 = ev.n2, c3 = ev.lo0;
#pragma unroll
  for (int r = 1; r < 10; ++r) {
    uint32_t hi0, lo0;
    mulhilo(0xD2511F53u, c0, hi0, lo0);
    mulhilo(0xCD9E8D57u, c2, hi1, lo1);
    uint32_t n0 = hi1 ^ c1 ^ rk[2 * r];
    uint32_t n2 = hi0 ^ c3 ^ rk[2 * r + 1];
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
  }
  return u128{c0, c1, c2, c3};
}

// 53-bit uniform on the OPEN interval (0,1).  The reference's rand_uniform is
// u64/(2^64-1) on the closed interval (RandomGen.cpp:39-42); the end points have
// probability 2^-64 each and only matter as log(0) hazards, which the open interval removes.
NB_HD double u53(uint32_t hi, uint32_t lo) {
  uint64_t v = (uint64_t(hi) << 21) | (uint64_t(lo) >> 11);
  const double u = (double(v) + 0.5) * 1.1102230246251565404e-16;  // 2^-53
  // v = 2^53 - 1 alone rounds up to 1.0 (2^53 - 1/2 is not a double): keep the interval open
  return u < 1. ? u : 0.99999999999999988898;  // 1 - 2^-53
}
// 32-bit uniform on (0,1) for yes/no decisions and angles
NB_HD double u32d(uint32_t a) { return (double(a) + 0.5) * 2.3283064365386962891e-10; }

// Sequential stream over one (seed, event, stream) triple; the seed comes as its 20 round keys
// (philox_round_keys), which the stream only points to.
struct Stream {
  const uint32_t* rk;
  uint32_t e0, e1, blk, sid;
  uint32_t s0, s1;  // second half of the last block
  bool have;

  NB_HD void init(const uint32_t* keys, uint64_t event, uint32_t stream) {
    rk = keys;
    e0 = uint32_t(event);
    e1 = uint32_t(event >> 32);
    blk = 0;
    sid = stream;
    have = false;
    s0 = s1 = 0;
  }
  NB_HD void next2(uint32_t& a, uint32_t& b) {
    if (have) {
      have = false;
      a = s0;
      b = s1;
      return;
    }
    u128 r = philox4x32_10_rk(e0, e1, blk, sid, rk);
    ++blk;
    s0 = r.z;
    s1 = r.w;
    have = true;
    a = r.x;
    b = r.y;
  }
  NB_HD double uniform() {
    uint32_t a, b;
    next2(a, b);
    return u53(a, b);
  }
};

#ifdef __CUDACC__
// ---- specialised FP64 math for the S1 hit loop -------------------------------------------
// The hit loop spends its time in -ln(u), sqrt and sincos of RANDOM BITS.  libm's log and
// sincospi pay for generic arguments (frexp, special values, large-argument reduction); here
// the argument is built from the bits, so the reduction is free and the remaining polynomials
// are short.  Both are accurate to ~1e-16 (tests/test_gpu_hitmath.py compares with libm).

// {1/c_j, ln c_j} for c_j = 1 + (j + 1/2)/32, j = 0..31
static __constant__ double c_lntab[64] = {
    0.98461538461538461538, 0.015504186535965254151, 0.95522388059701492537, 0.045809536031294203167,
    0.92753623188405797101, 0.075223421237587525699, 0.90140845070422535211, 0.10379679368164356483,
    0.87671232876712328767, 0.13157635778871927259,  0.85333333333333333333, 0.15860503017663858409,
    0.83116883116883116883, 0.18492233849401199266,  0.81012658227848101266, 0.21056476910734963767,
    0.79012345679012345679, 0.23556607131276690908,  0.77108433734939759036, 0.25995752443692606697,
    0.75294117647058823529, 0.28376817313064459835,  0.73563218390804597701, 0.30702503529491186208,
    0.71910112359550561798, 0.32975328637246798181,  0.7032967032967032967,  0.35197642315717818466,
    0.68817204301075268817, 0.37371640979358408082,  0.67368421052631578947, 0.39499380824086897811,
    0.65979381443298969072, 0.41582789514371096561,  0.64646464646464646465, 0.43623676677491807035,
    0.63366336633663366337, 0.45623743348158759438,  0.62135922330097087379, 0.47584590486996391427,
    0.60952380952380952381, 0.4950772667978515146,   0.5981308411214953271,  0.5139457511022343168,
    0.58715596330275229358, 0.53246479886947184387,  0.57657657657657657658, 0.55064711795266227926,
    0.56637168141592920354, 0.56850473535266871208,  0.55652173913043478261, 0.5860490450035782089,
    0.54700854700854700855, 0.60329085143808426234,  0.53781512605042016807, 0.62024040975185752885,
    0.5289256198347107438,  0.63690746223706923162,  0.52032520325203252033, 0.65330127201274563876,
    0.512,                  0.6694306539426292673,   0.50393700787401574803, 0.68530400309891941654};

// -ln(u) for u = x / 2^64 built from 64 random bits (x = 0 is mapped to x = 1).  With e leading
// zeros, u = m 2^-(e+1), m in [1,2) from the next 53 bits; ln m = ln c_j + ln(1 + d) with c_j the
// centre of m's 1/32-wide cell and d = m/c_j - 1, |d| <= 2^-6 (Taylor through d^9: 9e-20).
// tab: c_lntab staged in shared memory by the caller.
// polynomial coefficients live in constant memory so that each DFMA takes its coefficient as a
// constant-bank operand (64-bit literals would cost two UMOV each, every iteration)
static __constant__ double c_lnp[8] = {1.0 / 9.0,  -1.0 / 8.0, 1.0 / 7.0, -1.0 / 6.0,
                                       1.0 / 5.0,  -1.0 / 4.0, 1.0 / 3.0, -0.5};
static __constant__ double c_sinp[7] = {-1.0 / 1307674368000.0, 1.0 / 6227020800.0, -1.0 / 39916800.0,
                                        1.0 / 362880.0,         -1.0 / 5040.0,      1.0 / 120.0,
                                        -1.0 / 6.0};
static __constant__ double c_cosp[8] = {1.0 / 20922789888000.0, -1.0 / 87178291200.0, 1.0 / 479001600.0,
                                        -1.0 / 3628800.0,       1.0 / 40320.0,        -1.0 / 720.0,
                                        1.0 / 24.0,             -0.5};
static __constant__ double c_hitk[4] = {0.69314718055994530942, 1.4629180792671596e-9, 536870912.0,
                                        1073741824.0};

NB_D double neg_log_bits(uint32_t hi, uint32_t lo, const double* tab) {
  unsigned long long x = ((unsigned long long)hi << 32) | lo;
  x |= 1ull;
  const int e = __clzll((long long)x);
  const unsigned long long y = x << e;
  const double m = __longlong_as_double((long long)(0x3FF0000000000000ull | ((y << 1) >> 12)));
  const int j = int((y >> 58) & 31ull);
  const double2 t = *reinterpret_cast<const double2*>(tab + 2 * j);
  const double d = fma(m, t.x, -1.0);
  double q = c_lnp[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) q = fma(q, d, c_lnp[i]);
  const double ln1pd = fma(d * d, q, d);
  return fma(double(e + 1), c_hitk[0], -t.y) - ln1pd;
}

// (sin, cos) of the angle 2 pi (z + 1/2) / 2^32: quadrant from the top two bits, the remaining
// 30 bits folded onto [0, pi/4] (Taylor through x^15 / x^16: 5e-17).
NB_D void sincos_bits(uint32_t z, double& s, double& c) {
  double f = double(z & 0x3FFFFFFFu) + 0.5;  // in (0, 2^30)
  const bool fold = f > c_hitk[2];            // beyond pi/4: use the complementary angle
  if (fold) f = c_hitk[3] - f;
  const double x = f * c_hitk[1];  // (pi/2) / 2^30
  const double x2 = x * x;
  double ps = c_sinp[0];
#pragma unroll
  for (int i = 1; i < 7; ++i) ps = fma(ps, x2, c_sinp[i]);
  const double sn = fma(x * x2, ps, x);
  double pc = c_cosp[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) pc = fma(pc, x2, c_cosp[i]);
  const double cs = fma(x2, pc, 1.0);
  // quadrant q (top two bits): (cos, sin) of q pi/2 + a with a = x or pi/2 - x (fold).  The values
  // swap when exactly one of (q odd, fold) holds; the signs are cos < 0 for q in {1,2}, sin < 0
  // for q in {2,3}: applied to the sign bit of the high word.
  const bool swap = (((z >> 30) & 1u) != 0u) != fold;
  const double a = swap ? sn : cs, b = swap ? cs : sn;
  const uint32_t q = z >> 30;
  const uint32_t sc = ((q ^ (q >> 1)) & 1u) << 31;  // q = 1, 2
  const uint32_t ss = (q >> 1) << 31;               // q = 2, 3
  c = __hiloint2double(__double2hiint(a) ^ int(sc), __double2loint(a));
  s = __hiloint2double(__double2hiint(b) ^ int(ss), __double2loint(b));
}

// The same table in global memory for the per-event stages (k_pre / k_post and their samplers),
// which have no shared-memory staging: a random index into constant memory
// would serialise, a read-only global load is one instruction and stays in L1.
static __device__ const double g_lntab[64] = {
    0.98461538461538461538, 0.015504186535965254151, 0.95522388059701492537, 0.045809536031294203167,
    0.92753623188405797101, 0.075223421237587525699, 0.90140845070422535211, 0.10379679368164356483,
    0.87671232876712328767, 0.13157635778871927259,  0.85333333333333333333, 0.15860503017663858409,
    0.83116883116883116883, 0.18492233849401199266,  0.81012658227848101266, 0.21056476910734963767,
    0.79012345679012345679, 0.23556607131276690908,  0.77108433734939759036, 0.25995752443692606697,
    0.75294117647058823529, 0.28376817313064459835,  0.73563218390804597701, 0.30702503529491186208,
    0.71910112359550561798, 0.32975328637246798181,  0.7032967032967032967,  0.35197642315717818466,
    0.68817204301075268817, 0.37371640979358408082,  0.67368421052631578947, 0.39499380824086897811,
    0.65979381443298969072, 0.41582789514371096561,  0.64646464646464646465, 0.43623676677491807035,
    0.63366336633663366337, 0.45623743348158759438,  0.62135922330097087379, 0.47584590486996391427,
    0.60952380952380952381, 0.4950772667978515146,   0.5981308411214953271,  0.5139457511022343168,
    0.58715596330275229358, 0.53246479886947184387,  0.57657657657657657658, 0.55064711795266227926,
    0.56637168141592920354, 0.56850473535266871208,  0.55652173913043478261, 0.5860490450035782089,
    0.54700854700854700855, 0.60329085143808426234,  0.53781512605042016807, 0.62024040975185752885,
    0.5289256198347107438,  0.63690746223706923162,  0.52032520325203252033, 0.65330127201274563876,
    0.512,                  0.6694306539426292673,   0.50393700787401574803, 0.68530400309891941654};

// ln(1 + d) for |d| <= 2^-6, Taylor through d^9 (as in neg_log_bits)
NB_D double ln1p_small(double d) {
  double q = c_lnp[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) q = fma(q, d, c_lnp[i]);
  return fma(d * d, q, d);
}
// -ln(u), u = (hi:lo)/2^64, with the global table (per-event stages)
NB_D double neg_log_bits_g(uint32_t hi, uint32_t lo) {
  unsigned long long x = ((unsigned long long)hi << 32) | lo;
  x |= 1ull;
  const int e = __clzll((long long)x);
  const unsigned long long y = x << e;
  const double m = __longlong_as_double((long long)(0x3FF0000000000000ull | ((y << 1) >> 12)));
  const int j = int((y >> 58) & 31ull);
  const double2 t = __ldg(reinterpret_cast<const double2*>(g_lntab) + j);
  return fma(double(e + 1), c_hitk[0], -t.y) - ln1p_small(fma(m, t.x, -1.0));
}
// ln(x) for a positive, finite, normal x: exponent and mantissa taken from the bits, the same table
// and polynomial (absolute error ~1e-16; no special cases -- callers guarantee the domain)
NB_D double log_pos(double x) {
  const long long b = __double_as_longlong(x);
  const int e = int(b >> 52) - 1023;
  const double m = __longlong_as_double((b & 0x000FFFFFFFFFFFFFll) | 0x3FF0000000000000ll);
  const int j = int(b >> 47) & 31;
  const double2 t = __ldg(reinterpret_cast<const double2*>(g_lntab) + j);
  return fma(double(e), c_hitk[0], t.y) + ln1p_small(fma(m, t.x, -1.0));
}

// ---- ziggurat normal for the S1 hit loop ---------------------------------------------------
// Marsaglia & Tsang's ziggurat (J. Stat. Soft. 5(8), 2000) with 2048 layers of equal area, layer
// index and abscissa taken from DISJOINT bits as Doornik (2005) requires: an exact sampler of
// N(0,1) -- 99.77 % of the draws are one table look-up, one integer compare and one multiply;
// the rest (wedges, tail) are finished out of line with exact exp / log tests (zig_finish).
// (256 layers leave 1.5 % for the out-of-line path, which then costs as much as the core, 1024 leave
// 0.43 %, 2048 0.23 % at 24 KB of shared memory per block: see DESIGN.md section 5 for the measurements.)
// Layer i covers |x| < x_i at heights f(x_i)..f(x_{i+1}), x_0 = V/f(R) (base strip + tail),
// x_1 = R, x_2048 = 0.  Tables are built by nb200_create (host, double) and staged in shared
// memory by k_hits:  zk[i] = floor(2^28 x_{i+1}/x_i), zw[i] = x_i / 2^28.
// Abscissa: t = 8 s + 4 with s the signed 26-bit field at the top of the word (symmetric, no atom
// at zero), x = t zw[i]; it lies inside the next layer (certain accept) iff |t| < zk[i].
#define NB_ZIG_N (1 << NB_ZIG_BITS)
__device__ uint32_t g_zig_k[NB_ZIG_N];
__device__ double g_zig_w[NB_ZIG_N];
__device__ double g_zig_f[NB_ZIG_N + 1];  // f(x_i) = exp(-x_i^2/2), f[N] = 1
__device__ double g_zig_R;

static __constant__ double c_zigk[1] = {4503601774854144.0};  // 2^52 + 2^31 (constant-bank operand)
// abscissa from the top 27 bits of the word: t = 8 s + 4 with s the signed 26-bit field -- symmetric, no atom at
// zero, on the scale of the 28-bit tables above (|t| < 2^28).  Bits 4..1 of the word are free for the caller
// (the hit loop's truncation pre-test), bit 0 is unused.
NB_D int32_t zig_t(uint32_t w0) { return ((int32_t(w0) >> 3) & ~7) | 4; }
// double(t) without the conversion unit: 2^52 + 2^31 + t has t in its low mantissa bits
NB_D double zig_td(int32_t t) {
  return __hiloint2double(0x43300000, t ^ int32_t(0x80000000)) - c_zigk[0];
}
// fast path: z is the normal iff the return value is true.  w0: abscissa word, i: layer index
NB_D bool zig_fast(uint32_t w0, uint32_t i, const uint32_t* zk, const double* zw, double& z) {
  const int32_t t = zig_t(w0);
  z = zig_td(t) * zw[i];
  return uint32_t(abs(t)) < zk[i];
}
// everything else: wedge test of (i, t); on rejection a fresh layer and abscissa; layer 0 beyond
// R is the tail (Marsaglia 1964: x = -ln(u1)/R, y = -ln(u2), accept when 2y > x^2).  The first
// wedge test takes its uniform from `spare` (an unused word of the caller's block) when has_spare;
// further bits come from blocks (event, k, stream | attempt << 8): words x,y decide a pending
// (i, t), words z,w are the next (i, t).
NB_D bool zig_wedge(uint32_t i, int32_t t, double u, double& x) {
  x = zig_td(t) * g_zig_w[i];
  const double f0 = g_zig_f[i], f1 = g_zig_f[i + 1];
  return fma(u, f1 - f0, f0) < exp(-0.5 * x * x);
}
__device__ __noinline__ double zig_finish(uint32_t w0, uint32_t i, uint32_t spare, bool has_spare, uint32_t e0,
                                          uint32_t e1, uint32_t k, uint32_t stream, uint32_t k0, uint32_t k1) {
  int32_t t = zig_t(w0);
  bool have = true;  // (i, t) awaits its wedge / tail decision
  double x;
  if (has_spare && i != 0u) {
    if (zig_wedge(i, t, u32d(spare), x)) return x;
    have = false;
  }
#pragma unroll 1
  for (uint32_t att = 0; att < 4096u; ++att) {
    const u128 b = philox_block(e0, e1, k, stream | (att << 8), k0, k1);
    if (have) {
      if (i == 0u) {
        const double R = g_zig_R;
        x = neg_log_bits_g(b.x, b.y) / R;
        const double y = neg_log_bits_g(b.z, b.w);
        if (y + y > x * x) return t < 0 ? -(R + x) : (R + x);
        continue;  // not a tail sample yet: draw the tail again
      }
      if (zig_wedge(i, t, u53(b.x, b.y), x)) return x;
    }
    t = zig_t(b.z);
    i = b.w >> (32 - NB_ZIG_BITS);
    if (uint32_t(abs(t)) < g_zig_k[i]) return zig_td(t) * g_zig_w[i];
    const bool decided = !have && i != 0u;  // words x,y of this block are still unused
    have = true;
    if (decided) {
      if (zig_wedge(i, t, u53(b.x, b.y), x)) return x;
      have = false;
    }
  }
  return 0.;
}

// The hit loop packs THREE photo-electrons into one Philox block (128 bits): each takes 27 bits of abscissa,
// 4 bits of truncation pre-test and NB_ZIG_BITS (<= 11) bits of layer index.  Field `which` of block r:
// abscissa word A (top 27 bits abscissa, bits 4..1 pre-test, bit 0 unused) and layer index L.
//   0: A = x,                    L = top bits of y
//   1: A = y << B | z >> (32-B), L = the next B bits of z
//   2: A = z << 2B | w >> (32-2B), L = the low B bits of w (bit B-1 of w is A's unused bit 0 when B = 11)
NB_D void pe_fields(const u128& r, int which, uint32_t& A, uint32_t& L) {
  constexpr int B = NB_ZIG_BITS;
  static_assert(B <= 11, "three photo-electron fields must fit one Philox block");
  constexpr uint32_t M = (1u << B) - 1u;
  if (which == 0) {
    A = r.x;
    L = r.y >> (32 - B);
  } else if (which == 1) {
    A = __funnelshift_l(r.z, r.y, B);
    L = (r.z >> (32 - 2 * B)) & M;
  } else {
    A = __funnelshift_l(r.w, r.z, 2 * B);
    L = r.w & M;
  }
}
#define NB_PE_PRETEST_MASK 0x1Eu  // bits 4..1 of the abscissa word: all set = the 1-in-16 that needs the exact test

// ---- samplers (device) -------------------------------------------------------------

// two independent standard normals from one Box-Muller transform.  The reference's
// rand_gauss (RandomGen.cpp:44-53) evaluates mean + sigma*sqrt2*sqrt(-ln u)*cos(2 pi v)
// and throws the sine partner away; the partner is an independent N(0,1), so using it
// halves the log/sqrt work without changing any distribution.
// Out of line (shared by the many call sites of the per-event stages), and built from the random
// bits like the hit loop's: radius from 64 bits, angle from 32.
__device__ __noinline__ double2 box_muller(uint32_t a, uint32_t b, uint32_t c) {
  const double r = sqrt(2.0 * neg_log_bits_g(a, b));
  double s, co;
  sincos_bits(c, s, co);
  return make_double2(r * co, r * s);
}
NB_D void normal_pair(Stream& g, double& z0, double& z1) {
  uint32_t a, b, c, d;
  g.next2(a, b);
  g.next2(c, d);
  double2 z = box_muller(a, b, c);
  z0 = z.x;
  z1 = z.y;
}
// the same, also handing out the block's unused fourth word (a 32-bit uniform for a yes/no draw)
NB_D void normal_pair_spare(Stream& g, double& z0, double& z1, uint32_t& spare) {
  uint32_t a, b, c;
  g.next2(a, b);
  g.next2(c, spare);
  double2 z = box_muller(a, b, c);
  z0 = z.x;
  z1 = z.y;
}
NB_D double normal(Stream& g) {
  uint32_t a, b, c, d;
  g.next2(a, b);
  g.next2(c, d);
  return box_muller(a, b, c).x;
}
// rand_gauss(mean, sigma, zero_min) RandomGen.cpp:44-53
NB_D double gauss(Stream& g, double mean, double sigma, bool zero_min) {
  double d = mean + sigma * normal(g);
  // max(draw, 0.) with std::max semantics: NaN draw stays NaN (max(a,b) = a<b ? b : a)
  if (zero_min) return (d < 0.) ? 0. : d;
  return d;
}
// rand_zero_trunc_gauss RandomGen.cpp:55-61 (redraw while <= 0)
NB_D double zero_trunc_gauss(Stream& g, double mean, double sigma) {
  double r = mean + sigma * normal(g);
  int guard = 0;
  while (r <= 0. && ++guard < 100000) r = mean + sigma * normal(g);
  return r;
}
// rand_exponential RandomGen.cpp:72-85
NB_D double exponential(Stream& g, double half_life, double t_min, double t_max) {
  const double ln2 = 0.693147180559945309417;
  double r = g.uniform();
  double r_min = 0., r_max = 1.;
  if (t_min >= 0) r_min = 1 - exp(-t_min * ln2 / half_life);
  if (t_max >= 0) r_max = 1 - exp(-t_max * ln2 / half_life);
  r = (r_max - r_min) * r + r_min;
  return -log(1 - r) * half_life / ln2;
}
// rand_skewGauss RandomGen.cpp:87-126.  The reference samples the skew-normal density
// 2/omega phi(z) Phi(alpha z), z=(x-xi)/omega, by box rejection over xi +- 6 omega
// (acceptance ~14 %).  Same distribution, exact: z = delta|a| + sqrt(1-delta^2) b with
// a,b iid N(0,1), delta = alpha/sqrt(1+alpha^2) (Azzalini 1985), redrawn when |z|>6 to
// keep the reference's +-6 omega support.
NB_D double skew_gauss(Stream& g, double xi, double omega, double alpha) {
  double delta = alpha / sqrt(1. + alpha * alpha);
  double cd = sqrt(1. - delta * delta);
  double z;
  int guard = 0;
  do {
    double a, b;
    normal_pair(g, a, b);
    z = delta * fabs(a) + cd * b;
  } while (fabs(z) > 6. && ++guard < 1000);
  return xi + omega * z;
}
// SelectRanXeAtom RandomGen.cpp:159-182 (natural-abundance cumulative table in percent)
NB_D int select_xe_atom(Stream& g) {
  double iso = g.uniform() * 100.;
  if (iso <= 0.090) return 124;
  if (iso <= 0.180) return 126;
  if (iso <= 2.100) return 128;
  if (iso <= 28.54) return 129;
  if (iso <= 32.62) return 130;
  if (iso <= 53.80) return 131;
  if (iso <= 80.69) return 132;
  if (iso <= 91.13) return 134;
  return 136;
}

// Stirling-series tail log(k!) - [ (k+1/2)log(k+1) - (k+1) + log(2 pi)/2 ]
static __constant__ double c_stirling[10] = {
    0.08106146679532726, 0.04134069595540929, 0.02767792568499834, 0.02079067210376509,
    0.01664469118982119, 0.01387612882307075, 0.01189670994589177, 0.01041126526197209,
    0.009255462182712733, 0.008330563433362871};
NB_D double stirling_tail(double k) {
  if (k < 10.) return c_stirling[int(k)];
  const double rk = 1.0 / (k + 1.);
  const double rk2 = rk * rk;
  return (1.0 / 12. - (1.0 / 360. - (1.0 / 1260.) * rk2) * rk2) * rk;
}

// 1/x for x = 0..255 (filled by nb200_create): the CDF-inversion recurrence of the binomial sampler
// multiplies instead of dividing
#define NB_RCP_N 256
__device__ double g_rcptab[NB_RCP_N];

// position of the n-th (0-based) set bit of m; n < popc(m)
NB_D int nth_set_bit(unsigned m, int n) {
  int pos = 0, t;
  t = __popc(m & 0xFFFFu);
  if (n >= t) { n -= t; pos += 16; m >>= 16; }
  t = __popc(m & 0xFFu);
  if (n >= t) { n -= t; pos += 8; m >>= 8; }
  t = __popc(m & 0xFu);
  if (n >= t) { n -= t; pos += 4; m >>= 4; }
  t = __popc(m & 0x3u);
  if (n >= t) { n -= t; pos += 2; m >>= 2; }
  if (n >= int(m & 1u)) pos += 1;
  return pos;
}

// ln(k!) for k < NB_LFACT_N, filled by nb200_create (host lgammal): the exact acceptance test of the binomial
// sampler reads its four log-factorials instead of evaluating four logarithms and four Stirling tails
#define NB_LFACT_N (1 << 17)
__device__ double g_lfact[NB_LFACT_N];
NB_D double log_factorial(double k) {
  if (k < double(NB_LFACT_N)) return __ldg(&g_lfact[int(k)]);
  // Stirling: ln k! = (k + 1/2) ln(k + 1) - (k + 1) + ln(2 pi)/2 + tail(k)
  return fma(k + 0.5, log_pos(k + 1.), -(k + 1.)) + 0.91893853320467274178 + stirling_tail(k);
}

// Hormann's BTRS (see binomial below) in two pieces, so that callers that regroup their draws (k_lar: one candidate
// per lane and pass, rejected draws re-queued) run the very statements of the in-line loop: the per-draw constants,
// and one candidate from one Philox block.  pp = min(p, 1 - p), q = 1 - pp, n pp >= 10.
struct Btrs {
  double dn, b, a, c, vr, alpha, m, h, logr;
};
NB_D Btrs btrs_setup(double dn, double pp, double q) {
  Btrs s;
  s.dn = dn;
  const double spq = sqrt(dn * pp * q);
  s.b = 1.15 + 2.53 * spq;
  const double rb = 1. / s.b;
  s.a = -0.0873 + 0.0248 * s.b + 0.01 * pp;
  s.c = dn * pp + 0.5;
  s.vr = 0.92 - 4.2 * rb;
  s.alpha = (2.83 + 5.1 * rb) * spq;
  s.m = floor((dn + 1.) * pp);
  // the per-draw terms of the exact test: loaded here so that their (L2) latency is hidden behind the first candidate
  s.h = log_factorial(s.m) + log_factorial(dn - s.m);
  s.logr = log_pos(pp / q);
  return s;
}
// candidate from block w: true when accepted (kk: the variate for pp)
NB_D bool btrs_try(const Btrs& s, const u128& w, double& kk) {
  const double u = u53(w.x, w.y) - 0.5;
  const double v = u53(w.z, w.w);
  const double us = 0.5 - fabs(u);
  kk = floor((2. * s.a / us + s.b) * u + s.c);
  if (us >= 0.07 && v <= s.vr) return true;
  if (kk < 0. || kk > s.dn) return false;
  const double lk = log_factorial(kk), lnk = log_factorial(s.dn - kk);  // issued before the logarithms below
  const double us2 = us * us;
  // ln( v alpha / (a/us^2 + b) )
  const double lhs = log_pos(v * s.alpha * us2) - log_pos(s.a + s.b * us2);
  const double rhs = (s.h - lk) - lnk + (kk - s.m) * s.logr;
  return lhs <= rhs;
}

// binom_draw RandomGen.cpp:132-134 = std::binomial_distribution<int64_t>(n,p).  Exact binomial sampler:
// sequential CDF inversion while n*min(p,1-p) < 10, Hormann's BTRS transformed rejection (J. Stat. Comput.
// Simul. 46 (1993) 101) above.  Draws from its own counter stream (seed; event, stream): candidate k of a
// BTRS draw is Philox block k of that stream.  In line at its call sites (with the block function): the
// scalar stages run their blocks in step, so code size costs nothing, and the inlined Philox reads its round
// keys as constant operands (measured against the out-of-line version: k_pre -3.5 %, k_post -4 %).
//
// BTRS accepts ~75 % of its candidates on a cheap test; the others need the exact log-pmf ratio
//     ln f(k)/f(m) = ln m! + ln (n-m)! - ln k! - ln (n-k)! + (k - m) ln(p/q),   m = floor((n+1) p).
// Round 1 evaluated it with four logarithms and four Stirling tails (~330 instructions, shared out over the
// lanes of the warp: ~1200 warp-instructions per call, 35 % of k_post at 19 active lanes).  Here the
// log-factorials come from a table (g_lfact: 1 MB, L2-resident; the m-terms and ln(p/q) once per draw), so a
// test is two loads, one logarithm and a handful of FMAs, cheap enough to run per lane.  The table holds
// ln k! to 1e-16 relative; the sum of four terms of up to 1e6 carries ~1e-9 of absolute error, i.e. the
// acceptance probability of a candidate is exact to ~1e-9 relative (tests/test_gpu_samplers.py: chi^2
// against the exact pmf, n up to 1.2e6).
NB_D long long binomial(const uint32_t* rk, uint64_t event, uint32_t stream, long long n, double p) {
  if (!(n > 0) || !(p > 0.)) return 0;
  if (p >= 1.) return n;
  const bool flip = p > 0.5;
  const double pp = flip ? 1. - p : p;
  const double q = 1. - pp;
  const double dn = double(n);
  const uint32_t e0 = uint32_t(event), e1 = uint32_t(event >> 32);
  long long k = 0;
  if (dn * pp < 10.) {
    // q^n = exp(n ln q), q in [1/2, 1): the bit-built logarithm (absolute error ~1e-16, times n < 10/pp)
    const double r0 = exp(dn * log_pos(q));
    const double s = pp / q;
    for (uint32_t att = 0;; ++att) {  // one uniform per attempt: words (x,y), then (z,w) of block att/2
      const u128 w = philox4x32_10_rk(e0, e1, att >> 1, stream, rk);
      double u = (att & 1u) ? u53(w.z, w.w) : u53(w.x, w.y);
      double r = r0;
      // x counts up as an int, n - x + 1 down as a double (integers below 2^53: the same values as the
      // 64-bit integer arithmetic it replaces, without its conversions)
      int x = 0;
      double left = dn;  // n - x
      bool ok = true;
      while (u > r) {
        u -= r;
        ++x;
        if (left < 1. || x > 200) {  // x > n
          ok = false;
          break;
        }
        r *= s * (left * __ldg(&g_rcptab[x]));  // x <= 200
        left -= 1.;
      }
      if (ok) {
        k = x;
        break;
      }
    }
  } else {
    const Btrs B = btrs_setup(dn, pp, q);
    double kk = 0.;
    for (uint32_t att = 0;; ++att) {
      const u128 w = philox4x32_10_rk(e0, e1, att, stream, rk);
      if (btrs_try(B, w, kk)) break;
    }
    k = (long long)kk;
  }
  return flip ? n - k : k;
}

// Binomial(n, p) for a probability that is fixed for the whole run (P_dphe): n is split into its binary
// digits, n = 4096 q + sum of 2^b, and Binomial(2^b, p) is drawn from an alias table (Walker / Vose) built
// on the host for that p -- one 32-bit word, one multiply, one 8-byte load and one compare per digit,
// q + <= 12 draws in all (a 79-hit S1: 5; a 2300-hit S2: 6), against the ~1500 warp-instructions of a
// divergent BTRS call.  Table b (b = 0..12) holds 2^b + 1 entries {threshold on the 32-bit fraction,
// alias} at offset 2^b - 1 + b.  Outcome probabilities are exact to 2^-32 (the resolution of the other
// yes/no draws of the chain).
NB_D int alias_draw(const uint2* __restrict__ t, uint32_t K, uint32_t w) {
  const uint64_t pr = uint64_t(w) * K;
  const uint32_t idx = uint32_t(pr >> 32);
  const uint2 en = __ldg(t + idx);
  return uint32_t(pr) < en.x ? int(idx) : int(en.y);
}
NB_D int binomial_fixed_p(const uint32_t* rk, uint64_t event, uint32_t stream, int n,
                          const uint2* __restrict__ tab) {
  Stream g;
  g.init(rk, event, stream);
  uint32_t a = 0u, b = 0u;
  bool have = false;
  int k = 0;
#pragma unroll 1
  for (int q = n >> NB_ALIAS_BITS; q > 0; --q) {
    if (!have) g.next2(a, b);
    k += alias_draw(tab + ((1 << NB_ALIAS_BITS) - 1 + NB_ALIAS_BITS), (1u << NB_ALIAS_BITS) + 1u, have ? b : a);
    have = !have;
  }
#pragma unroll 1
  for (int rest = n & ((1 << NB_ALIAS_BITS) - 1); rest != 0; rest &= rest - 1) {  // its set bits, lowest first
    const int bit = __ffs(rest) - 1;
    if (!have) g.next2(a, b);
    k += alias_draw(tab + ((1 << bit) - 1 + bit), (1u << bit) + 1u, have ? b : a);
    have = !have;
  }
  return k;
}
#endif  // __CUDACC__

// host: the ziggurat of the hit loop's normal sampler.  x[0..N]: x_0 = V/f(R) (base strip incl. tail),
// x_1 = R, ..., x_N = 0; f[i] = exp(-x_i^2/2) (f[0] unused, f[N] = 1); k[i] = floor(2^28 x_{i+1}/x_i),
// w[i] = x_i / 2^28.  R is solved by bisection so that N layers of equal area V = R f(R) + tail(R)
// close exactly at f = 1.
inline void build_zig_tables(double* zx, double* zf, uint32_t* zk, double* zw) {
  const int N = 1 << NB_ZIG_BITS;
  auto f = [](double x) { return std::exp(-0.5 * x * x); };
  auto fill = [&](double R) -> double {  // returns f at the top minus 1 (> 0: R too small)
    const double V = R * f(R) + 1.2533141373155002512 * std::erfc(R * 0.70710678118654752440);
    zx[0] = V / f(R);
    zx[1] = R;
    for (int i = 1; i < N; ++i) {
      const double fi = f(zx[i]) + V / zx[i];
      if (i == N - 1) return fi - 1.;
      if (fi >= 1.) return double(N - i);
      zx[i + 1] = std::sqrt(-2. * std::log(fi));
    }
    return 0.;
  };
  double lo = 3.0, hi = 5.0;
  for (int it = 0; it < 200; ++it) {
    const double mid = 0.5 * (lo + hi);
    (fill(mid) > 0. ? lo : hi) = mid;
  }
  fill(hi);
  zx[N] = 0.;
  for (int i = 1; i <= N; ++i) zf[i] = f(zx[i]);
  zf[0] = 0.;
  zf[N] = 1.;
  for (int i = 0; i < N; ++i) {
    zk[i] = uint32_t(std::floor(268435456.0 * (zx[i + 1] / zx[i])));
    zw[i] = zx[i] / 268435456.0;
  }
}

// host: alias tables of Binomial(2^b, p), b = 0..12, for binomial_fixed_p (0 < p < 1)
struct AliasEntry {
  uint32_t thr, alias;
};
inline void build_binomial_alias(double p, AliasEntry* out /* NB_ALIAS_ENTRIES */) {
  const int KMAX = (1 << NB_ALIAS_BITS) + 1;
  double* pmf = new double[KMAX];
  int* small = new int[KMAX];
  int* large = new int[KMAX];
  const double lp = std::log(p), lq = std::log1p(-p);
  for (int b = 0; b <= NB_ALIAS_BITS; ++b) {
    const int m = 1 << b, K = m + 1;
    AliasEntry* t = out + (m - 1 + b);
    double tot = 0.;
    for (int k = 0; k < K; ++k) {  // log space: (1-p)^4096 underflows
      pmf[k] = std::exp(std::lgamma(double(m) + 1.) - std::lgamma(double(k) + 1.) - std::lgamma(double(m - k) + 1.) +
                        double(k) * lp + double(m - k) * lq);
      tot += pmf[k];
    }
    int ns = 0, nl = 0;
    for (int k = 0; k < K; ++k) {
      pmf[k] = pmf[k] / tot * K;
      (pmf[k] < 1. ? small[ns++] : large[nl++]) = k;
    }
    for (int k = 0; k < K; ++k) t[k] = AliasEntry{0xFFFFFFFFu, uint32_t(k)};
    while (ns > 0 && nl > 0) {
      const int s_ = small[--ns], l_ = large[--nl];
      const double v = pmf[s_] * 4294967296.0;
      t[s_].thr = v >= 4294967295.0 ? 0xFFFFFFFFu : uint32_t(v);
      t[s_].alias = uint32_t(l_);
      pmf[l_] = (pmf[l_] + pmf[s_]) - 1.;
      (pmf[l_] < 1. ? small[ns++] : large[nl++]) = l_;
    }
  }
  delete[] pmf;
  delete[] small;
  delete[] large;
}

}  // namespace nb
