// BN254 Fr / Fq arithmetic for sm_100a: 8 x 32-bit limbs, Montgomery form, canonical outputs.
//
// Replaces (reference, /root/reference): MontgomeryField / BN254Scalar in
// src/field/mod.zig:124-556, :583-1052 -- same representation on the wire
// ([4]u64 little-endian limbs == 8 x u32 little-endian limbs, Montgomery form a*2^256 mod p,
// always < p) so tables and points cross the C ABI without conversion.
//
// Multiplication is an interleaved (CIOS) Montgomery product on mad.lo.cc / madc.hi.cc carry
// chains.  Products of even and odd limbs of the multiplicand accumulate in two separate
// registers sets ("E" aligned at 2^0 and "O" aligned at 2^32) so that every lo/hi pair is one
// uninterrupted carry chain; ptxas fuses each pair into a single IMAD.WIDE.U32(.X) -- 128 wide
// multiplies + 8 IMAD per product (checked with cuobjdump).
//
// The primitive layer below has two back-ends with identical instruction sequences:
//   * device  (__CUDA_ARCH__): one PTX instruction per primitive, carry in the hardware CC flag;
//   * host    : exact emulation of the PTX carry semantics (thread-local CC bit).  The host
//     back-end exists ONLY so tests/ can validate the algorithms on a machine without a GPU
//     (tests/host_shim); no product entry point ever runs field math on the host.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define ZOLT_HD __host__ __device__ __forceinline__
#define ZOLT_D __device__ __forceinline__
#else
#define ZOLT_HD inline
#define ZOLT_D inline
#endif

namespace zolt {

// ---------------------------------------------------------------------------------------------
// carry-chain primitives
// ---------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
namespace ptx {
ZOLT_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZOLT_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZOLT_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZOLT_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZOLT_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZOLT_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZOLT_D uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZOLT_D uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZOLT_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZOLT_D uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZOLT_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZOLT_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZOLT_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
}  // namespace ptx
#else
namespace ptx {
// Host emulation of the PTX integer carry flag (PTX ISA "Extended-Precision Integer Arithmetic").
inline uint32_t& cc() { static thread_local uint32_t flag = 0; return flag; }
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; cc() = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b + cc(); cc() = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t addc(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a + b + cc()); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b; cc() = (uint32_t)((s >> 32) & 1); return (uint32_t)s; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b - cc(); cc() = (uint32_t)((s >> 32) & 1); return (uint32_t)s; }
inline uint32_t subc(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a - b - cc()); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_lo(a, b), c); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_hi(a, b), c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_lo(a, b), c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc(mul_hi(a, b), c); }
}  // namespace ptx
#endif

// ---------------------------------------------------------------------------------------------
// field parameters (reference src/field/mod.zig:16-41 Fr, :51-75 Fq; 64-bit limbs split in two)
// ---------------------------------------------------------------------------------------------
// Constants are exposed through constexpr accessors with function-local tables: static constexpr
// array members are not addressable from device code, while a local constexpr table indexed by an
// unrolled loop counter folds to immediates.
#define ZOLT_FIELD_TABLE(name, ...)                            \
  static constexpr ZOLT_HD uint32_t name(int i) {              \
    constexpr uint32_t t[8] = {__VA_ARGS__};                   \
    return t[i];                                               \
  }

struct FrParams {
  // r = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
  ZOLT_FIELD_TABLE(mod, 0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u,
                   0xe131a029u, 0x30644e72u)
  // R = 2^256 mod r  (== BN254Scalar.one())
  ZOLT_FIELD_TABLE(one, 0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu,
                   0x9a07df2fu, 0x0e0a77c1u)
  // R^2 mod r
  ZOLT_FIELD_TABLE(r2, 0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du,
                   0x7f4e44a5u, 0x0216d0b1u)
  // 2r (subtracted from the top half of a 512-bit accumulator by wide_normalize)
  ZOLT_FIELD_TABLE(mod2, 0xe0000002u, 0x87c3eb27u, 0xf372e122u, 0x5067d090u, 0x0302b0bau, 0x70a08b6du,
                   0xc2634053u, 0x60c89ce5u)
  // 1/2 and 1/6 in Montgomery form (Newton interpolation of the previous round polynomial)
  ZOLT_FIELD_TABLE(inv2, 0x1ffffffeu, 0x783c14d8u, 0x0c8d1eddu, 0xaf982f6fu, 0xfcfd4f45u, 0x8f5f7492u,
                   0x3d9cbfacu, 0x1f37631au)
  ZOLT_FIELD_TABLE(inv6, 0x0aaaaaaau, 0x7d695c48u, 0xaed9b4f4u, 0x3a880fcfu, 0xa9a9c517u, 0xda7526dbu,
                   0x69deea8eu, 0x0a67cbb3u)
  static constexpr uint32_t INV = 0xefffffffu;  // -r^-1 mod 2^32 (low word of 0xc2e1f593efffffff)
  static constexpr unsigned long long MU320_HI = 0x5ull, MU320_LO = 0x4a47462623a04a7aull;  // floor(2^320 / r) (Barrett, fp_fixed_mul_add)
};

struct FqParams {
  // q = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
  ZOLT_FIELD_TABLE(mod, 0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u,
                   0xe131a029u, 0x30644e72u)
  ZOLT_FIELD_TABLE(one, 0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu,
                   0x9a07df2fu, 0x0e0a77c1u)
  ZOLT_FIELD_TABLE(r2, 0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu,
                   0xcab8351fu, 0x06d89f71u)
  ZOLT_FIELD_TABLE(mod2, 0xb0f9fa8eu, 0x7841182du, 0xd0e3951au, 0x2f02d522u, 0x0302b0bbu, 0x70a08b6du,
                   0xc2634053u, 0x60c89ce5u)
  ZOLT_FIELD_TABLE(inv2, 0x4f060572u, 0x87bee7d2u, 0x2f1c6ae5u, 0xd0fd2addu, 0xfcfd4f44u, 0x8f5f7492u,
                   0x3d9cbfacu, 0x1f37631au)
  ZOLT_FIELD_TABLE(inv6, 0x1a575726u, 0xd7ea4d46u, 0x0fb423a1u, 0xf054639fu, 0xa9a9c516u, 0xda7526dbu,
                   0x69deea8eu, 0x0a67cbb3u)
  static constexpr uint32_t INV = 0xe4866389u;  // -q^-1 mod 2^32 (low word of 0x87d20782e4866389)
  static constexpr unsigned long long MU320_HI = 0x5ull, MU320_LO = 0x4a47462623a04a7aull;  // floor(2^320 / q)
};

// ---------------------------------------------------------------------------------------------
// field element
// ---------------------------------------------------------------------------------------------
template <class P>
struct alignas(16) Fp {
  uint32_t v[8];

  static ZOLT_HD Fp zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
  }
  static ZOLT_HD Fp one() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::one(i);
    return r;
  }
  static ZOLT_HD Fp r2() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::r2(i);
    return r;
  }
  ZOLT_HD bool is_zero() const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= v[i];
    return o == 0;
  }
  ZOLT_HD bool operator==(const Fp& b) const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= v[i] ^ b.v[i];
    return o == 0;
  }
  ZOLT_HD bool operator!=(const Fp& b) const { return !(*this == b); }
};

// r = a - MOD if a >= MOD else a   (a < 2*MOD)
template <class P>
ZOLT_HD void fp_reduce_once(uint32_t* a) {
  uint32_t s[8];
  s[0] = ptx::sub_cc(a[0], P::mod(0));
#pragma unroll
  for (int i = 1; i < 8; i++) s[i] = ptx::subc_cc(a[i], P::mod(i));
  uint32_t borrow = ptx::subc(0u, 0u);  // 0xffffffff if a < MOD
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = borrow ? a[i] : s[i];
}

// canonical a + b  (reference add, src/field/mod.zig:782-798)
template <class P>
ZOLT_HD Fp<P> fp_add(const Fp<P>& a, const Fp<P>& b) {
  Fp<P> r;
  r.v[0] = ptx::add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = ptx::addc_cc(a.v[i], b.v[i]);
  r.v[7] = ptx::addc(a.v[7], b.v[7]);  // a,b < 2^254: no carry out of limb 7
  fp_reduce_once<P>(r.v);
  return r;
}

// a + b WITHOUT the final conditional subtraction: a value in [0, 2p) < 2^255 for canonical a, b.  Only for operands of
// a product whose other factor keeps the result in range: mont_mul_raw accepts a < 2^255 with a * b < p * 2^256, the wide
// accumulators accept factors below 4p (see their invariants below).
template <class P>
ZOLT_HD Fp<P> fp_add_lazy(const Fp<P>& a, const Fp<P>& b) {
  Fp<P> r;
  r.v[0] = ptx::add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = ptx::addc_cc(a.v[i], b.v[i]);
  r.v[7] = ptx::addc(a.v[7], b.v[7]);
  return r;
}

// canonical a - b  (reference sub, src/field/mod.zig:801-816)
template <class P>
ZOLT_HD Fp<P> fp_sub(const Fp<P>& a, const Fp<P>& b) {
  Fp<P> r;
  r.v[0] = ptx::sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) r.v[i] = ptx::subc_cc(a.v[i], b.v[i]);
  uint32_t borrow = ptx::subc(0u, 0u);  // all-ones when a < b
  r.v[0] = ptx::add_cc(r.v[0], P::mod(0) & borrow);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = ptx::addc_cc(r.v[i], P::mod(i) & borrow);
  r.v[7] = ptx::addc(r.v[7], P::mod(7) & borrow);
  return r;
}

template <class P>
ZOLT_HD Fp<P> fp_neg(const Fp<P>& a) {
  return fp_sub<P>(Fp<P>::zero(), a);
}

template <class P>
ZOLT_HD Fp<P> fp_dbl(const Fp<P>& a) {
  return fp_add<P>(a, a);
}

// ---------------------------------------------------------------------------------------------
// Montgomery product, interleaved reduction, even/odd accumulators.
//
// Invariant between rows: T = E + 2^32 * O, with E in e[0..8] (9 limbs) and O in o[0..7].
// For a < 2^255 (any b): O < 2^256 (its chain never carries out) and E < 2^288.
// After a row's reduction e[0] == 0 and the division by 2^32 is done by exchanging roles:
//   E' = O + e[1]            (carry of the single limb add feeds the O' chain)
//   O' = e[2..8] + a_odd*b_i
// ---------------------------------------------------------------------------------------------
template <class P>
ZOLT_HD void mont_row_reduce(uint32_t* e, uint32_t* o) {
  const uint32_t m = ptx::mul_lo(e[0], P::INV);
  e[0] = ptx::mad_lo_cc(P::mod(0), m, e[0]);
  e[1] = ptx::madc_hi_cc(P::mod(0), m, e[1]);
  e[2] = ptx::madc_lo_cc(P::mod(2), m, e[2]);
  e[3] = ptx::madc_hi_cc(P::mod(2), m, e[3]);
  e[4] = ptx::madc_lo_cc(P::mod(4), m, e[4]);
  e[5] = ptx::madc_hi_cc(P::mod(4), m, e[5]);
  e[6] = ptx::madc_lo_cc(P::mod(6), m, e[6]);
  e[7] = ptx::madc_hi_cc(P::mod(6), m, e[7]);
  e[8] = ptx::addc(e[8], 0u);
  o[0] = ptx::mad_lo_cc(P::mod(1), m, o[0]);
  o[1] = ptx::madc_hi_cc(P::mod(1), m, o[1]);
  o[2] = ptx::madc_lo_cc(P::mod(3), m, o[2]);
  o[3] = ptx::madc_hi_cc(P::mod(3), m, o[3]);
  o[4] = ptx::madc_lo_cc(P::mod(5), m, o[4]);
  o[5] = ptx::madc_hi_cc(P::mod(5), m, o[5]);
  o[6] = ptx::madc_lo_cc(P::mod(7), m, o[6]);
  o[7] = ptx::madc_hi(P::mod(7), m, o[7]);
}

// First row: T = a * b0.
template <class P>
ZOLT_HD void mont_row_first(uint32_t* e, uint32_t* o, const uint32_t* a, uint32_t bi) {
#pragma unroll
  for (int t = 0; t < 4; t++) {
    o[2 * t] = ptx::mul_lo(a[2 * t + 1], bi);
    o[2 * t + 1] = ptx::mul_hi(a[2 * t + 1], bi);
    e[2 * t] = ptx::mul_lo(a[2 * t], bi);
    e[2 * t + 1] = ptx::mul_hi(a[2 * t], bi);
  }
  e[8] = 0;
  mont_row_reduce<P>(e, o);
}

// Later rows.  `e` is the accumulator that was O in the previous row (8 live limbs, e[8] is
// scratch), `o` is the previous row's E (9 live limbs, o[0] == 0).
template <class P>
ZOLT_HD void mont_row_next(uint32_t* e, uint32_t* o, const uint32_t* a, uint32_t bi) {
  e[0] = ptx::add_cc(e[0], o[1]);
  o[0] = ptx::madc_lo_cc(a[1], bi, o[2]);
  o[1] = ptx::madc_hi_cc(a[1], bi, o[3]);
  o[2] = ptx::madc_lo_cc(a[3], bi, o[4]);
  o[3] = ptx::madc_hi_cc(a[3], bi, o[5]);
  o[4] = ptx::madc_lo_cc(a[5], bi, o[6]);
  o[5] = ptx::madc_hi_cc(a[5], bi, o[7]);
  o[6] = ptx::madc_lo_cc(a[7], bi, o[8]);
  o[7] = ptx::madc_hi(a[7], bi, 0u);
  e[0] = ptx::mad_lo_cc(a[0], bi, e[0]);
  e[1] = ptx::madc_hi_cc(a[0], bi, e[1]);
  e[2] = ptx::madc_lo_cc(a[2], bi, e[2]);
  e[3] = ptx::madc_hi_cc(a[2], bi, e[3]);
  e[4] = ptx::madc_lo_cc(a[4], bi, e[4]);
  e[5] = ptx::madc_hi_cc(a[4], bi, e[5]);
  e[6] = ptx::madc_lo_cc(a[6], bi, e[6]);
  e[7] = ptx::madc_hi_cc(a[6], bi, e[7]);
  e[8] = ptx::addc(0u, 0u);
  mont_row_reduce<P>(e, o);
}

// r = a * b * 2^-256 mod p.  Requires a < 2^255 (b arbitrary 256-bit).  Output in [0, 2p) when
// a*b < p * 2^256; fp_mul below makes it canonical.
template <class P>
ZOLT_HD void mont_mul_raw(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t x[9], y[9];
  mont_row_first<P>(x, y, a, b[0]);
  mont_row_next<P>(y, x, a, b[1]);
  mont_row_next<P>(x, y, a, b[2]);
  mont_row_next<P>(y, x, a, b[3]);
  mont_row_next<P>(x, y, a, b[4]);
  mont_row_next<P>(y, x, a, b[5]);
  mont_row_next<P>(x, y, a, b[6]);
  mont_row_next<P>(y, x, a, b[7]);
  // last row had E = y (y[0] == 0), O = x:  result = O + (E >> 32)
  r[0] = ptx::add_cc(x[0], y[1]);
#pragma unroll
  for (int i = 1; i < 7; i++) r[i] = ptx::addc_cc(x[i], y[i + 1]);
  r[7] = ptx::addc(x[7], y[8]);
}

// ---- two products, one reduction -------------------------------------------------------------------------
// (a * b + c * d) * 2^-256 mod p with the reduction rows shared between the two products: 8 x (8 + 8 + 8) wide
// multiplies instead of 2 x 8 x (8 + 8).  Requires a, c < 2^254 (canonical elements, or p itself); b, d arbitrary.
// Bounds: every odd-limb partial (a_odd * b_i, c_odd * d_i, m * p_odd) is below 2^254 because the top limbs of a, c, p
// are below 2^30, so O = carried-in (< 2^224) + three of them stays below 2^256 and never carries out of o[7]; E gains
// at most three carries per row in e[8].  The result is below a + c + p + 1 <= 3p.
template <class P>
ZOLT_HD void mont_row_extra(uint32_t* e, uint32_t* o, const uint32_t* c, uint32_t di) {
  o[0] = ptx::mad_lo_cc(c[1], di, o[0]);
  o[1] = ptx::madc_hi_cc(c[1], di, o[1]);
  o[2] = ptx::madc_lo_cc(c[3], di, o[2]);
  o[3] = ptx::madc_hi_cc(c[3], di, o[3]);
  o[4] = ptx::madc_lo_cc(c[5], di, o[4]);
  o[5] = ptx::madc_hi_cc(c[5], di, o[5]);
  o[6] = ptx::madc_lo_cc(c[7], di, o[6]);
  o[7] = ptx::madc_hi(c[7], di, o[7]);
  e[0] = ptx::mad_lo_cc(c[0], di, e[0]);
  e[1] = ptx::madc_hi_cc(c[0], di, e[1]);
  e[2] = ptx::madc_lo_cc(c[2], di, e[2]);
  e[3] = ptx::madc_hi_cc(c[2], di, e[3]);
  e[4] = ptx::madc_lo_cc(c[4], di, e[4]);
  e[5] = ptx::madc_hi_cc(c[4], di, e[5]);
  e[6] = ptx::madc_lo_cc(c[6], di, e[6]);
  e[7] = ptx::madc_hi_cc(c[6], di, e[7]);
  e[8] = ptx::addc(e[8], 0u);
}

// the product rows of mont_row_first / mont_row_next without their reduction step
template <class P>
ZOLT_HD void mont_row_first_noreduce(uint32_t* e, uint32_t* o, const uint32_t* a, uint32_t bi) {
#pragma unroll
  for (int t = 0; t < 4; t++) {
    o[2 * t] = ptx::mul_lo(a[2 * t + 1], bi);
    o[2 * t + 1] = ptx::mul_hi(a[2 * t + 1], bi);
    e[2 * t] = ptx::mul_lo(a[2 * t], bi);
    e[2 * t + 1] = ptx::mul_hi(a[2 * t], bi);
  }
  e[8] = 0;
}
template <class P>
ZOLT_HD void mont_row_next_noreduce(uint32_t* e, uint32_t* o, const uint32_t* a, uint32_t bi) {
  e[0] = ptx::add_cc(e[0], o[1]);
  o[0] = ptx::madc_lo_cc(a[1], bi, o[2]);
  o[1] = ptx::madc_hi_cc(a[1], bi, o[3]);
  o[2] = ptx::madc_lo_cc(a[3], bi, o[4]);
  o[3] = ptx::madc_hi_cc(a[3], bi, o[5]);
  o[4] = ptx::madc_lo_cc(a[5], bi, o[6]);
  o[5] = ptx::madc_hi_cc(a[5], bi, o[7]);
  o[6] = ptx::madc_lo_cc(a[7], bi, o[8]);
  o[7] = ptx::madc_hi(a[7], bi, 0u);
  e[0] = ptx::mad_lo_cc(a[0], bi, e[0]);
  e[1] = ptx::madc_hi_cc(a[0], bi, e[1]);
  e[2] = ptx::madc_lo_cc(a[2], bi, e[2]);
  e[3] = ptx::madc_hi_cc(a[2], bi, e[3]);
  e[4] = ptx::madc_lo_cc(a[4], bi, e[4]);
  e[5] = ptx::madc_hi_cc(a[4], bi, e[5]);
  e[6] = ptx::madc_lo_cc(a[6], bi, e[6]);
  e[7] = ptx::madc_hi_cc(a[6], bi, e[7]);
  e[8] = ptx::addc(0u, 0u);
}

// canonical a * b + c * d (Montgomery); a, c canonical (c may also be p)
template <class P>
ZOLT_HD Fp<P> fp_mul2_add(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d) {
  uint32_t x[9], y[9];
  mont_row_first_noreduce<P>(x, y, a.v, b.v[0]);
  mont_row_extra<P>(x, y, c.v, d.v[0]);
  mont_row_reduce<P>(x, y);
#define ZOLT_ROW2(E, O, I)                          \
  mont_row_next_noreduce<P>(E, O, a.v, b.v[I]);     \
  mont_row_extra<P>(E, O, c.v, d.v[I]);             \
  mont_row_reduce<P>(E, O);
  ZOLT_ROW2(y, x, 1)
  ZOLT_ROW2(x, y, 2)
  ZOLT_ROW2(y, x, 3)
  ZOLT_ROW2(x, y, 4)
  ZOLT_ROW2(y, x, 5)
  ZOLT_ROW2(x, y, 6)
  ZOLT_ROW2(y, x, 7)
#undef ZOLT_ROW2
  Fp<P> r;
  r.v[0] = ptx::add_cc(x[0], y[1]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = ptx::addc_cc(x[i], y[i + 1]);
  r.v[7] = ptx::addc(x[7], y[8]);
  fp_reduce_once<P>(r.v);  // [0, 3p) -> [0, p)
  fp_reduce_once<P>(r.v);
  return r;
}

// canonical a * b - c * d:  the subtrahend enters as (p - c) * d
template <class P>
ZOLT_HD Fp<P> fp_mul2_sub(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d) {
  Fp<P> nc;
  nc.v[0] = ptx::sub_cc(P::mod(0), c.v[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) nc.v[i] = ptx::subc_cc(P::mod(i), c.v[i]);
  nc.v[7] = ptx::subc(P::mod(7), c.v[7]);
  return fp_mul2_add<P>(a, b, nc, d);
}

// canonical Montgomery product (reference montgomeryMul, src/field/mod.zig:269-308, :735-779)
template <class P>
ZOLT_HD Fp<P> fp_mul(const Fp<P>& a, const Fp<P>& b) {
  Fp<P> r;
  mont_mul_raw<P>(r.v, a.v, b.v);
  fp_reduce_once<P>(r.v);
  return r;
}

// ---- squaring ---------------------------------------------------------------------------------------------
// a^2 = sum_i a_i * (a_i 2^(32 i) + 2 sum_{j > i} a_j 2^(32 j)) 2^(32 i): row i of the interleaved product only needs the
// partial products at limb positions k >= i, taken against V_i = (a_i at k = i, the limbs of 2a above: D[i+1] with its
// lowest bit -- the bit shifted in from a_i -- cleared, D[k] for k > i+1).  36 wide multiplies for the product part
// instead of 64; the reduction rows are those of mont_mul_raw.  Positions below i only move the accumulator down (adds).
// Odd-position partials are below 2^255 (doubled limbs) and m * p_odd below 2^254, so O still never carries out of o[7].
template <class P, int I>
ZOLT_HD void sqr_row_next(uint32_t* e, uint32_t* o, const uint32_t* v, uint32_t ai) {
  e[0] = ptx::add_cc(e[0], o[1]);
  // O' = (previous E >> 64) + carry + odd positions of V_i * a_i
  if constexpr (1 >= I) { o[0] = ptx::madc_lo_cc(v[1], ai, o[2]); o[1] = ptx::madc_hi_cc(v[1], ai, o[3]); }
  else { o[0] = ptx::addc_cc(o[2], 0u); o[1] = ptx::addc_cc(o[3], 0u); }
  if constexpr (3 >= I) { o[2] = ptx::madc_lo_cc(v[3], ai, o[4]); o[3] = ptx::madc_hi_cc(v[3], ai, o[5]); }
  else { o[2] = ptx::addc_cc(o[4], 0u); o[3] = ptx::addc_cc(o[5], 0u); }
  if constexpr (5 >= I) { o[4] = ptx::madc_lo_cc(v[5], ai, o[6]); o[5] = ptx::madc_hi_cc(v[5], ai, o[7]); }
  else { o[4] = ptx::addc_cc(o[6], 0u); o[5] = ptx::addc_cc(o[7], 0u); }
  o[6] = ptx::madc_lo_cc(v[7], ai, o[8]);
  o[7] = ptx::madc_hi(v[7], ai, 0u);
  // E += even positions of V_i * a_i (first present pair opens the chain)
  constexpr int kFirstEven = (I + 1) / 2 * 2;  // smallest even k >= I
  if constexpr (kFirstEven <= 6) {
    if constexpr (0 >= I) {
      e[0] = ptx::mad_lo_cc(v[0], ai, e[0]);
      e[1] = ptx::madc_hi_cc(v[0], ai, e[1]);
    }
    if constexpr (2 >= I) {
      e[2] = (kFirstEven == 2) ? ptx::mad_lo_cc(v[2], ai, e[2]) : ptx::madc_lo_cc(v[2], ai, e[2]);
      e[3] = ptx::madc_hi_cc(v[2], ai, e[3]);
    }
    if constexpr (4 >= I) {
      e[4] = (kFirstEven == 4) ? ptx::mad_lo_cc(v[4], ai, e[4]) : ptx::madc_lo_cc(v[4], ai, e[4]);
      e[5] = ptx::madc_hi_cc(v[4], ai, e[5]);
    }
    e[6] = (kFirstEven == 6) ? ptx::mad_lo_cc(v[6], ai, e[6]) : ptx::madc_lo_cc(v[6], ai, e[6]);
    e[7] = ptx::madc_hi_cc(v[6], ai, e[7]);
    e[8] = ptx::addc(0u, 0u);
  } else {
    e[8] = 0u;
  }
  mont_row_reduce<P>(e, o);
}

// canonical a^2 (Montgomery); a canonical
template <class P>
ZOLT_HD Fp<P> fp_sqr(const Fp<P>& a) {
  // D = 2a (a < 2^254: no overflow)
  uint32_t d[8];
  d[0] = a.v[0] << 1;
#pragma unroll
  for (int k = 1; k < 8; k++) d[k] = (a.v[k] << 1) | (a.v[k - 1] >> 31);
  uint32_t x[9], y[9], v[8];
  // row 0: V_0 = (a_0, D[1] & ~1, D[2..7])
  v[0] = a.v[0];
  v[1] = d[1] & ~1u;
#pragma unroll
  for (int k = 2; k < 8; k++) v[k] = d[k];
  mont_row_first<P>(x, y, v, a.v[0]);
  // rows 1..7: V_i differs from V_(i-1) at positions i and i+1 only; positions below i are never read
#define ZOLT_SQR_ROW(E, O, I)                         \
  v[I] = a.v[I];                                      \
  if (I + 1 < 8) v[(I + 1) & 7] = d[(I + 1) & 7] & ~1u; \
  sqr_row_next<P, I>(E, O, v, a.v[I]);
  ZOLT_SQR_ROW(y, x, 1)
  ZOLT_SQR_ROW(x, y, 2)
  ZOLT_SQR_ROW(y, x, 3)
  ZOLT_SQR_ROW(x, y, 4)
  ZOLT_SQR_ROW(y, x, 5)
  ZOLT_SQR_ROW(x, y, 6)
  ZOLT_SQR_ROW(y, x, 7)
#undef ZOLT_SQR_ROW
  Fp<P> r;
  r.v[0] = ptx::add_cc(x[0], y[1]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = ptx::addc_cc(x[i], y[i + 1]);
  r.v[7] = ptx::addc(x[7], y[8]);
  fp_reduce_once<P>(r.v);
  return r;
}

// Montgomery -> standard integer limbs (reference fromMontgomery, src/field/mod.zig:642-645)
template <class P>
ZOLT_HD Fp<P> fp_from_mont(const Fp<P>& a) {
  Fp<P> one_raw = Fp<P>::zero();
  one_raw.v[0] = 1;
  return fp_mul<P>(a, one_raw);
}

// standard integer limbs (any 256-bit value) -> Montgomery (reference fromBytes/toMontgomery,
// src/field/mod.zig:625-652).  R2 < p is the row-scanned operand so the bound a < 2^255 holds.
template <class P>
ZOLT_HD Fp<P> fp_to_mont(const Fp<P>& raw) {
  return fp_mul<P>(Fp<P>::r2(), raw);
}

// small unsigned constant -> Montgomery form
template <class P>
ZOLT_HD Fp<P> fp_from_u32(uint32_t k) {
  Fp<P> raw = Fp<P>::zero();
  raw.v[0] = k;
  return fp_to_mont<P>(raw);
}

// a^(p-2) (reference inverse, src/field/mod.zig:955-983 / :500-518); 0 -> 0.
template <class P>
ZOLT_HD Fp<P> fp_inv(const Fp<P>& a) {
  // exponent p-2, MSB first square-and-multiply
  uint32_t e[8];
#pragma unroll
  for (int i = 0; i < 8; i++) e[i] = P::mod(i);
  e[0] -= 2u;  // MOD[0] >= 2 for both fields, no borrow
  Fp<P> acc = Fp<P>::one();
  for (int i = 7; i >= 0; i--) {
    for (int bit = 31; bit >= 0; bit--) {
      acc = fp_sqr<P>(acc);
      if ((e[i] >> bit) & 1u) acc = fp_mul<P>(acc, a);
    }
  }
  return acc;
}

// a^-1 by the binary extended Euclidean algorithm: same value as fp_inv (0 -> 0), about 4x less latency on a single
// thread (shifts and subtractions instead of 380 dependent Montgomery products).  Variable time and divergent, so it is
// used where ONE thread normalises ONE result (xyzz_to_affine at the end of an MSM), not in batch kernels.
template <class P>
ZOLT_HD Fp<P> fp_inv_vartime(const Fp<P>& a) {
  if (a.is_zero()) return a;
  uint32_t u[8], v[8], x1[8], x2[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    u[i] = a.v[i];
    v[i] = P::mod(i);
    x1[i] = i == 0 ? 1u : 0u;
    x2[i] = 0u;
  }
  auto is_one = [](const uint32_t* z) {
    uint32_t o = z[0] ^ 1u;
    for (int i = 1; i < 8; i++) o |= z[i];
    return o == 0;
  };
  auto shr1 = [](uint32_t* z) {
    for (int i = 0; i < 7; i++) z[i] = (z[i] >> 1) | (z[i + 1] << 31);
    z[7] >>= 1;
  };
  auto add_mod = [](uint32_t* z) {  // z += p  (z < p < 2^254: no overflow)
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
      c += (uint64_t)z[i] + P::mod(i);
      z[i] = (uint32_t)c;
      c >>= 32;
    }
  };
  auto geq = [](const uint32_t* x, const uint32_t* y) {
    for (int i = 7; i >= 0; i--) {
      if (x[i] != y[i]) return x[i] > y[i];
    }
    return true;
  };
  auto sub = [](uint32_t* x, const uint32_t* y) {  // x -= y, returns the borrow
    uint64_t b = 0;
    for (int i = 0; i < 8; i++) {
      const uint64_t d = (uint64_t)x[i] - y[i] - b;
      x[i] = (uint32_t)d;
      b = (d >> 32) & 1;
    }
    return (uint32_t)b;
  };
  auto halve = [&](uint32_t* z, uint32_t* x) {  // z even: z /= 2, x /= 2 mod p
    shr1(z);
    if (x[0] & 1u) add_mod(x);
    shr1(x);
  };
  while (!is_one(u) && !is_one(v)) {
    while (!(u[0] & 1u)) halve(u, x1);
    while (!(v[0] & 1u)) halve(v, x2);
    if (geq(u, v)) {
      sub(u, v);
      if (sub(x1, x2)) add_mod(x1);
    } else {
      sub(v, u);
      if (sub(x2, x1)) add_mod(x2);
    }
  }
  // y = (a * 2^256)^-1 as a plain integer; the Montgomery form of a^-1 is y * 2^512 = mont(mont(y, R^2), R^2)
  Fp<P> y;
  const uint32_t* res = is_one(u) ? x1 : x2;
#pragma unroll
  for (int i = 0; i < 8; i++) y.v[i] = res[i];
  return fp_mul<P>(fp_mul<P>(y, Fp<P>::r2()), Fp<P>::r2());
}

// ---------------------------------------------------------------------------------------------
// Delayed reduction: sums of products kept as unreduced 512-bit integers.
//
// A sum  S = sum_i a_i * b_i  of Montgomery-form products needs ONE Montgomery reduction at the end
// instead of one per term: each term then costs the 64 wide multiplies of the plain product instead
// of the 128 + 8 of a Montgomery product.  The result  S * 2^-256 mod p  is the same canonical field
// element as the sum of the individually reduced products, so outputs stay bit-identical to the
// reference's per-term montgomeryMul + add (src/field/mod.zig:735-798).
//
// Invariant of an accumulator acc[16]: acc < 2^511.  Every term is < 2^509.3 (operands < p and
// < 4p), so acc + term < 2^512 never carries out; wide_normalize then subtracts (2p) * 2^256
// (a multiple of p, 0.378 * 2^512) while bit 511 is set, which restores the invariant.
// ---------------------------------------------------------------------------------------------
// acc[s .. s+7] += (m0, m1, m2, m3) * bi laid out as four (lo, hi) pairs; the carry out goes to
// acc[s+8], which at that point only ever holds earlier carries (see wide_mul_add).
template <int S>
ZOLT_HD void wide_chain(uint32_t* acc, uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3, uint32_t bi) {
  acc[S + 0] = ptx::mad_lo_cc(m0, bi, acc[S + 0]);
  acc[S + 1] = ptx::madc_hi_cc(m0, bi, acc[S + 1]);
  acc[S + 2] = ptx::madc_lo_cc(m1, bi, acc[S + 2]);
  acc[S + 3] = ptx::madc_hi_cc(m1, bi, acc[S + 3]);
  acc[S + 4] = ptx::madc_lo_cc(m2, bi, acc[S + 4]);
  acc[S + 5] = ptx::madc_hi_cc(m2, bi, acc[S + 5]);
  acc[S + 6] = ptx::madc_lo_cc(m3, bi, acc[S + 6]);
  if constexpr (S + 8 < 16) {
    acc[S + 7] = ptx::madc_hi_cc(m3, bi, acc[S + 7]);
    acc[S + 8] = ptx::addc(acc[S + 8], 0u);
  } else {
    acc[S + 7] = ptx::madc_hi(m3, bi, acc[S + 7]);
  }
}

// prod[0..15] = a * b  (plain 256 x 256 -> 512-bit product, no reduction).
// The product is formed in two 16-limb temporaries: E collects the partial products a_j*b_i with
// i+j even (their low words sit on even limbs), O those with i+j odd, shifted down one limb, so that
// every (lo, hi) pair is one carry chain that ptxas turns into IMAD.WIDE.  Rows are visited in order,
// so the limb a chain carries into has so far only received carries and cannot ripple.
ZOLT_HD void wide_mul(uint32_t* e, const uint32_t* a, const uint32_t* b) {
  uint32_t o[16];
#pragma unroll
  for (int t = 0; t < 4; t++) {
    e[2 * t] = ptx::mul_lo(a[2 * t], b[0]);
    e[2 * t + 1] = ptx::mul_hi(a[2 * t], b[0]);
    o[2 * t] = ptx::mul_lo(a[2 * t + 1], b[0]);
    o[2 * t + 1] = ptx::mul_hi(a[2 * t + 1], b[0]);
  }
#pragma unroll
  for (int k = 8; k < 16; k++) e[k] = o[k] = 0;
  // odd rows i: E takes a1,a3,a5,a7 at limb i+1, O takes a0,a2,a4,a6 at limb i-1
  // even rows i: E takes a0,a2,a4,a6 at limb i,  O takes a1,a3,a5,a7 at limb i
  wide_chain<2>(e, a[1], a[3], a[5], a[7], b[1]);
  wide_chain<0>(o, a[0], a[2], a[4], a[6], b[1]);
  wide_chain<2>(e, a[0], a[2], a[4], a[6], b[2]);
  wide_chain<2>(o, a[1], a[3], a[5], a[7], b[2]);
  wide_chain<4>(e, a[1], a[3], a[5], a[7], b[3]);
  wide_chain<2>(o, a[0], a[2], a[4], a[6], b[3]);
  wide_chain<4>(e, a[0], a[2], a[4], a[6], b[4]);
  wide_chain<4>(o, a[1], a[3], a[5], a[7], b[4]);
  wide_chain<6>(e, a[1], a[3], a[5], a[7], b[5]);
  wide_chain<4>(o, a[0], a[2], a[4], a[6], b[5]);
  wide_chain<6>(e, a[0], a[2], a[4], a[6], b[6]);
  wide_chain<6>(o, a[1], a[3], a[5], a[7], b[6]);
  wide_chain<8>(e, a[1], a[3], a[5], a[7], b[7]);
  wide_chain<6>(o, a[0], a[2], a[4], a[6], b[7]);
  // prod = E + (O << 32)
  e[1] = ptx::add_cc(e[1], o[0]);
#pragma unroll
  for (int k = 2; k < 15; k++) e[k] = ptx::addc_cc(e[k], o[k - 1]);
  e[15] = ptx::addc(e[15], o[14]);
}

// acc += prod (no carry out: see the invariant above)
ZOLT_HD void wide_add(uint32_t* acc, const uint32_t* prod) {
  acc[0] = ptx::add_cc(acc[0], prod[0]);
#pragma unroll
  for (int k = 1; k < 15; k++) acc[k] = ptx::addc_cc(acc[k], prod[k]);
  acc[15] = ptx::addc(acc[15], prod[15]);
}

ZOLT_HD void wide_mul_add(uint32_t* acc, const uint32_t* a, const uint32_t* b) {
  uint32_t prod[16];
  wide_mul(prod, a, b);
  wide_add(acc, prod);
}

template <class P>
ZOLT_HD void wide_normalize(uint32_t* acc) {
  while (acc[15] & 0x80000000u) {
    acc[8] = ptx::sub_cc(acc[8], P::mod2(0));
#pragma unroll
    for (int i = 1; i < 7; i++) acc[8 + i] = ptx::subc_cc(acc[8 + i], P::mod2(i));
    acc[15] = ptx::subc(acc[15], P::mod2(7));
  }
}

ZOLT_HD void wide_zero(uint32_t* acc) {
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = 0;
}

// canonical  acc * 2^-256 mod p  for acc < 2^511 (runs once per thread per kernel: written for
// clarity, not speed).  (acc + M*p) / 2^256 < 2^255 + p < 4p, hence three conditional subtractions.
template <class P>
ZOLT_HD Fp<P> wide_reduce(const uint32_t* acc_in) {
  uint32_t t[17];
#pragma unroll
  for (int i = 0; i < 16; i++) t[i] = acc_in[i];
  t[16] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t m = t[i] * P::INV;
    uint64_t carry = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const uint64_t x = (uint64_t)m * P::mod(j) + t[i + j] + carry;
      t[i + j] = (uint32_t)x;
      carry = x >> 32;
    }
#pragma unroll
    for (int k = i + 8; k < 17; k++) {
      const uint64_t x = (uint64_t)t[k] + carry;
      t[k] = (uint32_t)x;
      carry = x >> 32;
    }
  }
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = t[8 + i];
  fp_reduce_once<P>(r.v);
  fp_reduce_once<P>(r.v);
  fp_reduce_once<P>(r.v);
  return r;
}

// ---------------------------------------------------------------------------------------------
// Multiplication by a per-launch constant (the sumcheck bind: every element of every table is
// multiplied by the same challenge r).
//
// With c[i] = (r_std * 2^(32 i)) mod p precomputed once per launch (r_std = the plain integer the
// Montgomery limbs of r stand for),  sum_i d_i * c[i]  is congruent to r_std * d for ANY 256-bit d, so for
// d = Montgomery limbs of a field element it is the Montgomery form of the product -- the value
// montgomeryMul (src/field/mod.zig:735-779) returns -- reached with the 64 wide multiplies of a plain
// 8 x 8 product and NO interleaved reduction rows.  The sum (plus an addend < p) is below 2^289 and is
// brought back to [0, p) by one Barrett step on its top 65 bits (quotient estimate short by at most 1) and
// one conditional subtraction: 128 + 28 + ~12 integer multiply-adds instead of 264.
// Table rows are kernel parameters, i.e. constant-bank operands of the multiply instructions.
// ---------------------------------------------------------------------------------------------
struct FixedMulTable {
  uint32_t c[8][8];
};

// c[i] = mont_mul(r, 2^(32 i) as a raw integer) = r_std * 2^(32 i) mod p
template <class P>
ZOLT_HD FixedMulTable fixed_mul_table_barrett(const Fp<P>& r) {
  FixedMulTable t;
  for (int i = 0; i < 8; i++) {
    Fp<P> e = Fp<P>::zero();
    e.v[i] = 1u;
    const Fp<P> ci = fp_mul<P>(r, e);
    for (int j = 0; j < 8; j++) t.c[i][j] = ci.v[j];
  }
  return t;
}

// acc[0..7] += (m0, m1, m2, m3) * d as four (lo, hi) pairs, carry into acc[8]
ZOLT_HD void fixed_chain(uint32_t* acc, uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3, uint32_t d) {
  acc[0] = ptx::mad_lo_cc(m0, d, acc[0]);
  acc[1] = ptx::madc_hi_cc(m0, d, acc[1]);
  acc[2] = ptx::madc_lo_cc(m1, d, acc[2]);
  acc[3] = ptx::madc_hi_cc(m1, d, acc[3]);
  acc[4] = ptx::madc_lo_cc(m2, d, acc[4]);
  acc[5] = ptx::madc_hi_cc(m2, d, acc[5]);
  acc[6] = ptx::madc_lo_cc(m3, d, acc[6]);
  acc[7] = ptx::madc_hi_cc(m3, d, acc[7]);
  acc[8] = ptx::addc(acc[8], 0u);
}

// x[0..9] (< 2^289, the value of a table product plus an addend) -> the canonical representative of x mod p.
template <class P>
ZOLT_HD Fp<P> fp_barrett_289(const uint32_t* x) {
  // Barrett: q = floor(h * mu / 2^96), h = floor(x / 2^224) = x[9] 2^64 + H, mu = floor(2^320 / p) = MU_HI 2^64 + MU_LO.
  // Both truncations together cost less than 2^-29 of a unit, so q is floor(x / p) or one less.  The low 64 bits of
  // H * MU_LO cannot change floor(. / 2^96):  q = x[9] MU_HI 2^32 + floor((x[9] MU_LO + H MU_HI + hi64(H MU_LO)) / 2^32).
  const unsigned long long H = ((unsigned long long)x[8] << 32) | x[7];
#if defined(__CUDA_ARCH__)
  const unsigned long long hi64 = __umul64hi(H, P::MU320_LO);
#else
  const unsigned long long hi64 = (unsigned long long)(((unsigned __int128)H * P::MU320_LO) >> 64);
#endif
  const unsigned __int128 bsum = (unsigned __int128)H * P::MU320_HI + hi64 + (x[9] ? (unsigned __int128)x[9] * P::MU320_LO : 0);
  const unsigned long long q = (unsigned long long)(bsum >> 32) + (((unsigned long long)x[9] * P::MU320_HI) << 32);
  const uint32_t q0 = (uint32_t)q, q1 = (uint32_t)(q >> 32);
  // q * p mod 2^256 = TE + (TO << 32)
  uint32_t te[8], to[7];
  te[0] = ptx::mul_lo(P::mod(0), q0);
  te[1] = ptx::mul_hi(P::mod(0), q0);
  te[2] = ptx::mul_lo(P::mod(2), q0);
  te[3] = ptx::mul_hi(P::mod(2), q0);
  te[4] = ptx::mul_lo(P::mod(4), q0);
  te[5] = ptx::mul_hi(P::mod(4), q0);
  te[6] = ptx::mul_lo(P::mod(6), q0);
  te[7] = ptx::mul_hi(P::mod(6), q0);
  to[0] = ptx::mul_lo(P::mod(1), q0);
  to[1] = ptx::mul_hi(P::mod(1), q0);
  to[2] = ptx::mul_lo(P::mod(3), q0);
  to[3] = ptx::mul_hi(P::mod(3), q0);
  to[4] = ptx::mul_lo(P::mod(5), q0);
  to[5] = ptx::mul_hi(P::mod(5), q0);
  to[6] = ptx::mul_lo(P::mod(7), q0);
  // + q1 * p * 2^32: even limbs of p land on the TO alignment, odd limbs on TE two limbs up
  to[0] = ptx::mad_lo_cc(P::mod(0), q1, to[0]);
  to[1] = ptx::madc_hi_cc(P::mod(0), q1, to[1]);
  to[2] = ptx::madc_lo_cc(P::mod(2), q1, to[2]);
  to[3] = ptx::madc_hi_cc(P::mod(2), q1, to[3]);
  to[4] = ptx::madc_lo_cc(P::mod(4), q1, to[4]);
  to[5] = ptx::madc_hi_cc(P::mod(4), q1, to[5]);
  to[6] = ptx::madc_lo_cc(P::mod(6), q1, to[6]);
  te[2] = ptx::mad_lo_cc(P::mod(1), q1, te[2]);
  te[3] = ptx::madc_hi_cc(P::mod(1), q1, te[3]);
  te[4] = ptx::madc_lo_cc(P::mod(3), q1, te[4]);
  te[5] = ptx::madc_hi_cc(P::mod(3), q1, te[5]);
  te[6] = ptx::madc_lo_cc(P::mod(5), q1, te[6]);
  te[7] = ptx::madc_hi(P::mod(5), q1, te[7]);
  // rem = x - q p  (true value in [0, 2p) < 2^256: the low 8 limbs are exact)
  Fp<P> r;
  r.v[0] = ptx::sub_cc(x[0], te[0]);
#pragma unroll
  for (int j = 1; j < 7; j++) r.v[j] = ptx::subc_cc(x[j], te[j]);
  r.v[7] = ptx::subc(x[7], te[7]);
  r.v[1] = ptx::sub_cc(r.v[1], to[0]);
#pragma unroll
  for (int j = 2; j < 7; j++) r.v[j] = ptx::subc_cc(r.v[j], to[j - 1]);
  r.v[7] = ptx::subc(r.v[7], to[6]);
  fp_reduce_once<P>(r.v);  // [0, 2p) -> [0, p)
  return r;
}

// canonical (addend + r * d) for the table of r;  addend < p canonical, d any 256-bit representative.
template <class P>
ZOLT_HD Fp<P> fp_fixed_mul_add_barrett(const Fp<P>& addend, const uint32_t* d, const FixedMulTable& t) {
  // x = addend + sum_i d_i * c[i] = E + (O << 32): E takes the even limbs of c[i], O the odd ones
  uint32_t e[9], o[9];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    e[j] = addend.v[j];
    o[j] = 0;
  }
  e[8] = o[8] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    fixed_chain(e, t.c[i][0], t.c[i][2], t.c[i][4], t.c[i][6], d[i]);
    fixed_chain(o, t.c[i][1], t.c[i][3], t.c[i][5], t.c[i][7], d[i]);
  }
  uint32_t x[10];
  x[0] = e[0];
  x[1] = ptx::add_cc(e[1], o[0]);
#pragma unroll
  for (int j = 2; j < 9; j++) x[j] = ptx::addc_cc(e[j], o[j - 1]);
  x[9] = ptx::addc(o[8], 0u);
  return fp_barrett_289<P>(x);
}

// ---------------------------------------------------------------------------------------------
// The same product with the table scaled by 2^64 and a two-limb Montgomery reduction instead of the Barrett step (r2 v23).
//   c64[i] = (r_std * 2^(32 i) * 2^64) mod p
//   x = (addend << 64) + sum_i d_i * c64[i]          < 2^64 p + 8 * 2^32 p
//   two reduction rows (m0 = x_0 * INV, m1 = x_1 * INV) clear the two low limbs:  (x + m0 p + m1 p 2^32) / 2^64
//   is congruent to addend + r_std * d and below 2p + p / 2^29: two conditional subtractions give the canonical value --
//   the field element montgomeryMul / the Barrett form return.  The addend enters shifted by two limbs, so it costs no
//   multiplication.  64 + 16 wide multiplies and 2 narrow ones, against 64 + 15 (q p) + ~8 (the quotient estimate: a 64 x 64
//   high product and its corrections) of the Barrett form, and no 64-bit quotient arithmetic on the ALU.
// ---------------------------------------------------------------------------------------------
template <class P>
ZOLT_HD FixedMulTable fixed_mul_table_mont2(const Fp<P>& r) {
  FixedMulTable t;
  Fp<P> e = Fp<P>::zero();
  e.v[2] = 1u;
  const Fp<P> base_std = fp_mul<P>(r, e);                 // r_std * 2^64 mod p, as a plain integer
  const Fp<P> base = fp_mul<P>(base_std, Fp<P>::r2());     // its Montgomery form
  for (int i = 0; i < 8; i++) {
    Fp<P> s = Fp<P>::zero();
    s.v[i] = 1u;
    const Fp<P> ci = fp_mul<P>(base, s);                   // r_std * 2^64 * 2^(32 i) mod p
    for (int j = 0; j < 8; j++) t.c[i][j] = ci.v[j];
  }
  return t;
}

template <class P>
ZOLT_HD Fp<P> fp_fixed_mul_add_mont2(const Fp<P>& addend, const uint32_t* d, const FixedMulTable& t) {
  // sum_i d_i * c64[i] = E + (O << 32): E takes the even limbs of the rows, O the odd ones
  uint32_t e[11], o[9];
#pragma unroll
  for (int j = 0; j < 11; j++) e[j] = 0;
#pragma unroll
  for (int j = 0; j < 9; j++) o[j] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    fixed_chain(e, t.c[i][0], t.c[i][2], t.c[i][4], t.c[i][6], d[i]);
    fixed_chain(o, t.c[i][1], t.c[i][3], t.c[i][5], t.c[i][7], d[i]);
  }
  // row m0: limb 0 of x is e[0] (the addend starts at limb 2, O at limb 1)
  const uint32_t m0 = ptx::mul_lo(e[0], P::INV);
  fixed_chain(e, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m0);
  fixed_chain(o, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m0);
  // row m1 (shifted by one limb): limb 1 of x is e[1] + o[0]; the even limbs of p land on the O alignment, the odd ones on
  // E two limbs up
  const uint32_t m1 = ptx::mul_lo(e[1] + o[0], P::INV);
  fixed_chain(o, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m1);
  fixed_chain(e + 2, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m1);
  // (E + (O << 32) + (addend << 64)) >> 64: limb 0 is zero, limb 1 sums to zero and leaves its carry
  uint32_t x[8];
  (void)ptx::add_cc(e[1], o[0]);
#pragma unroll
  for (int j = 0; j < 7; j++) x[j] = ptx::addc_cc(e[j + 2], o[j + 1]);
  x[7] = ptx::addc(e[9], o[8]);
  Fp<P> r;
  r.v[0] = ptx::add_cc(x[0], addend.v[0]);
#pragma unroll
  for (int j = 1; j < 7; j++) r.v[j] = ptx::addc_cc(x[j], addend.v[j]);
  r.v[7] = ptx::addc(x[7], addend.v[7]);
  fp_reduce_once<P>(r.v);  // [0, 2p + p / 2^29) -> [0, p + p / 2^29)
  fp_reduce_once<P>(r.v);  // -> [0, p)
  return r;
}

// ---------------------------------------------------------------------------------------------
// One reduction row (r2 v24).  Table scaled by 2^32: x = (addend << 32) + sum_i d_i * c32[i] < 9 * 2^32 p; one row
// (m0 = x_0 * INV) clears the low limb and y = (x + m0 p) / 2^32 < 10 p is brought to [0, p) by a subtraction ladder
// (8p, 4p, 2p, p: compare-and-select on the ALU pipe, which has the headroom the multiplier lacks).  64 + 8 wide multiplies
// and 1 narrow one.
// ---------------------------------------------------------------------------------------------
// limb i (0..8) of (p << sh), sh < 32
template <class P>
ZOLT_HD constexpr uint32_t mod_shl(int sh, int i) {
  const uint32_t lo = i < 8 ? P::mod(i) : 0u;
  const uint32_t below = i > 0 ? P::mod(i - 1) : 0u;
  return sh == 0 ? lo : ((lo << sh) | (below >> (32 - sh)));
}
// y (9 limbs) -= (p << SH) if that leaves it non-negative
template <class P, int SH>
ZOLT_HD void fp_ladder_step(uint32_t* y) {
  uint32_t t[9];
  t[0] = ptx::sub_cc(y[0], mod_shl<P>(SH, 0));
#pragma unroll
  for (int i = 1; i < 9; i++) t[i] = ptx::subc_cc(y[i], mod_shl<P>(SH, i));
  const uint32_t borrow = ptx::subc(0u, 0u);  // 0xffffffff if y < (p << SH)
#pragma unroll
  for (int i = 0; i < 9; i++) y[i] = borrow ? y[i] : t[i];
}

template <class P>
ZOLT_HD FixedMulTable fixed_mul_table_mont1(const Fp<P>& r) {
  FixedMulTable t;
  Fp<P> e = Fp<P>::zero();
  e.v[1] = 1u;
  const Fp<P> base_std = fp_mul<P>(r, e);                 // r_std * 2^32 mod p, as a plain integer
  const Fp<P> base = fp_mul<P>(base_std, Fp<P>::r2());     // its Montgomery form
  for (int i = 0; i < 8; i++) {
    Fp<P> s = Fp<P>::zero();
    s.v[i] = 1u;
    const Fp<P> ci = fp_mul<P>(base, s);                   // r_std * 2^32 * 2^(32 i) mod p
    for (int j = 0; j < 8; j++) t.c[i][j] = ci.v[j];
  }
  return t;
}

template <class P>
ZOLT_HD Fp<P> fp_fixed_mul_add_mont1(const Fp<P>& addend, const uint32_t* d, const FixedMulTable& t) {
  uint32_t e[9], o[9];
#pragma unroll
  for (int j = 0; j < 9; j++) e[j] = o[j] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    fixed_chain(e, t.c[i][0], t.c[i][2], t.c[i][4], t.c[i][6], d[i]);
    fixed_chain(o, t.c[i][1], t.c[i][3], t.c[i][5], t.c[i][7], d[i]);
  }
  const uint32_t m0 = ptx::mul_lo(e[0], P::INV);  // limb 0 of x is e[0]: O and the addend start one limb up
  fixed_chain(e, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m0);
  fixed_chain(o, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m0);
  // y = (E + ((O + addend) << 32)) >> 32; e[0] is zero now
  uint32_t y[9];
  y[0] = ptx::add_cc(e[1], o[0]);
#pragma unroll
  for (int j = 1; j < 8; j++) y[j] = ptx::addc_cc(e[j + 1], o[j]);
  y[8] = ptx::addc(o[8], 0u);
  y[0] = ptx::add_cc(y[0], addend.v[0]);
#pragma unroll
  for (int j = 1; j < 8; j++) y[j] = ptx::addc_cc(y[j], addend.v[j]);
  y[8] = ptx::addc(y[8], 0u);
  fp_ladder_step<P, 3>(y);  // [0, 10p) -> [0, 8p)
  fp_ladder_step<P, 2>(y);  // -> [0, 4p)
  fp_ladder_step<P, 1>(y);  // -> [0, 2p)
  fp_ladder_step<P, 0>(y);  // -> [0, p)
  Fp<P> r;
#pragma unroll
  for (int j = 0; j < 8; j++) r.v[j] = y[j];
  return r;
}

// The form in use.  ZOLT_BIND_BARRETT selects the r2 v7 Barrett reduction (kept for A/B timing; same values).
#if defined(ZOLT_BIND_BARRETT)
template <class P>
ZOLT_HD FixedMulTable fixed_mul_table(const Fp<P>& r) { return fixed_mul_table_barrett<P>(r); }
template <class P>
ZOLT_HD Fp<P> fp_fixed_mul_add(const Fp<P>& addend, const uint32_t* d, const FixedMulTable& t) { return fp_fixed_mul_add_barrett<P>(addend, d, t); }
#elif defined(ZOLT_BIND_MONT1)
template <class P>
ZOLT_HD FixedMulTable fixed_mul_table(const Fp<P>& r) { return fixed_mul_table_mont1<P>(r); }
template <class P>
ZOLT_HD Fp<P> fp_fixed_mul_add(const Fp<P>& addend, const uint32_t* d, const FixedMulTable& t) { return fp_fixed_mul_add_mont1<P>(addend, d, t); }
#else
template <class P>
ZOLT_HD FixedMulTable fixed_mul_table(const Fp<P>& r) { return fixed_mul_table_mont2<P>(r); }
template <class P>
ZOLT_HD Fp<P> fp_fixed_mul_add(const Fp<P>& addend, const uint32_t* d, const FixedMulTable& t) { return fp_fixed_mul_add_mont2<P>(addend, d, t); }
#endif

// lo + r * (hi - lo)  (DensePolynomial.bindLow, src/poly/mod.zig:171) through the table of r.  The difference
// is taken as hi + (p - lo), a non-canonical representative below 2p: the table product accepts any limbs.
template <class P>
ZOLT_HD Fp<P> fp_bind_fixed(const Fp<P>& lo, const Fp<P>& hi, const FixedMulTable& t) {
  uint32_t d[8];
  d[0] = ptx::sub_cc(P::mod(0), lo.v[0]);
#pragma unroll
  for (int j = 1; j < 7; j++) d[j] = ptx::subc_cc(P::mod(j), lo.v[j]);
  d[7] = ptx::subc(P::mod(7), lo.v[7]);
  d[0] = ptx::add_cc(d[0], hi.v[0]);
#pragma unroll
  for (int j = 1; j < 7; j++) d[j] = ptx::addc_cc(d[j], hi.v[j]);
  d[7] = ptx::addc(d[7], hi.v[7]);
  return fp_fixed_mul_add<P>(lo, d, t);
}

// ---------------------------------------------------------------------------------------------
// The same fixed-multiplier product on the FP64 pipe.
//
// On B200 the FP64 pipe (DFMA, 16 lanes / SM sub-partition / clock, the rate of a 32-bit IMAD) sits idle beside
// the integer multiplier that bounds the sumcheck passes, and it issues in parallel with it.  The bind's product has a
// per-launch constant on one side, so it maps onto exact double-precision arithmetic with no carry handling at all:
//   d  = 16 limbs of 16 bits  (d_l < 2^16)
//   C_l = (r_std * 2^(16 l)) mod p  as 8 limbs of 32 bits, converted to doubles ON THE HOST (kernel parameter, i.e.
//         uniform-register operands of the DFMAs)
//   col_j = sum_l d_l * C_l[j]   -- 16 terms below 2^48: every partial sum is an integer below 2^52, exact in a double
// so  lo + sum_j col_j 2^(32 j)  is congruent to  lo + r_std * d  and below 2^277; eight (low 32, high 20) splits,
// two carry chains and a one-word Barrett step (quotient < 2^23) bring it to the canonical value -- the same field
// element fp_fixed_mul_add / montgomeryMul produce.  128 DFMA + 24 DADD + ~50 integer instructions instead of 64 wide
// integer multiplies (256 multiplier-pipe cycles per warp) + 15 more in the two-word Barrett step.
// ---------------------------------------------------------------------------------------------
struct alignas(16) F64Pair {
  double x, y;
};
struct alignas(16) FixedMulTableF64 {
  F64Pair c[16][4];  // row l: limbs (0,1) (2,3) (4,5) (6,7) of C_l; pairs so that one 128-bit constant load feeds two DFMAs
};

// table of r for fp_fixed_mul_add_f64: row l = limbs of r_std * 2^(16 l) mod p  (O(1) per launch; host or one thread)
template <class P>
ZOLT_HD FixedMulTableF64 fixed_mul_table_f64(const Fp<P>& r) {
  FixedMulTableF64 t;
  // r_std * 2^(16 l) = mont_mul(r, 2^(16 l) as a raw integer) for l < 16 (2^(16 l) < 2^256)
  for (int l = 0; l < 16; l++) {
    Fp<P> e = Fp<P>::zero();
    e.v[l >> 1] = (l & 1) ? 0x10000u : 1u;
    const Fp<P> cl = fp_mul<P>(r, e);
    for (int j = 0; j < 4; j++) t.c[l][j] = F64Pair{(double)cl.v[2 * j], (double)cl.v[2 * j + 1]};
  }
  return t;
}

namespace f64 {
#if defined(__CUDA_ARCH__)
// exact u32 -> double without the conversion unit: 2^52 + x by bit pattern, minus 2^52
ZOLT_D double from_u32(uint32_t x) { return __hiloint2double(0x43300000, (int)x) - 4503599627370496.0; }
// integer-valued v in [0, 2^52) -> (low 32 bits, high 20 bits)
ZOLT_D void split52(double v, uint32_t& lo, uint32_t& hi) {
  const double y = v + 4503599627370496.0;
  lo = (uint32_t)__double2loint(y);
  hi = (uint32_t)__double2hiint(y) & 0xfffffu;
}
ZOLT_D double fma_(double a, double b, double c) { return fma(a, b, c); }
#else
inline double from_u32(uint32_t x) { return (double)x; }
inline void split52(double v, uint32_t& lo, uint32_t& hi) {
  const unsigned long long y = (unsigned long long)v;
  lo = (uint32_t)y;
  hi = (uint32_t)(y >> 32);
}
inline double fma_(double a, double b, double c) { return __builtin_fma(a, b, c); }
#endif
}  // namespace f64

// x[0..8] < 2^277 with floor(x / p) < 2^32 -> canonical x mod p (one-word Barrett: see fp_barrett_289 for the error bound)
template <class P>
ZOLT_HD Fp<P> fp_barrett_small(const uint32_t* x) {
  const unsigned long long H = ((unsigned long long)x[8] << 32) | x[7];  // floor(x / 2^224) < 2^53
#if defined(__CUDA_ARCH__)
  const unsigned long long hi64 = __umul64hi(H, P::MU320_LO);
#else
  const unsigned long long hi64 = (unsigned long long)(((unsigned __int128)H * P::MU320_LO) >> 64);
#endif
  const uint32_t q = (uint32_t)((H * P::MU320_HI + hi64) >> 32);  // floor(x / p) or one less
  // q * p mod 2^256 = TE + (TO << 32)
  uint32_t te[8], to[7];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    te[2 * k] = ptx::mul_lo(P::mod(2 * k), q);
    te[2 * k + 1] = ptx::mul_hi(P::mod(2 * k), q);
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    to[2 * k] = ptx::mul_lo(P::mod(2 * k + 1), q);
    to[2 * k + 1] = ptx::mul_hi(P::mod(2 * k + 1), q);
  }
  to[6] = ptx::mul_lo(P::mod(7), q);
  Fp<P> r;
  r.v[0] = ptx::sub_cc(x[0], te[0]);
#pragma unroll
  for (int j = 1; j < 7; j++) r.v[j] = ptx::subc_cc(x[j], te[j]);
  r.v[7] = ptx::subc(x[7], te[7]);
  r.v[1] = ptx::sub_cc(r.v[1], to[0]);
#pragma unroll
  for (int j = 2; j < 7; j++) r.v[j] = ptx::subc_cc(r.v[j], to[j - 1]);
  r.v[7] = ptx::subc(r.v[7], to[6]);
  fp_reduce_once<P>(r.v);  // [0, 2p) -> [0, p)
  return r;
}

// canonical (addend + r * d) through the FP64 table of r;  addend < p canonical, d any 256-bit value
template <class P>
ZOLT_HD Fp<P> fp_fixed_mul_add_f64(const Fp<P>& addend, const uint32_t* d, const FixedMulTableF64& t) {
  double col[8];
#pragma unroll
  for (int j = 0; j < 8; j++) col[j] = 0.0;
#pragma unroll
  for (int l = 0; l < 16; l++) {
    const uint32_t w = d[l >> 1];
    const double x = f64::from_u32((l & 1) ? (w >> 16) : (w & 0xffffu));
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const F64Pair c = t.c[l][j];
      col[2 * j] = f64::fma_(x, c.x, col[2 * j]);
      col[2 * j + 1] = f64::fma_(x, c.y, col[2 * j + 1]);
    }
  }
  uint32_t lo[8], hi[8];
#pragma unroll
  for (int j = 0; j < 8; j++) f64::split52(col[j], lo[j], hi[j]);
  // x = addend + sum_j lo[j] 2^(32 j) + sum_j hi[j] 2^(32 (j + 1))
  uint32_t x[9];
  x[0] = ptx::add_cc(addend.v[0], lo[0]);
#pragma unroll
  for (int j = 1; j < 8; j++) x[j] = ptx::addc_cc(addend.v[j], lo[j]);
  x[8] = ptx::addc(0u, 0u);
  x[1] = ptx::add_cc(x[1], hi[0]);
#pragma unroll
  for (int j = 2; j < 8; j++) x[j] = ptx::addc_cc(x[j], hi[j - 1]);
  x[8] = ptx::addc(x[8], hi[7]);
  return fp_barrett_small<P>(x);
}

// lo + r * (hi - lo) through the FP64 table of r (see fp_bind_fixed)
template <class P>
ZOLT_HD Fp<P> fp_bind_fixed_f64(const Fp<P>& lo, const Fp<P>& hi, const FixedMulTableF64& t) {
  uint32_t d[8];
  d[0] = ptx::sub_cc(P::mod(0), lo.v[0]);
#pragma unroll
  for (int j = 1; j < 7; j++) d[j] = ptx::subc_cc(P::mod(j), lo.v[j]);
  d[7] = ptx::subc(P::mod(7), lo.v[7]);
  d[0] = ptx::add_cc(d[0], hi.v[0]);
#pragma unroll
  for (int j = 1; j < 7; j++) d[j] = ptx::addc_cc(d[j], hi.v[j]);
  d[7] = ptx::addc(d[7], hi.v[7]);
  return fp_fixed_mul_add_f64<P>(lo, d, t);
}

// Two binds against the same table, interleaved: 16 independent DFMA chains instead of 8 and every table constant
// loaded once for two products.  (lo0, hi0) -> out0, (lo1, hi1) -> out1; same values as two fp_bind_fixed_f64 calls.
template <class P>
ZOLT_HD void fp_bind2_fixed_f64(const Fp<P>& lo0, const Fp<P>& hi0, const Fp<P>& lo1, const Fp<P>& hi1,
                                const FixedMulTableF64& t, Fp<P>& out0, Fp<P>& out1) {
  uint32_t d0[8], d1[8];
  d0[0] = ptx::sub_cc(P::mod(0), lo0.v[0]);
#pragma unroll
  for (int j = 1; j < 7; j++) d0[j] = ptx::subc_cc(P::mod(j), lo0.v[j]);
  d0[7] = ptx::subc(P::mod(7), lo0.v[7]);
  d0[0] = ptx::add_cc(d0[0], hi0.v[0]);
#pragma unroll
  for (int j = 1; j < 7; j++) d0[j] = ptx::addc_cc(d0[j], hi0.v[j]);
  d0[7] = ptx::addc(d0[7], hi0.v[7]);
  d1[0] = ptx::sub_cc(P::mod(0), lo1.v[0]);
#pragma unroll
  for (int j = 1; j < 7; j++) d1[j] = ptx::subc_cc(P::mod(j), lo1.v[j]);
  d1[7] = ptx::subc(P::mod(7), lo1.v[7]);
  d1[0] = ptx::add_cc(d1[0], hi1.v[0]);
#pragma unroll
  for (int j = 1; j < 7; j++) d1[j] = ptx::addc_cc(d1[j], hi1.v[j]);
  d1[7] = ptx::addc(d1[7], hi1.v[7]);
  double c0[8], c1[8];
#pragma unroll
  for (int j = 0; j < 8; j++) c0[j] = c1[j] = 0.0;
#pragma unroll
  for (int l = 0; l < 16; l++) {
    const uint32_t w0 = d0[l >> 1], w1 = d1[l >> 1];
    const double x0 = f64::from_u32((l & 1) ? (w0 >> 16) : (w0 & 0xffffu));
    const double x1 = f64::from_u32((l & 1) ? (w1 >> 16) : (w1 & 0xffffu));
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const F64Pair c = t.c[l][j];
      c0[2 * j] = f64::fma_(x0, c.x, c0[2 * j]);
      c1[2 * j] = f64::fma_(x1, c.x, c1[2 * j]);
      c0[2 * j + 1] = f64::fma_(x0, c.y, c0[2 * j + 1]);
      c1[2 * j + 1] = f64::fma_(x1, c.y, c1[2 * j + 1]);
    }
  }
  auto finish = [](const Fp<P>& addend, const double* col) {
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int j = 0; j < 8; j++) f64::split52(col[j], lo[j], hi[j]);
    uint32_t x[9];
    x[0] = ptx::add_cc(addend.v[0], lo[0]);
#pragma unroll
    for (int j = 1; j < 8; j++) x[j] = ptx::addc_cc(addend.v[j], lo[j]);
    x[8] = ptx::addc(0u, 0u);
    x[1] = ptx::add_cc(x[1], hi[0]);
#pragma unroll
    for (int j = 2; j < 8; j++) x[j] = ptx::addc_cc(x[j], hi[j - 1]);
    x[8] = ptx::addc(x[8], hi[7]);
    return fp_barrett_small<P>(x);
  };
  out0 = finish(lo0, c0);
  out1 = finish(lo1, c1);
}

using Fr = Fp<FrParams>;
using Fq = Fp<FqParams>;

}  // namespace zolt
