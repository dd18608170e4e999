// Host-side BN254 scalar for the O(1)-per-round glue that stays on the CPU in the reference too:
// claim updates, UniPoly interpolation, the Gruen cubic assembly and Fiat-Shamir challenges
// (reference src/field/mod.zig:583-1052; src/poly/mod.zig:584-687; src/poly/split_eq.zig:353-434).
// The O(N) loops never run here -- they are CUDA kernels behind include/zolt_gpu.h.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>

namespace zolt::host {

class Fr {
 public:
  std::array<uint64_t, 4> limbs{};  // Montgomery form, canonical, little-endian limbs (== BN254Scalar.limbs)

  static constexpr std::array<uint64_t, 4> kModulus = {0x43e1f593f0000001ull, 0x2833e84879b97091ull,
                                                       0xb85045b68181585dull, 0x30644e72e131a029ull};
  static constexpr std::array<uint64_t, 4> kR = {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull,
                                                 0x666ea36f7879462eull, 0x0e0a77c19a07df2full};
  static constexpr std::array<uint64_t, 4> kR2 = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull,
                                                  0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
  static constexpr uint64_t kInv = 0xc2e1f593efffffffull;  // -r^-1 mod 2^64

  static Fr zero() { return Fr{}; }
  static Fr one() { return fromLimbs(kR); }
  static Fr fromLimbs(const std::array<uint64_t, 4>& l) {
    Fr r;
    r.limbs = l;
    return r;
  }
  static Fr fromRaw(const uint64_t* p) {
    Fr r;
    std::memcpy(r.limbs.data(), p, 32);
    return r;
  }
  void store(uint64_t* p) const { std::memcpy(p, limbs.data(), 32); }

  static Fr fromU64(uint64_t v) { return montMul(Fr::fromLimbs({v, 0, 0, 0}), Fr::fromLimbs(kR2)); }
  // fromBytes: little-endian 256-bit value, reduced, to Montgomery form (src/field/mod.zig:625-639)
  static Fr fromBytes(const uint8_t* bytes32) {
    Fr raw;
    std::memcpy(raw.limbs.data(), bytes32, 32);
    return montMul(Fr::fromLimbs(kR2), raw);
  }
  // toBytes: canonical little-endian (src/field/mod.zig:685-695)
  void toBytes(uint8_t* out32) const {
    Fr s = montMul(*this, Fr::fromLimbs({1, 0, 0, 0}));
    std::memcpy(out32, s.limbs.data(), 32);
  }

  bool isZero() const { return (limbs[0] | limbs[1] | limbs[2] | limbs[3]) == 0; }
  bool eql(const Fr& o) const { return limbs == o.limbs; }

  Fr add(const Fr& o) const {
    Fr r;
    unsigned __int128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (unsigned __int128)limbs[i] + o.limbs[i];
      r.limbs[i] = (uint64_t)c;
      c >>= 64;
    }
    if (c || !r.ltModulus()) r.subModulus();
    return r;
  }
  Fr sub(const Fr& o) const {
    Fr r;
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
      unsigned __int128 d = (unsigned __int128)limbs[i] - o.limbs[i] - borrow;
      r.limbs[i] = (uint64_t)d;
      borrow = (uint64_t)(d >> 64) & 1;
    }
    if (borrow) {
      unsigned __int128 c = 0;
      for (int i = 0; i < 4; i++) {
        c += (unsigned __int128)r.limbs[i] + kModulus[i];
        r.limbs[i] = (uint64_t)c;
        c >>= 64;
      }
    }
    return r;
  }
  Fr neg() const { return isZero() ? *this : Fr::fromLimbs(kModulus).sub(*this); }
  Fr mul(const Fr& o) const { return montMul(*this, o); }
  Fr square() const { return montMul(*this, *this); }
  Fr dbl() const { return add(*this); }
  // a^(r-2); the reference returns null for zero (src/field/mod.zig:955-983) -- callers test isZero
  Fr inverse() const {
    std::array<uint64_t, 4> e = kModulus;
    e[0] -= 2;
    Fr result = one(), base = *this;
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 64; j++) {
        if ((e[i] >> j) & 1) result = result.mul(base);
        base = base.square();
      }
    return result;
  }

 private:
  bool ltModulus() const {
    for (int i = 3; i >= 0; i--) {
      if (limbs[i] < kModulus[i]) return true;
      if (limbs[i] > kModulus[i]) return false;
    }
    return false;
  }
  void subModulus() {
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
      unsigned __int128 d = (unsigned __int128)limbs[i] - kModulus[i] - borrow;
      limbs[i] = (uint64_t)d;
      borrow = (uint64_t)(d >> 64) & 1;
    }
  }
  // Montgomery product by separated operand scanning: full 512-bit product, then four reduction
  // steps.  Valid for a < 2^256 arbitrary when b < r (used by fromBytes).
  static Fr montMul(const Fr& a, const Fr& b) {
    uint64_t t[9] = {0};
    for (int i = 0; i < 4; i++) {
      unsigned __int128 carry = 0;
      for (int j = 0; j < 4; j++) {
        carry += (unsigned __int128)a.limbs[i] * b.limbs[j] + t[i + j];
        t[i + j] = (uint64_t)carry;
        carry >>= 64;
      }
      t[i + 4] = (uint64_t)carry;
    }
    for (int i = 0; i < 4; i++) {
      const uint64_t m = t[i] * kInv;
      unsigned __int128 carry = 0;
      for (int j = 0; j < 4; j++) {
        carry += (unsigned __int128)m * kModulus[j] + t[i + j];
        t[i + j] = (uint64_t)carry;
        carry >>= 64;
      }
      for (int j = i + 4; j < 9 && carry; j++) {
        carry += t[j];
        t[j] = (uint64_t)carry;
        carry >>= 64;
      }
    }
    Fr r = Fr::fromLimbs({t[4], t[5], t[6], t[7]});
    if (t[8] || !r.ltModulus()) r.subModulus();
    return r;
  }
};

}  // namespace zolt::host
