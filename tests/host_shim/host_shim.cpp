// TEST INFRASTRUCTURE: compiles the device headers' host (PTX-emulation) back-end so the
// algorithms in zolt_b200/csrc/*.cuh can be validated against the oracle without a GPU.
// Never linked into the product library.
#include "../../zolt_b200/csrc/curve.cuh"
#include <cstddef>
#include <cstring>
using namespace zolt;

template <class P>
static void binop(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    Fp<P> x, y, r;
    memcpy(x.v, a + 8 * i, 32);
    memcpy(y.v, b + 8 * i, 32);
    switch (op) {
      case 0: r = fp_mul<P>(x, y); break;
      case 1: r = fp_add<P>(x, y); break;
      case 2: r = fp_sub<P>(x, y); break;
      case 3: r = fp_inv<P>(x); break;
      case 4: r = fp_from_mont<P>(x); break;
      case 5: r = fp_to_mont<P>(x); break;
      case 6: r = fp_neg<P>(x); break;
      case 7: r = fp_inv_vartime<P>(x); break;
      case 8: r = fp_sqr<P>(x); break;
      default: r = Fp<P>::zero();
    }
    memcpy(out + 8 * i, r.v, 32);
  }
}

extern "C" void shim_field_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
  if (field == 0) binop<FrParams>(op, a, b, out, n);
  else binop<FqParams>(op, a, b, out, n);
}

static AffineG1 load_affine(const uint64_t* xy) {
  AffineG1 p;
  memcpy(p.x.v, xy, 32);
  memcpy(p.y.v, xy + 4, 32);
  return p;
}

// Fold a list of affine points with the XYZZ formulas.  mode 0: sequential mixed additions;
// mode 1: pairwise tree of full additions; mode 2: sum, then doubled `dbl` times; mode 3: sum
// multiplied by the small constant `dbl` (xyzz_mul_small); mode 4: sequential xyzz_madd_uniform.  Output: canonical affine + flag.
extern "C" void shim_g1_fold(int mode, const uint64_t* xy, const uint8_t* neg, size_t n, uint32_t dbl, uint64_t* out_xy,
                             uint8_t* out_inf) {
  XYZZ acc = XYZZ::identity();
  if (mode == 1) {
    XYZZ* tmp = new XYZZ[n ? n : 1];
    for (size_t i = 0; i < n; i++) {
      AffineG1 p = load_affine(xy + 8 * i);
      if (neg && neg[i]) p = affine_neg(p);
      tmp[i] = XYZZ::identity();
      xyzz_madd(tmp[i], p);
    }
    for (size_t m = n; m > 1;) {
      size_t half = (m + 1) / 2;
      for (size_t i = 0; i + half < m; i++) xyzz_add(tmp[i], tmp[i + half]);
      m = half;
    }
    if (n) acc = tmp[0];
    delete[] tmp;
  } else {
    for (size_t i = 0; i < n; i++) {
      AffineG1 p = load_affine(xy + 8 * i);
      if (neg && neg[i]) p = affine_neg(p);
      if (mode == 4) xyzz_madd_uniform(acc, p);
      else xyzz_madd(acc, p);
    }
    if (mode == 2)
      for (uint32_t j = 0; j < dbl; j++) acc = xyzz_dbl(acc);
    if (mode == 3) acc = xyzz_mul_small(acc, dbl);
  }
  AffineG1 a;
  bool finite = xyzz_to_affine(acc, a);
  memcpy(out_xy, a.x.v, 32);
  memcpy(out_xy + 4, a.y.v, 32);
  *out_inf = finite ? 0 : 1;
}

// Delayed reduction (field.cuh wide_*): out = sum_i a_i * b_i * 2^-256 mod p through one 512-bit
// accumulator that is normalised after every term, exactly like the sumcheck kernels use it.
template <class P>
static void wide_dot(const uint32_t* a, const uint32_t* b, size_t n, uint32_t* out) {
  uint32_t acc[16];
  wide_zero(acc);
  for (size_t i = 0; i < n; i++) {
    wide_mul_add(acc, a + 8 * i, b + 8 * i);
    wide_normalize<P>(acc);
  }
  Fp<P> r = wide_reduce<P>(acc);
  memcpy(out, r.v, 32);
}

extern "C" void shim_wide_dot(int field, const uint32_t* a, const uint32_t* b, size_t n, uint32_t* out) {
  if (field == 0) wide_dot<FrParams>(a, b, n, out);
  else wide_dot<FqParams>(a, b, n, out);
}

// fp_bind_fixed: out[i] = lo[i] + r * (hi[i] - lo[i]) through the per-launch table of r (raw limbs in, raw limbs out;
// the arithmetic is agnostic of what the limbs stand for: out = lo + r_std * (hi - lo) mod p with r_std = r / 2^256).
template <class P>
static void bind_fixed(const uint32_t* r, const uint32_t* lo, const uint32_t* hi, size_t n, uint32_t* out) {
  Fp<P> rr;
  memcpy(rr.v, r, 32);
  const FixedMulTable t = fixed_mul_table<P>(rr);
  for (size_t i = 0; i < n; i++) {
    Fp<P> a, b;
    memcpy(a.v, lo + 8 * i, 32);
    memcpy(b.v, hi + 8 * i, 32);
    const Fp<P> v = fp_bind_fixed<P>(a, b, t);
    memcpy(out + 8 * i, v.v, 32);
  }
}
template <class P>
static void bind_fixed_f64(const uint32_t* r, const uint32_t* lo, const uint32_t* hi, size_t n, uint32_t* out) {
  Fp<P> rr;
  memcpy(rr.v, r, 32);
  const FixedMulTableF64 t = fixed_mul_table_f64<P>(rr);
  for (size_t i = 0; i < n; i++) {
    Fp<P> a, b;
    memcpy(a.v, lo + 8 * i, 32);
    memcpy(b.v, hi + 8 * i, 32);
    const Fp<P> v = fp_bind_fixed_f64<P>(a, b, t);
    memcpy(out + 8 * i, v.v, 32);
  }
}
// the FP64-pipe form of the same bind (field.cuh fp_bind_fixed_f64)
extern "C" void shim_bind_fixed_f64(int field, const uint32_t* r, const uint32_t* lo, const uint32_t* hi, size_t n, uint32_t* out) {
  if (field == 0) bind_fixed_f64<FrParams>(r, lo, hi, n, out);
  else bind_fixed_f64<FqParams>(r, lo, hi, n, out);
}
extern "C" void shim_bind_fixed(int field, const uint32_t* r, const uint32_t* lo, const uint32_t* hi, size_t n, uint32_t* out) {
  if (field == 0) bind_fixed<FrParams>(r, lo, hi, n, out);
  else bind_fixed<FqParams>(r, lo, hi, n, out);
}

// fp_mul2_sub: out = (a * b - c * d) * 2^-256 mod p with one shared reduction (the MSM's y3)
template <class P>
static void mul2_sub(const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, size_t n, uint32_t* out) {
  for (size_t i = 0; i < n; i++) {
    Fp<P> x, y, z, w;
    memcpy(x.v, a + 8 * i, 32);
    memcpy(y.v, b + 8 * i, 32);
    memcpy(z.v, c + 8 * i, 32);
    memcpy(w.v, d + 8 * i, 32);
    const Fp<P> r = fp_mul2_sub<P>(x, y, z, w);
    memcpy(out + 8 * i, r.v, 32);
  }
}
extern "C" void shim_mul2_sub(int field, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, size_t n,
                              uint32_t* out) {
  if (field == 0) mul2_sub<FrParams>(a, b, c, d, n, out);
  else mul2_sub<FqParams>(a, b, c, d, n, out);
}
