// BN254 G1 group law for the MSM kernels: y^2 = x^3 + 3 over Fq (reference src/msm/mod.zig:11-12,
// :106-115).  Accumulators use extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ,
// ZZ^3 == ZZZ^2, identity <=> ZZ == 0): a mixed addition costs 8M + 2S against 8M + 4S for the
// Jacobian addAffine the reference uses (src/msm/mod.zig:229-274).  The group element computed is
// the same, and every result leaves the device as canonical affine coordinates (toAffine,
// src/msm/mod.zig:178-189), so outputs are bit-identical to the reference's.
//
// All special cases the reference handles are handled here: identity operands, P + P (doubling)
// and P + (-P) (src/msm/mod.zig:259-267, :312-320).
#pragma once
#include "field.cuh"

namespace zolt {

struct AffineG1 {  // 64 B, Montgomery coordinates; the infinity flag travels separately
  Fq x, y;
};

struct XYZZ {  // 128 B
  Fq x, y, zz, zzz;
  static ZOLT_HD XYZZ identity() {
    XYZZ r;
    r.x = Fq::zero();
    r.y = Fq::zero();
    r.zz = Fq::zero();
    r.zzz = Fq::zero();
    return r;
  }
  ZOLT_HD bool is_identity() const { return zz.is_zero(); }
};

ZOLT_HD Fq fq_add(const Fq& a, const Fq& b) { return fp_add<FqParams>(a, b); }
ZOLT_HD Fq fq_sub(const Fq& a, const Fq& b) { return fp_sub<FqParams>(a, b); }
ZOLT_HD Fq fq_mul(const Fq& a, const Fq& b) { return fp_mul<FqParams>(a, b); }
ZOLT_HD Fq fq_sqr(const Fq& a) { return fp_sqr<FqParams>(a); }
// a * b - c * d with one shared Montgomery reduction (field.cuh): the y coordinate of every addition / doubling
ZOLT_HD Fq fq_mul2_sub(const Fq& a, const Fq& b, const Fq& c, const Fq& d) { return fp_mul2_sub<FqParams>(a, b, c, d); }

ZOLT_HD XYZZ xyzz_from_affine(const AffineG1& p) {
  XYZZ r;
  r.x = p.x;
  r.y = p.y;
  r.zz = Fq::one();
  r.zzz = Fq::one();
  return r;
}

// 2*(x, y) for an affine point (mdbl-2008-s-1).  y != 0 on a prime-order curve.
ZOLT_HD XYZZ xyzz_dbl_affine(const AffineG1& p) {
  XYZZ r;
  Fq u = fq_add(p.y, p.y);
  Fq v = fq_sqr(u);
  Fq w = fq_mul(u, v);
  Fq s = fq_mul(p.x, v);
  Fq xx = fq_sqr(p.x);
  Fq m = fq_add(fq_add(xx, xx), xx);
  r.x = fq_sub(fq_sqr(m), fq_add(s, s));
  r.y = fq_mul2_sub(m, fq_sub(s, r.x), w, p.y);
  r.zz = v;
  r.zzz = w;
  return r;
}

// 2*P (dbl-2008-s-1)
ZOLT_HD XYZZ xyzz_dbl(const XYZZ& p) {
  if (p.is_identity()) return p;
  XYZZ r;
  Fq u = fq_add(p.y, p.y);
  Fq v = fq_sqr(u);
  Fq w = fq_mul(u, v);
  Fq s = fq_mul(p.x, v);
  Fq xx = fq_sqr(p.x);
  Fq m = fq_add(fq_add(xx, xx), xx);
  r.x = fq_sub(fq_sqr(m), fq_add(s, s));
  r.y = fq_mul2_sub(m, fq_sub(s, r.x), w, p.y);
  r.zz = fq_mul(v, p.zz);
  r.zzz = fq_mul(w, p.zzz);
  return r;
}

// acc += (x2, y2)   (madd-2008-s), with the reference's special cases
ZOLT_HD void xyzz_madd(XYZZ& acc, const AffineG1& q) {
  if (acc.is_identity()) {
    acc = xyzz_from_affine(q);
    return;
  }
  Fq u2 = fq_mul(q.x, acc.zz);
  Fq s2 = fq_mul(q.y, acc.zzz);
  Fq p = fq_sub(u2, acc.x);
  Fq r = fq_sub(s2, acc.y);
  if (p.is_zero()) {
    if (r.is_zero()) acc = xyzz_dbl_affine(q);  // same point: double
    else acc = XYZZ::identity();                // opposite points
    return;
  }
  Fq pp = fq_sqr(p);
  Fq ppp = fq_mul(p, pp);
  Fq qq = fq_mul(acc.x, pp);
  Fq x3 = fq_sub(fq_sub(fq_sqr(r), ppp), fq_add(qq, qq));
  Fq y3 = fq_mul2_sub(r, fq_sub(qq, x3), acc.y, ppp);
  acc.x = x3;
  acc.y = y3;
  acc.zz = fq_mul(acc.zz, pp);
  acc.zzz = fq_mul(acc.zzz, ppp);
}

// acc += (x2, y2) with warp-uniform control flow: every lane runs the ten field products of the
// general case; identity accumulators are fixed up by a select afterwards, so lanes of a warp that
// sit at different positions of their buckets stay converged.  Only the (rare) P == +-Q cases branch.
ZOLT_HD void xyzz_madd_uniform(XYZZ& acc, const AffineG1& q) {
  const bool ident = acc.is_identity();
  Fq u2 = fq_mul(q.x, acc.zz);
  Fq s2 = fq_mul(q.y, acc.zzz);
  Fq p = fq_sub(u2, acc.x);
  Fq r = fq_sub(s2, acc.y);
  Fq pp = fq_sqr(p);
  Fq ppp = fq_mul(p, pp);
  Fq qq = fq_mul(acc.x, pp);
  Fq x3 = fq_sub(fq_sub(fq_sqr(r), ppp), fq_add(qq, qq));
  Fq y3 = fq_mul2_sub(r, fq_sub(qq, x3), acc.y, ppp);
  Fq zz3 = fq_mul(acc.zz, pp);
  Fq zzz3 = fq_mul(acc.zzz, ppp);
  if (!ident && p.is_zero()) {  // same x: P + P or P + (-P)
    XYZZ d = r.is_zero() ? xyzz_dbl_affine(q) : XYZZ::identity();
    x3 = d.x;
    y3 = d.y;
    zz3 = d.zz;
    zzz3 = d.zzz;
  }
  const Fq one = Fq::one();
#pragma unroll
  for (int i = 0; i < 8; i++) {
    acc.x.v[i] = ident ? q.x.v[i] : x3.v[i];
    acc.y.v[i] = ident ? q.y.v[i] : y3.v[i];
    acc.zz.v[i] = ident ? one.v[i] : zz3.v[i];
    acc.zzz.v[i] = ident ? one.v[i] : zzz3.v[i];
  }
}

// acc += b   (add-2008-s)
ZOLT_HD void xyzz_add(XYZZ& acc, const XYZZ& b) {
  if (b.is_identity()) return;
  if (acc.is_identity()) {
    acc = b;
    return;
  }
  Fq u1 = fq_mul(acc.x, b.zz);
  Fq u2 = fq_mul(b.x, acc.zz);
  Fq s1 = fq_mul(acc.y, b.zzz);
  Fq s2 = fq_mul(b.y, acc.zzz);
  Fq p = fq_sub(u2, u1);
  Fq r = fq_sub(s2, s1);
  if (p.is_zero()) {
    if (r.is_zero()) acc = xyzz_dbl(acc);
    else acc = XYZZ::identity();
    return;
  }
  Fq pp = fq_sqr(p);
  Fq ppp = fq_mul(p, pp);
  Fq qq = fq_mul(u1, pp);
  Fq x3 = fq_sub(fq_sub(fq_sqr(r), ppp), fq_add(qq, qq));
  Fq y3 = fq_mul2_sub(r, fq_sub(qq, x3), s1, ppp);
  acc.x = x3;
  acc.y = y3;
  acc.zz = fq_mul(fq_mul(acc.zz, b.zz), pp);
  acc.zzz = fq_mul(fq_mul(acc.zzz, b.zzz), ppp);
}

ZOLT_HD AffineG1 affine_neg(const AffineG1& p) {
  AffineG1 r;
  r.x = p.x;
  r.y = fp_neg<FqParams>(p.y);
  return r;
}

// k * P for a small unsigned k (double-and-add, MSB first)
ZOLT_HD XYZZ xyzz_mul_small(const XYZZ& p, uint32_t k) {
  if (k == 0) return XYZZ::identity();
  int top = 31;
  while (!((k >> top) & 1u)) top--;  // start at the leading one: no doublings of the identity
  XYZZ acc = p;
  for (int bit = top - 1; bit >= 0; bit--) {
    acc = xyzz_dbl(acc);
    if ((k >> bit) & 1u) xyzz_add(acc, p);
  }
  return acc;
}

// XYZZ -> canonical affine (reference toAffine, src/msm/mod.zig:178-189).  Returns false (and
// x = y = 0) for the identity, which the reference encodes as (0, 0, infinity = true) (:24-30).
ZOLT_HD bool xyzz_to_affine(const XYZZ& p, AffineG1& out) {
  if (p.is_identity()) {
    out.x = Fq::zero();
    out.y = Fq::zero();
    return false;
  }
  // one inversion: ZZ^3 == ZZZ^2, hence ZZ^-1 = ZZZ^-2 * ZZ^2
  Fq izzz = fp_inv_vartime<FqParams>(p.zzz);
  Fq izz = fq_mul(fq_sqr(izzz), fq_sqr(p.zz));  // ZZZ^-2 * ZZ^2 = ZZ^-3 * ZZ^2 = ZZ^-1
  out.x = fq_mul(p.x, izz);
  out.y = fq_mul(p.y, izzz);
  return true;
}

}  // namespace zolt
