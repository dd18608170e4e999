// ORACLE / CPU BASELINE (test infrastructure only -- never linked into or called by the product).
//
// C++17 restatement of Zolt's CPU algorithms for the hot path, used (a) as a fast checker at sizes the
// Python big-int oracle cannot reach and (b) as the TIMED CPU baseline (bench.py cpu_baseline /
// --impl reference), because the reference itself cannot be built here (no zig toolchain).
// The algorithms follow the reference as written so the timing is representative:
//   montgomeryMul 4x64 CIOS with u128         src/field/mod.zig:735-779 (generic :269-308)
//   add / sub / neg / inverse (Fermat)        src/field/mod.zig:782-816, :944-983
//   Jacobian double / addAffine / add         src/msm/mod.zig:195-327 (incl. the repeated Y*Z, Y1*J)
//   toAffine                                  src/msm/mod.zig:178-189
//   MSM.compute / pippengerMSM / getWindow    src/msm/mod.zig:355-471 (per-call fromMontgomery)
//   optimalWindowSize / naiveMSM / scalarMul  src/msm/mod.zig:475-540
//   ParallelMSM.compute (<= 8 threads)        src/msm/mod.zig:588-678
//   bindLow / bindFirst / eq evals            src/poly/mod.zig:128-175, :252-290
//   per-prover round loops                    see each function
//   Keccak Transcript                         src/transcripts/mod.zig:49-221
// Pinned against: the Python big-int oracle (tests/test_oracle_cpu_ref.py) and, through it, the
// reference's own unit-test vectors and logs/zolt.log KATs (tests/test_oracle_golden.py).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

typedef unsigned __int128 u128;

namespace ref {

struct FrTag {
  static constexpr uint64_t MOD[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  static constexpr uint64_t ONE[4] = {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full};
  static constexpr uint64_t R2[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
  static constexpr uint64_t INV = 0xc2e1f593efffffffull;
};
struct FqTag {
  static constexpr uint64_t MOD[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  static constexpr uint64_t ONE[4] = {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full};
  static constexpr uint64_t R2[4] = {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full};
  static constexpr uint64_t INV = 0x87d20782e4866389ull;
};

template <class T>
struct Fe {
  uint64_t l[4];

  static Fe zero() { return Fe{{0, 0, 0, 0}}; }
  static Fe one() { return Fe{{T::ONE[0], T::ONE[1], T::ONE[2], T::ONE[3]}}; }
  bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
  bool eq(const Fe& o) const { return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3]; }

  bool lt_mod() const {
    for (int i = 3; i >= 0; i--) {
      if (l[i] < T::MOD[i]) return true;
      if (l[i] > T::MOD[i]) return false;
    }
    return false;
  }
  Fe sub_mod() const {
    Fe r;
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
      u128 d = (u128)l[i] - T::MOD[i] - borrow;
      r.l[i] = (uint64_t)d;
      borrow = (uint64_t)(d >> 64) & 1;
    }
    return r;
  }
  // CIOS, src/field/mod.zig:735-779
  Fe mul(const Fe& o) const {
    uint64_t t[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
      uint64_t carry = 0;
      for (int j = 0; j < 4; j++) {
        u128 s = (u128)t[j] + (u128)l[i] * o.l[j] + carry;
        t[j] = (uint64_t)s;
        carry = (uint64_t)(s >> 64);
      }
      t[4] += carry;
      uint64_t m = t[0] * T::INV;
      u128 s0 = (u128)t[0] + (u128)m * T::MOD[0];
      carry = (uint64_t)(s0 >> 64);
      for (int j = 1; j < 4; j++) {
        u128 s = (u128)t[j] + (u128)m * T::MOD[j] + carry;
        t[j - 1] = (uint64_t)s;
        carry = (uint64_t)(s >> 64);
      }
      u128 f = (u128)t[4] + carry;
      t[3] = (uint64_t)f;
      t[4] = (uint64_t)(f >> 64);
    }
    Fe r{{t[0], t[1], t[2], t[3]}};
    if (t[4] != 0 || !r.lt_mod()) r = r.sub_mod();
    return r;
  }
  Fe sqr() const { return mul(*this); }
  Fe add(const Fe& o) const {
    Fe r;
    uint64_t carry = 0;
    for (int i = 0; i < 4; i++) {
      u128 s = (u128)l[i] + o.l[i] + carry;
      r.l[i] = (uint64_t)s;
      carry = (uint64_t)(s >> 64);
    }
    if (carry || !r.lt_mod()) r = r.sub_mod();
    return r;
  }
  Fe sub(const Fe& o) const {
    Fe r;
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
      u128 d = (u128)l[i] - o.l[i] - borrow;
      r.l[i] = (uint64_t)d;
      borrow = (uint64_t)(d >> 64) & 1;
    }
    if (borrow) {
      uint64_t carry = 0;
      for (int i = 0; i < 4; i++) {
        u128 s = (u128)r.l[i] + T::MOD[i] + carry;
        r.l[i] = (uint64_t)s;
        carry = (uint64_t)(s >> 64);
      }
    }
    return r;
  }
  Fe neg() const {
    if (is_zero()) return *this;
    Fe m{{T::MOD[0], T::MOD[1], T::MOD[2], T::MOD[3]}};
    return m.sub(*this);
  }
  Fe from_mont() const { return mul(Fe{{1, 0, 0, 0}}); }
  Fe to_mont() const { return mul(Fe{{T::R2[0], T::R2[1], T::R2[2], T::R2[3]}}); }
  static Fe from_u64(uint64_t v) { return Fe{{v, 0, 0, 0}}.to_mont(); }
  // Fermat, LSB-first over all 256 exponent bits (src/field/mod.zig:955-983); 0 -> 0
  Fe inverse() const {
    if (is_zero()) return *this;
    uint64_t e[4] = {T::MOD[0] - 2, T::MOD[1], T::MOD[2], T::MOD[3]};
    Fe result = one(), base = *this;
    for (int i = 0; i < 4; i++) {
      uint64_t bits = e[i];
      for (int j = 0; j < 64; j++) {
        if (bits & 1) result = result.mul(base);
        base = base.sqr();
        bits >>= 1;
      }
    }
    return result;
  }
};

using Fr = Fe<FrTag>;
using Fq = Fe<FqTag>;

// ---- curve (src/msm/mod.zig:15-329) -----------------------------------------------------------------
struct Affine {
  Fq x, y;
  bool inf;
  static Affine identity() { return Affine{Fq::zero(), Fq::zero(), true}; }
};

struct Jac {
  Fq x, y, z;
  static Jac identity() { return Jac{Fq::one(), Fq::one(), Fq::zero()}; }
  static Jac from_affine(const Affine& p) { return p.inf ? identity() : Jac{p.x, p.y, Fq::one()}; }
  bool is_identity() const { return z.is_zero(); }

  Affine to_affine() const {
    if (z.is_zero()) return Affine::identity();
    Fq zi = z.inverse();
    Fq zi2 = zi.sqr();
    Fq zi3 = zi2.mul(zi);
    return Affine{x.mul(zi2), y.mul(zi3), false};
  }
  // dbl-2009-l as written, src/msm/mod.zig:195-226
  Jac dbl() const {
    if (z.is_zero()) return *this;
    Fq A = x.sqr(), B = y.sqr(), C = B.sqr();
    Fq xpb = x.add(B);
    Fq half_D = xpb.sqr().sub(A).sub(C);
    Fq D = half_D.add(half_D);
    Fq E = A.add(A).add(A);
    Fq FF = E.sqr();
    Fq X3 = FF.sub(D.add(D));
    Fq C8 = C.add(C).add(C).add(C).add(C).add(C).add(C).add(C);
    Fq Y3 = E.mul(D.sub(X3)).sub(C8);
    Fq Z3 = y.mul(z).add(y.mul(z));
    return Jac{X3, Y3, Z3};
  }
  // madd as written, src/msm/mod.zig:229-274
  Jac add_affine(const Affine& o) const {
    if (o.inf) return *this;
    if (z.is_zero()) return from_affine(o);
    Fq z1z1 = z.sqr();
    Fq U2 = o.x.mul(z1z1);
    Fq S2 = o.y.mul(z).mul(z1z1);
    Fq H = U2.sub(x);
    Fq HH = H.sqr();
    Fq I = HH.add(HH).add(HH).add(HH);
    Fq J = H.mul(I);
    Fq r = S2.sub(y).add(S2.sub(y));
    Fq V = x.mul(I);
    Fq X3 = r.sqr().sub(J).sub(V).sub(V);
    Fq Y3 = r.mul(V.sub(X3)).sub(y.mul(J)).sub(y.mul(J));
    Fq Z3 = z.add(H).sqr().sub(z1z1).sub(HH);
    if (H.is_zero()) return r.is_zero() ? dbl() : identity();
    return Jac{X3, Y3, Z3};
  }
  // add-2007-bl as written, src/msm/mod.zig:277-327
  Jac add(const Jac& o) const {
    if (z.is_zero()) return o;
    if (o.z.is_zero()) return *this;
    Fq z1z1 = z.sqr(), z2z2 = o.z.sqr();
    Fq U1 = x.mul(z2z2), U2 = o.x.mul(z1z1);
    Fq S1 = y.mul(o.z).mul(z2z2), S2 = o.y.mul(z).mul(z1z1);
    Fq H = U2.sub(U1);
    Fq twoH = H.add(H);
    Fq I = twoH.sqr();
    Fq J = H.mul(I);
    Fq r = S2.sub(S1).add(S2.sub(S1));
    Fq V = U1.mul(I);
    Fq X3 = r.sqr().sub(J).sub(V).sub(V);
    Fq Y3 = r.mul(V.sub(X3)).sub(S1.mul(J)).sub(S1.mul(J));
    Fq Z3 = z.add(o.z).sqr().sub(z1z1).sub(z2z2).mul(H);
    if (H.is_zero()) return r.is_zero() ? dbl() : identity();
    return Jac{X3, Y3, Z3};
  }
};

// ---- MSM (src/msm/mod.zig:345-542) --------------------------------------------------------------------
static size_t optimal_window_size(size_t n) {
  if (n < 8) return 1;
  if (n < 32) return 2;
  if (n < 128) return 3;
  if (n < 512) return 4;
  if (n < 2048) return 5;
  if (n < 8192) return 6;
  if (n < 32768) return 7;
  return 8;
}

static size_t get_window(const Fr& scalar, size_t window_idx, size_t c) {
  const Fr s = scalar.from_mont();  // per call, as the reference does (:442)
  const size_t bit_offset = window_idx * c;
  const size_t limb_idx = bit_offset / 64, bit_in_limb = bit_offset % 64;
  if (limb_idx >= 4) return 0;
  const uint64_t mask = (1ull << c) - 1;
  uint64_t value = (s.l[limb_idx] >> bit_in_limb) & mask;
  if (bit_in_limb + c > 64 && limb_idx + 1 < 4) {
    const size_t remaining = bit_in_limb + c - 64;
    if (remaining > 0 && remaining <= 63 && bit_in_limb > 0) {
      const uint64_t next_mask = (1ull << remaining) - 1;
      value |= (s.l[limb_idx + 1] & next_mask) << (64 - bit_in_limb);
    }
  }
  return (size_t)(value & mask);
}

static Jac scalar_mul(const Affine& base, const Fr& scalar) {
  if (base.inf || scalar.is_zero()) return Jac::identity();
  Jac result = Jac::identity();
  const Fr s = scalar.from_mont();
  for (int limb = 3; limb >= 0; limb--) {
    for (int bit = 63; bit >= 0; bit--) {
      if (!result.is_identity()) result = result.dbl();
      if ((s.l[limb] >> bit) & 1) result = result.is_identity() ? Jac::from_affine(base) : result.add_affine(base);
    }
  }
  return result;
}

static Affine pippenger(const Affine* bases, const Fr* scalars, size_t n) {
  const size_t c = optimal_window_size(n);
  const size_t num_windows = (256 + c - 1) / c;
  const size_t num_buckets = (size_t(1) << c) - 1;
  Jac final_result = Jac::identity();
  for (size_t w = num_windows; w-- > 0;) {
    if (!final_result.is_identity())
      for (size_t i = 0; i < c; i++) final_result = final_result.dbl();
    Jac buckets[256];
    for (size_t j = 0; j < num_buckets; j++) buckets[j] = Jac::identity();
    for (size_t i = 0; i < n; i++) {
      if (bases[i].inf) continue;
      const size_t d = get_window(scalars[i], w, c);
      if (d == 0) continue;
      buckets[d - 1] = buckets[d - 1].add_affine(bases[i]);
    }
    Jac running = Jac::identity(), window_sum = Jac::identity();
    for (size_t b = num_buckets; b-- > 0;) {
      running = running.add(buckets[b]);
      window_sum = window_sum.add(running);
    }
    final_result = final_result.add(window_sum);
  }
  return final_result.to_affine();
}

static Affine msm_compute(const Affine* bases, const Fr* scalars, size_t n) {
  if (n == 0) return Affine::identity();
  if (n < 8) {
    Jac result = Jac::identity();
    for (size_t i = 0; i < n; i++) result = result.add(scalar_mul(bases[i], scalars[i]));
    return result.to_affine();
  }
  return pippenger(bases, scalars, n);
}

// ParallelMSM.compute (src/msm/mod.zig:588-653): contiguous chunks, one thread each, partials added.
static Affine msm_parallel(const Affine* bases, const Fr* scalars, size_t n, size_t num_threads) {
  if (n == 0) return Affine::identity();
  const size_t min_points = 1024;
  size_t actual = std::min(num_threads, std::max<size_t>(1, n / min_points));
  if (actual <= 1) return msm_compute(bases, scalars, n);
  const size_t chunk = (n + actual - 1) / actual;
  std::vector<Affine> partial(actual, Affine::identity());
  std::vector<std::thread> threads;
  for (size_t t = 0; t < actual; t++) {
    const size_t lo = t * chunk, hi = std::min(n, lo + chunk);
    if (lo >= hi) break;
    threads.emplace_back([&, t, lo, hi] { partial[t] = msm_compute(bases + lo, scalars + lo, hi - lo); });
  }
  for (auto& th : threads) th.join();
  Jac acc = Jac::identity();
  for (auto& p : partial) acc = acc.add(Jac::from_affine(p));
  return acc.to_affine();
}

// ---- Keccak transcript (src/transcripts/mod.zig:49-221) ----------------------------------------------------
struct Transcript {
  uint8_t state[200];
  size_t position;
};
static const uint64_t RC[24] = {
    0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull,
    0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull,
    0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull,
    0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
    0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
static const int ROTC[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
static const int PILN[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
static inline uint64_t rotl(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

static void keccak_f(uint8_t* state) {
  uint64_t st[25];
  memcpy(st, state, 200);  // little-endian host
  for (int rnd = 0; rnd < 24; rnd++) {
    uint64_t bc[5];
    for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
    for (int i = 0; i < 5; i++) {
      uint64_t t = bc[(i + 4) % 5] ^ rotl(bc[(i + 1) % 5], 1);
      for (int j = i; j < 25; j += 5) st[j] ^= t;
    }
    uint64_t t = st[1];
    for (int i = 0; i < 24; i++) {
      int j = PILN[i];
      uint64_t tmp = st[j];
      st[j] = rotl(t, ROTC[i]);
      t = tmp;
    }
    for (int row = 0; row < 25; row += 5) {
      for (int i = 0; i < 5; i++) bc[i] = st[row + i];
      for (int i = 0; i < 5; i++) st[row + i] = bc[i] ^ (~bc[(i + 1) % 5] & bc[(i + 2) % 5]);
    }
    st[0] ^= RC[rnd];
  }
  memcpy(state, st, 200);
}
static void tr_append(Transcript* t, const uint8_t* data, size_t n) {
  for (size_t i = 0; i < n; i++) {
    t->state[t->position] ^= data[i];
    if (++t->position >= 136) {
      keccak_f(t->state);
      t->position = 0;
    }
  }
}

}  // namespace ref

using namespace ref;

// ------------------------------------------------------------------------------------------------------
// C exports (ctypes)
// ------------------------------------------------------------------------------------------------------
extern "C" {

// op: 0 mul, 1 add, 2 sub, 3 inverse, 4 fromMontgomery, 5 toMontgomery, 6 neg
void ref_field_op(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    if (field == 0) {
      Fr x, y = Fr::zero(), r;
      memcpy(x.l, a + 4 * i, 32);
      if (b) memcpy(y.l, b + 4 * i, 32);
      switch (op) {
        case 0: r = x.mul(y); break;
        case 1: r = x.add(y); break;
        case 2: r = x.sub(y); break;
        case 3: r = x.inverse(); break;
        case 4: r = x.from_mont(); break;
        case 5: r = x.to_mont(); break;
        default: r = x.neg(); break;
      }
      memcpy(out + 4 * i, r.l, 32);
    } else {
      Fq x, y = Fq::zero(), r;
      memcpy(x.l, a + 4 * i, 32);
      if (b) memcpy(y.l, b + 4 * i, 32);
      switch (op) {
        case 0: r = x.mul(y); break;
        case 1: r = x.add(y); break;
        case 2: r = x.sub(y); break;
        case 3: r = x.inverse(); break;
        case 4: r = x.from_mont(); break;
        case 5: r = x.to_mont(); break;
        default: r = x.neg(); break;
      }
      memcpy(out + 4 * i, r.l, 32);
    }
  }
}

// sum_i a_i * b_i (Montgomery form in, Montgomery form out): BatchOps.innerProduct (src/field/mod.zig:1164-1280), split
// over `threads` contiguous chunks.  bench.py uses it, untimed, for the closed-form check (sum s_i k_i) * G of the MSM.
void ref_fr_dot(const uint64_t* a, const uint64_t* b, size_t n, int threads, uint64_t* out) {
  if (threads < 1) threads = 1;
  std::vector<Fr> part((size_t)threads, Fr::zero());
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++) {
    pool.emplace_back([&, t]() {
      const size_t lo = n * (size_t)t / (size_t)threads, hi = n * (size_t)(t + 1) / (size_t)threads;
      Fr acc = Fr::zero();
      for (size_t i = lo; i < hi; i++) {
        Fr x, y;
        memcpy(x.l, a + 4 * i, 32);
        memcpy(y.l, b + 4 * i, 32);
        acc = acc.add(x.mul(y));
      }
      part[(size_t)t] = acc;
    });
  }
  for (auto& th : pool) th.join();
  Fr acc = Fr::zero();
  for (auto& p : part) acc = acc.add(p);
  memcpy(out, acc.l, 32);
}

static std::vector<Affine> load_points(const uint64_t* xy, const uint8_t* inf, size_t n) {
  std::vector<Affine> pts(n);
  for (size_t i = 0; i < n; i++) {
    memcpy(pts[i].x.l, xy + 8 * i, 32);
    memcpy(pts[i].y.l, xy + 8 * i + 4, 32);
    pts[i].inf = inf ? inf[i] != 0 : false;
  }
  return pts;
}
static void store_point(const Affine& p, uint64_t* out_xy, uint8_t* out_inf) {
  memcpy(out_xy, p.x.l, 32);
  memcpy(out_xy + 4, p.y.l, 32);
  *out_inf = p.inf ? 1 : 0;
}

// MSM(F,Fp).compute; threads <= 1 is the path HyperKZG.commit actually calls, threads > 1 is
// ParallelMSM.compute with that many threads (the reference caps detectOptimalThreads at 8).
void ref_msm(const uint64_t* xy, const uint8_t* inf, const uint64_t* scalars, size_t n, int threads, uint64_t* out_xy,
             uint8_t* out_inf) {
  std::vector<Affine> pts = load_points(xy, inf, n);
  const Fr* sc = reinterpret_cast<const Fr*>(scalars);
  Affine r = threads > 1 ? msm_parallel(pts.data(), sc, n, (size_t)threads) : msm_compute(pts.data(), sc, n);
  store_point(r, out_xy, out_inf);
}

// bases[i] = scalars[i] * G via scalarMul (src/msm/mod.zig:503-540), `threads`-way split for speed
void ref_g1_fixed_mul(const uint64_t* scalars, size_t n, int threads, uint64_t* out_xy, uint8_t* out_inf) {
  Affine gen{Fq::one(), Fq::one().add(Fq::one()), false};
  const Fr* sc = reinterpret_cast<const Fr*>(scalars);
  auto work = [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) store_point(scalar_mul(gen, sc[i]).to_affine(), out_xy + 8 * i, out_inf + i);
  };
  if (threads <= 1) { work(0, n); return; }
  std::vector<std::thread> th;
  const size_t chunk = (n + threads - 1) / threads;
  for (int t = 0; t < threads; t++) {
    size_t lo = t * chunk, hi = std::min(n, lo + chunk);
    if (lo < hi) th.emplace_back(work, lo, hi);
  }
  for (auto& x : th) x.join();
}

// bases[i] = (i+1) * G (the bench.zig base set, src/bench.zig:251-262) by repeated mixed addition and
// one batch inversion -- a fast way to get many distinct valid bases for the timed CPU baseline.
void ref_g1_multiples(size_t n, uint64_t* out_xy, uint8_t* out_inf) {
  Affine gen{Fq::one(), Fq::one().add(Fq::one()), false};
  std::vector<Jac> pts(n);
  Jac acc = Jac::identity();
  for (size_t i = 0; i < n; i++) {
    acc = acc.add_affine(gen);
    pts[i] = acc;
  }
  std::vector<Fq> prefix(n);
  Fq run = Fq::one();
  for (size_t i = 0; i < n; i++) {
    prefix[i] = run;
    run = run.mul(pts[i].z);
  }
  Fq inv = run.inverse();
  for (size_t i = n; i-- > 0;) {
    Fq zi = inv.mul(prefix[i]);
    inv = inv.mul(pts[i].z);
    Fq zi2 = zi.sqr();
    store_point(Affine{pts[i].x.mul(zi2), pts[i].y.mul(zi2.mul(zi)), false}, out_xy + 8 * i, out_inf + i);
  }
}

// DensePolynomial.bindLow in place (src/poly/mod.zig:160-175): n = current length
void ref_bind_low(uint64_t* table, size_t n, const uint64_t* r) {
  Fr* t = reinterpret_cast<Fr*>(table);
  Fr rr;
  memcpy(rr.l, r, 32);
  for (size_t i = 0; i < n / 2; i++) {
    Fr lo = t[2 * i], hi = t[2 * i + 1];
    t[i] = lo.add(rr.mul(hi.sub(lo)));
  }
}
// DensePolynomial.bindFirst (src/poly/mod.zig:128-149): out has n/2 entries
void ref_bind_first(const uint64_t* table, size_t n, const uint64_t* r, uint64_t* out) {
  const Fr* t = reinterpret_cast<const Fr*>(table);
  Fr* o = reinterpret_cast<Fr*>(out);
  Fr rr;
  memcpy(rr.l, r, 32);
  Fr omr = Fr::one().sub(rr);
  const size_t h = n / 2;
  for (size_t i = 0; i < h; i++) o[i] = t[i].mul(omr).add(t[i + h].mul(rr));
}
// the (1-r)*lo + r*hi adjacent-pair fold used by the stage provers (val_evaluation.zig:624-631)
void ref_bind_adjacent_2mul(uint64_t* table, size_t n, const uint64_t* r) {
  Fr* t = reinterpret_cast<Fr*>(table);
  Fr rr;
  memcpy(rr.l, r, 32);
  Fr omr = Fr::one().sub(rr);
  for (size_t i = 0; i < n / 2; i++) t[i] = omr.mul(t[2 * i]).add(rr.mul(t[2 * i + 1]));
}
// EqPolynomial.evalsSliceWithScaling (src/poly/mod.zig:252-290)
void ref_eq_evals(const uint64_t* r, int n_vars, const uint64_t* scale, uint64_t* out) {
  Fr* res = reinterpret_cast<Fr*>(out);
  const Fr* rv = reinterpret_cast<const Fr*>(r);
  const size_t total = size_t(1) << n_vars;
  for (size_t i = 0; i < total; i++) res[i] = Fr::zero();
  if (scale) memcpy(res[0].l, scale, 32); else res[0] = Fr::one();
  size_t size = 1;
  for (int j = n_vars - 1; j >= 0; j--) {
    for (size_t i = 0; i < size; i++) {
      Fr x = res[i];
      Fr y = x.mul(rv[j]);
      res[i + size] = y;
      res[i] = x.sub(y);
    }
    size *= 2;
  }
}

// Round loops "as written".  kind: 0 sum (subprotocols/mod.zig:84-93, jolt_r1cs.zig:436-439),
// 1 prod2 (val_final.zig:164-184), 2 prod3 (val_evaluation.zig:567-601), 3 eq*a*(b-c)
// (output_check.zig:385-433), 4 eq*(az*bz-cz) (config-3 shape, same slope-stepping style as 3).
// pairing: 0 adjacent, 1 half split.  out: 4 evaluations (kind 0: [lo, hi, 0, 0]).
void ref_sc_round(int kind, int pairing, const uint64_t* const* tables, size_t n, uint64_t* out) {
  const size_t half = n / 2;
  auto T = [&](int k) { return reinterpret_cast<const Fr*>(tables[k]); };
  auto LO = [&](int k, size_t i) { return pairing == 0 ? T(k)[2 * i] : T(k)[i]; };
  auto HI = [&](int k, size_t i) { return pairing == 0 ? T(k)[2 * i + 1] : T(k)[i + half]; };
  Fr e[4] = {Fr::zero(), Fr::zero(), Fr::zero(), Fr::zero()};
  if (kind == 0) {
    for (size_t i = 0; i < half; i++) e[0] = e[0].add(LO(0, i));
    for (size_t i = 0; i < half; i++) e[1] = e[1].add(HI(0, i));
  } else if (kind == 1) {
    const Fr two = Fr::from_u64(2), three = Fr::from_u64(3);
    for (size_t i = 0; i < half; i++) {
      Fr a0 = LO(0, i), b0 = LO(1, i), a1 = HI(0, i), b1 = HI(1, i);
      e[0] = e[0].add(a0.mul(b0));
      e[1] = e[1].add(a1.mul(b1));
      Fr a2 = two.mul(a1).sub(a0), b2 = two.mul(b1).sub(b0);
      e[2] = e[2].add(a2.mul(b2));
      Fr a3 = three.mul(a1).sub(two.mul(a0)), b3 = three.mul(b1).sub(two.mul(b0));
      e[3] = e[3].add(a3.mul(b3));
    }
  } else if (kind == 2) {
    for (size_t i = 0; i < half; i++) {
      Fr a0 = LO(0, i), b0 = LO(1, i), c0 = LO(2, i), a1 = HI(0, i), b1 = HI(1, i), c1 = HI(2, i);
      e[0] = e[0].add(a0.mul(b0).mul(c0));
      e[1] = e[1].add(a1.mul(b1).mul(c1));
      const Fr two = Fr::from_u64(2), three = Fr::from_u64(3);  // inside the loop, as written (:587-588)
      Fr a2 = two.mul(a1).sub(a0), b2 = two.mul(b1).sub(b0), c2 = two.mul(c1).sub(c0);
      e[2] = e[2].add(a2.mul(b2).mul(c2));
      Fr a3 = three.mul(a1).sub(two.mul(a0)), b3 = three.mul(b1).sub(two.mul(b0)), c3 = three.mul(c1).sub(two.mul(c0));
      e[3] = e[3].add(a3.mul(b3).mul(c3));
    }
  } else if (kind == 3) {
    for (size_t g = 0; g < half; g++) {
      Fr eq0 = LO(0, g), eq1 = HI(0, g), io0 = LO(1, g), io1 = HI(1, g);
      Fr v0 = LO(2, g).sub(LO(3, g)), v1 = HI(2, g).sub(HI(3, g));
      Fr deq = eq1.sub(eq0), dio = io1.sub(io0), dv = v1.sub(v0);
      Fr p0 = eq0.mul(io0).mul(v0), p1 = eq1.mul(io1).mul(v1);
      Fr eq2 = eq0.add(deq).add(deq), io2 = io0.add(dio).add(dio), v2 = v0.add(dv).add(dv);
      Fr p2 = eq2.mul(io2).mul(v2);
      Fr p3 = eq2.add(deq).mul(io2.add(dio)).mul(v2.add(dv));
      e[0] = e[0].add(p0); e[1] = e[1].add(p1); e[2] = e[2].add(p2); e[3] = e[3].add(p3);
    }
  } else {
    for (size_t g = 0; g < half; g++) {
      Fr q = LO(0, g), a = LO(1, g), b = LO(2, g), c = LO(3, g);
      Fr dq = HI(0, g).sub(q), da = HI(1, g).sub(a), db = HI(2, g).sub(b), dc = HI(3, g).sub(c);
      for (int x = 0; x < 4; x++) {
        e[x] = e[x].add(q.mul(a.mul(b).sub(c)));
        q = q.add(dq); a = a.add(da); b = b.add(db); c = c.add(dc);
      }
    }
  }
  memcpy(out, e, 128);
}

// Gruen round (product_remainder.zig:283-322): out = [t0, t_inf]; with_c adds the -cz term of config 3
void ref_sc_gruen_round(const uint64_t* const* tables, size_t n, const uint64_t* e_out, size_t n_out,
                        const uint64_t* e_in, size_t n_in, int with_c, uint64_t* out) {
  const Fr* left = reinterpret_cast<const Fr*>(tables[0]);
  const Fr* right = reinterpret_cast<const Fr*>(tables[1]);
  const Fr* cz = with_c ? reinterpret_cast<const Fr*>(tables[2]) : nullptr;
  const Fr* Eo = reinterpret_cast<const Fr*>(e_out);
  const Fr* Ei = reinterpret_cast<const Fr*>(e_in);
  const size_t groups = n / 2;
  int bits = 0;
  while ((size_t(1) << bits) < n_in) bits++;
  Fr t0 = Fr::zero(), tinf = Fr::zero();
  for (size_t xo = 0; xo < n_out; xo++) {
    Fr i0 = Fr::zero(), ii = Fr::zero();
    for (size_t xi = 0; xi < n_in; xi++) {
      const size_t g = (xo << bits) | xi;
      if (g < groups) {
        Fr l_lo = left[2 * g], l_hi = left[2 * g + 1], r_lo = right[2 * g], r_hi = right[2 * g + 1];
        Fr p0 = l_lo.mul(r_lo);
        if (cz) p0 = p0.sub(cz[2 * g]);
        Fr slope = l_hi.sub(l_lo).mul(r_hi.sub(r_lo));
        i0 = i0.add(p0.mul(Ei[xi]));
        ii = ii.add(slope.mul(Ei[xi]));
      }
    }
    t0 = t0.add(i0.mul(Eo[xo]));
    tinf = tinf.add(ii.mul(Eo[xo]));
  }
  memcpy(out, t0.l, 32);
  memcpy(out + 4, tinf.l, 32);
}

// Keccak transcript
void ref_transcript_init(void* t, const uint8_t* domain, size_t n) {
  Transcript* tr = static_cast<Transcript*>(t);
  memset(tr, 0, sizeof(Transcript));
  tr_append(tr, domain, n);
}
size_t ref_transcript_size(void) { return sizeof(Transcript); }
void ref_transcript_append_bytes(void* t, const uint8_t* data, size_t n) { tr_append(static_cast<Transcript*>(t), data, n); }
void ref_transcript_append_scalar(void* t, const uint8_t* label, size_t n, const uint64_t* scalar) {
  tr_append(static_cast<Transcript*>(t), label, n);
  tr_append(static_cast<Transcript*>(t), reinterpret_cast<const uint8_t*>(scalar), 32);
}
void ref_transcript_challenge_scalar(void* t, const uint8_t* label, size_t n, uint64_t* out) {
  Transcript* tr = static_cast<Transcript*>(t);
  tr_append(tr, label, n);
  keccak_f(tr->state);
  Fr raw;
  memcpy(raw.l, tr->state, 32);
  Fr m = raw.to_mont();  // fromBytes (src/field/mod.zig:625-639)
  memcpy(out, m.l, 32);
}

// Whole dense sumcheck "as the reference runs it": per round compute evals (ref_sc_round), absorb
// them, squeeze the challenge, bind every table (2-mul fold for kinds 1-2 as in val_*.zig, 1-mul
// bindLow form for kinds 3-4 as in output_check.zig, bindFirst for half-split).  Used as the timed
// CPU baseline.  tables are modified in place; evals_out receives rounds*4 elements.
void ref_sc_prove(int kind, int pairing, uint64_t* const* tables, int k, size_t n, const uint8_t* domain, size_t dn,
                  uint64_t* evals_out, uint64_t* challenges_out) {
  Transcript tr;
  ref_transcript_init(&tr, domain, dn);
  std::vector<uint64_t> tmp;
  int rounds = 0;
  while ((size_t(1) << rounds) < n) rounds++;
  for (int rnd = 0; rnd < rounds; rnd++) {
    uint64_t ev[16];
    ref_sc_round(kind, pairing, tables, n, ev);
    const int ne = kind == 0 ? 2 : (kind == 1 ? 3 : 4);  // degree + 1 evaluations are absorbed
    for (int j = 0; j < ne; j++) {
      char label[32];
      int ln = snprintf(label, sizeof label, "round_poly_%d", j);
      ref_transcript_append_scalar(&tr, reinterpret_cast<const uint8_t*>(label), (size_t)ln, ev + 4 * j);
    }
    memcpy(evals_out + (size_t)rnd * 16, ev, 128);
    uint64_t r[4];
    ref_transcript_challenge_scalar(&tr, reinterpret_cast<const uint8_t*>("round"), 5, r);
    memcpy(challenges_out + (size_t)rnd * 4, r, 32);
    for (int t = 0; t < k; t++) {
      if (pairing == 0) {
        if (kind == 1 || kind == 2) ref_bind_adjacent_2mul(tables[t], n, r);
        else ref_bind_low(tables[t], n, r);
      } else {
        tmp.resize(n / 2 * 4);
        ref_bind_first(tables[t], n, r, tmp.data());
        memcpy(tables[t], tmp.data(), n / 2 * 32);
      }
    }
    n /= 2;
  }
}

}  // extern "C"
