// C++ host mirror of the reference's interfaces for the hot path, above the C ABI (include/zolt_gpu.h).
//
// The reference is Zig; no Zig toolchain exists in this image, so the host side that a
// `zig/zolt_gpu.zig` shim would provide is written in C++ with the reference's names, argument
// meaning and error behaviour (INTEGRATION.md shows the Zig binding).  Paths are relative to
// /root/reference:
//   AffinePoint / MSM.compute            src/msm/mod.zig:15-140, :355-372   (infallible: panics on error)
//   HyperKZG.setup / commit / batchCommit src/poly/commitment/mod.zig:174-255, :558-570
//   DensePolynomial bindLow / bindFirst  src/poly/mod.zig:128-175
//   EqPolynomial.evals                   src/poly/mod.zig:240-290
//   UniPoly interpolateDegree3 / ...     src/poly/mod.zig:608-685
//   GruenSplitEqPolynomial               src/poly/split_eq.zig:63-434 (scalar state + cubic on host,
//                                        E tables and the O(N) sums on the device)
//   SumcheckInstance                     src/zkvm/batched_sumcheck.zig:34-74
//   stage round loop with transcript     src/zkvm/prover.zig:405-428
// Everything O(N) is a CUDA kernel; the host keeps only O(1)-per-round field glue and Keccak.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/zolt_gpu.h"
#include "fr_host.hpp"
#include "transcript.hpp"

namespace zolt::host {

// MSM.compute / HyperKZG.commit have no error channel in the reference (no error union): a failing
// device call is a panic, exactly what the Zig shim does with @panic.
[[noreturn]] inline void panic(const char* what, int status) {
  std::fprintf(stderr, "zolt_gpu panic: %s failed with status %d (no CPU fallback)\n", what, status);
  std::abort();
}
inline void must(int status, const char* what) {
  if (status != ZOLT_OK) panic(what, status);
}

class Context {
 public:
  explicit Context(int device = 0, void* stream = nullptr) { must(zolt_gpu_init(device, stream, &ctx_), "zolt_gpu_init"); }
  explicit Context(zolt_ctx* borrowed) : ctx_(borrowed), owned_(false) {}
  ~Context() {
    if (owned_) zolt_gpu_shutdown(ctx_);
  }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  zolt_ctx* raw() const { return ctx_; }

 private:
  zolt_ctx* ctx_ = nullptr;
  bool owned_ = true;
};

// AffinePoint(Fq): x, y in Montgomery form + infinity flag; identity = (0, 0, true) (msm/mod.zig:24-30)
struct AffinePoint {
  uint64_t x[4] = {0, 0, 0, 0};
  uint64_t y[4] = {0, 0, 0, 0};
  bool infinity = true;
  static AffinePoint identity() { return AffinePoint{}; }
  bool eql(const AffinePoint& o) const {
    if (infinity || o.infinity) return infinity == o.infinity;
    return std::memcmp(x, o.x, 32) == 0 && std::memcmp(y, o.y, 32) == 0;
  }
};

// Device-resident bases: SetupParams.powers_of_tau_g1 uploaded once (commitment/mod.zig:122-140).
class DeviceBases {
 public:
  DeviceBases(Context& ctx, const std::vector<AffinePoint>& pts) : ctx_(ctx) {
    std::vector<uint64_t> xy(pts.size() * 8);
    std::vector<uint8_t> inf(pts.size());
    for (size_t i = 0; i < pts.size(); i++) {
      std::memcpy(&xy[8 * i], pts[i].x, 32);
      std::memcpy(&xy[8 * i + 4], pts[i].y, 32);
      inf[i] = pts[i].infinity ? 1 : 0;
    }
    must(zolt_msm_bases_upload(ctx.raw(), xy.data(), inf.data(), pts.size(), &h_), "zolt_msm_bases_upload");
  }
  DeviceBases(Context& ctx, zolt_bases* h) : ctx_(ctx), h_(h) {}
  ~DeviceBases() { zolt_msm_bases_free(h_); }
  DeviceBases(const DeviceBases&) = delete;
  size_t len() const {
    size_t n = 0;
    must(zolt_msm_bases_len(h_, &n), "zolt_msm_bases_len");
    return n;
  }
  zolt_bases* raw() const { return h_; }
  Context& ctx() const { return ctx_; }

 private:
  Context& ctx_;
  zolt_bases* h_ = nullptr;
};

// msm.MSM(F, Fp)
struct MSM {
  // compute(bases, scalars): sum_i scalars[i] * bases[i]; lengths must match (assert, msm/mod.zig:359)
  static AffinePoint compute(const DeviceBases& bases, const std::vector<Fr>& scalars, size_t offset = 0) {
    AffinePoint out;
    uint64_t xy[8];
    uint8_t inf = 0;
    must(zolt_msm(bases.ctx().raw(), bases.raw(), offset, reinterpret_cast<const uint64_t*>(scalars.data()),
                  scalars.size(), xy, &inf),
         "zolt_msm");
    std::memcpy(out.x, xy, 32);
    std::memcpy(out.y, xy + 4, 32);
    out.infinity = inf != 0;
    return out;
  }
};

// HyperKZG(F): commit path only (open/verify are outside this tier)
struct HyperKZG {
  struct SetupParams {
    std::unique_ptr<DeviceBases> powers_of_tau_g1;
  };
  // setup(max_degree): mock SRS tau = 0x12345678, powers[i] = tau^i * G, generated on the device
  static SetupParams setup(Context& ctx, size_t max_degree) {
    zolt_bases* h = nullptr;
    must(zolt_hyperkzg_setup(ctx.raw(), max_degree, &h), "zolt_hyperkzg_setup");
    SetupParams p;
    p.powers_of_tau_g1 = std::make_unique<DeviceBases>(ctx, h);
    return p;
  }
  // commit(params, evals): n = min(len(evals), len(powers)); empty -> identity (commitment/mod.zig:239-255)
  static AffinePoint commit(const SetupParams& params, const std::vector<Fr>& evals) {
    const size_t n = std::min(evals.size(), params.powers_of_tau_g1->len());
    if (n == 0) return AffinePoint::identity();
    std::vector<Fr> head(evals.begin(), evals.begin() + n);
    return MSM::compute(*params.powers_of_tau_g1, head);
  }
  // open(params, evals, point, value) (commitment/mod.zig:261-324): quotient commitments per variable + final_eval
  struct Proof {
    std::vector<AffinePoint> quotient_commitments;
    Fr final_eval;
  };
  static Proof open(const SetupParams& params, const std::vector<Fr>& evals, const std::vector<Fr>& point, const Fr& value) {
    Proof pr;
    pr.final_eval = value;
    const int nv = (int)point.size();
    if (nv == 0) return pr;
    std::vector<uint64_t> xy((size_t)nv * 8);
    std::vector<uint8_t> inf((size_t)nv);
    int nq = 0;
    uint64_t fin[4];
    must(zolt_hyperkzg_open(params.powers_of_tau_g1->ctx().raw(), params.powers_of_tau_g1->raw(),
                            reinterpret_cast<const uint64_t*>(evals.data()), evals.size(), reinterpret_cast<const uint64_t*>(point.data()),
                            nv, xy.data(), inf.data(), &nq, fin),
         "zolt_hyperkzg_open");
    for (int i = 0; i < nq; i++) {
      AffinePoint p;
      std::memcpy(p.x, &xy[8 * i], 32);
      std::memcpy(p.y, &xy[8 * i + 4], 32);
      p.infinity = inf[i] != 0;
      pr.quotient_commitments.push_back(p);
    }
    pr.final_eval = Fr::fromRaw(fin);
    return pr;
  }
  static std::vector<AffinePoint> batchCommit(const SetupParams& params, const std::vector<std::vector<Fr>>& polys) {
    std::vector<AffinePoint> out;
    for (const auto& p : polys) out.push_back(commit(params, p));
    return out;
  }
};

// SRS containers (src/poly/commitment/srs.zig, dory.zig): host glue that walks the headers and hands the G1 array to the
// device, where the points are converted, checked and kept resident (zolt_msm_bases_upload_wire).  G2 / pairing data is
// located and skipped: it belongs to the CPU side of the verifier.
namespace srs {
enum class Error { Ok = 0, TruncatedData, InvalidFileFormat, UnsupportedFormat, PointNotOnCurve, Device };
struct G1Srs {
  zolt_bases* powers = nullptr;  // caller frees with zolt_msm_bases_free
  size_t count = 0;
  uint64_t max_num_vars = 0;     // Dory SRS only
  size_t bad_index = (size_t)-1;
};
inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
inline uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | (uint64_t)rd32(p + 4) << 32; }
inline Error upload(Context& ctx, const uint8_t* pts, size_t n, int fmt, G1Srs* out) {
  const int st = zolt_msm_bases_upload_wire(ctx.raw(), pts, n, fmt, &out->powers, &out->bad_index);
  out->count = n;
  if (st == ZOLT_OK) return Error::Ok;
  return out->bad_index != (size_t)-1 ? Error::PointNotOnCurve : Error::Device;
}
// loadFromRawBinary (srs.zig:256-306): u32 count, count * 64-byte big-endian points, then tau*G2 (128), G1 generator (64),
// G2 generator (128)
inline Error loadFromRawBinary(Context& ctx, const uint8_t* data, size_t len, G1Srs* out) {
  if (len < 4) return Error::TruncatedData;
  const size_t n = rd32(data);
  if (len < 4 + n * 64 + 128 + 64 + 128) return Error::TruncatedData;
  return upload(ctx, data + 4, n, ZOLT_G1_WIRE_BE, out);
}
// loadFromPtau (srs.zig:733-900): "ptau", version 1, section table; header section: field size 32, prime, power; the tauG1
// section holds min(2^(power+1) - 1, size / 64) little-endian points
inline Error loadFromPtau(Context& ctx, const uint8_t* data, size_t len, G1Srs* out) {
  if (len < 12) return Error::TruncatedData;
  if (std::memcmp(data, "ptau", 4) != 0) return Error::InvalidFileFormat;
  if (rd32(data + 4) != 1) return Error::UnsupportedFormat;
  const uint32_t sections = rd32(data + 8);
  size_t off = 12, hdr_off = 0, hdr_size = 0, g1_off = 0, g1_size = 0;
  bool have_hdr = false, have_g1 = false;
  for (uint32_t i = 0; i < sections; i++) {
    if (off + 12 > len) return Error::TruncatedData;
    const uint32_t type = rd32(data + off);
    const uint64_t size = rd64(data + off + 4);
    off += 12;
    if (type == 1) { hdr_off = off; hdr_size = (size_t)size; have_hdr = true; }
    if (type == 2) { g1_off = off; g1_size = (size_t)size; have_g1 = true; }
    off += (size_t)size;
  }
  if (!have_hdr) return Error::InvalidFileFormat;
  if (hdr_off + hdr_size > len || hdr_size < 8) return Error::TruncatedData;
  if (rd32(data + hdr_off) != 32) return Error::UnsupportedFormat;
  if (hdr_size < 4 + 32 + 4 + 4) return Error::TruncatedData;
  const uint32_t power = rd32(data + hdr_off + 36);
  size_t n = 0;
  if (have_g1) {
    if (g1_off + g1_size > len) return Error::TruncatedData;
    const size_t want = (size_t(1) << power) * 2 - 1;
    n = want < g1_size / 64 ? want : g1_size / 64;
  }
  return upload(ctx, data + g1_off, n, ZOLT_G1_WIRE_LE, out);
}
// DoryCommitmentScheme.loadFromFile (dory.zig:740-826): "JOLT_DORY_SRS_V1", u64 max_num_vars, u64 g1_count, the G1 vector
// (arkworks uncompressed), u64 g2_count, the G2 vector, the blinding generators
inline Error loadDorySrs(Context& ctx, const uint8_t* data, size_t len, G1Srs* out) {
  if (len < 32) return Error::TruncatedData;
  if (std::memcmp(data, "JOLT_DORY_SRS_V1", 16) != 0) return Error::InvalidFileFormat;
  out->max_num_vars = rd64(data + 16);
  const size_t n = (size_t)rd64(data + 24);
  if (len < 32 + n * 64) return Error::TruncatedData;
  return upload(ctx, data + 32, n, ZOLT_G1_WIRE_ARK_LE, out);
}
}  // namespace srs

// UniPoly(F) helpers used around the round loop (src/poly/mod.zig:608-685)
struct UniPoly {
  static Fr evaluate(const std::vector<Fr>& coeffs, const Fr& x) {
    Fr acc = Fr::zero();
    for (size_t i = coeffs.size(); i-- > 0;) acc = acc.mul(x).add(coeffs[i]);
    return acc;
  }
  static std::array<Fr, 4> interpolateDegree3(const std::array<Fr, 4>& e) {
    static const Fr inv6 = Fr::fromU64(6).inverse(), inv2 = Fr::fromU64(2).inverse();  // once, not per call as the reference does
    auto k = [](uint64_t v) { return Fr::fromU64(v); };
    Fr c1 = Fr::zero().sub(k(11).mul(e[0])).add(k(18).mul(e[1])).sub(k(9).mul(e[2])).add(k(2).mul(e[3])).mul(inv6);
    Fr c2 = k(2).mul(e[0]).sub(k(5).mul(e[1])).add(k(4).mul(e[2])).sub(e[3]).mul(inv2);
    Fr c3 = Fr::zero().sub(e[0]).add(k(3).mul(e[1])).sub(k(3).mul(e[2])).add(e[3]).mul(inv6);
    return {e[0], c1, c2, c3};
  }
  static std::array<Fr, 3> evalsToCompressed(const std::array<Fr, 4>& e) {
    auto c = interpolateDegree3(e);
    return {c[0], c[2], c[3]};
  }
};

// evaluateCubicAtPointGeneric (src/zkvm/batched_sumcheck.zig:437-457)
inline Fr evaluateCubicAtPoint(const std::array<Fr, 4>& e, const Fr& x) {
  // the reference inverts 6, 2, -2, -6 on every call (four Fermat inversions per round); same values, computed once
  static const Fr one = Fr::one(), two = Fr::fromU64(2), three = Fr::fromU64(3);
  static const Fr inv_m6 = Fr::fromU64(6).neg().inverse(), inv_2 = Fr::fromU64(2).inverse(), inv_m2 = Fr::fromU64(2).neg().inverse(),
                  inv_6 = Fr::fromU64(6).inverse();
  const Fr x1 = x.sub(one), x2 = x.sub(two), x3 = x.sub(three);
  const Fr L0 = x1.mul(x2).mul(x3).mul(inv_m6);
  const Fr L1 = x.mul(x2).mul(x3).mul(inv_2);
  const Fr L2 = x.mul(x1).mul(x3).mul(inv_m2);
  const Fr L3 = x.mul(x1).mul(x2).mul(inv_6);
  return e[0].mul(L0).add(e[1].mul(L1)).add(e[2].mul(L2)).add(e[3].mul(L3));
}

// DensePolynomial(F) with the evaluation table on the device
class DensePolynomial {
 public:
  DensePolynomial(Context& ctx, const std::vector<Fr>& evaluations) : ctx_(ctx), len_(evaluations.size()) {
    if (len_ == 0 || (len_ & (len_ - 1))) panic("DensePolynomial.init: length must be a power of two", ZOLT_ERR_ARG);
    must(zolt_dev_alloc(ctx.raw(), len_ * 32, &d_), "zolt_dev_alloc");
    must(zolt_dev_upload(ctx.raw(), d_, evaluations.data(), len_ * 32), "zolt_dev_upload");
  }
  ~DensePolynomial() { zolt_dev_free(ctx_.raw(), d_); }
  DensePolynomial(const DensePolynomial&) = delete;
  size_t boundLen() const { return len_; }
  size_t numVars() const {
    size_t v = 0;
    while ((size_t(1) << v) < len_) v++;
    return v;
  }
  void bindLow(const Fr& r) {  // in place, src/poly/mod.zig:160-175
    must(zolt_poly_bind_low(ctx_.raw(), d_, d_, len_, r.limbs.data()), "zolt_poly_bind_low");
    len_ /= 2;
  }
  void bindFirst(const Fr& r) {  // src/poly/mod.zig:128-149 (the table is reused instead of reallocated)
    must(zolt_poly_bind_first(ctx_.raw(), d_, d_, len_, r.limbs.data()), "zolt_poly_bind_first");
    len_ /= 2;
  }
  Fr sum() const {
    Fr out;
    must(zolt_poly_sum(ctx_.raw(), d_, len_, out.limbs.data()), "zolt_poly_sum");
    return out;
  }
  std::vector<Fr> evaluations() const {
    std::vector<Fr> out(len_);
    must(zolt_dev_download(ctx_.raw(), out.data(), d_, len_ * 32), "zolt_dev_download");
    return out;
  }
  const void* devicePtr() const { return d_; }
  // evaluate(point) (src/poly/mod.zig:73-92): the multilinear extension at `point`, point[j] <-> variable j (index bit j).
  // The reference folds one variable at a time on a copy; here the same folds run on the device (bindLow per variable on
  // a scratch copy), the table itself is left untouched.
  Fr evaluate(const std::vector<Fr>& point) const {
    if (point.size() != numVars()) panic("DensePolynomial.evaluate: point length != num_vars", ZOLT_ERR_ARG);
    if (len_ == 1) return evaluations()[0];
    void* tmp = nullptr;
    must(zolt_dev_alloc(ctx_.raw(), len_ * 32, &tmp), "zolt_dev_alloc");
    must(zolt_dev_copy(ctx_.raw(), tmp, d_, len_ * 32), "zolt_dev_copy");
    size_t n = len_;
    for (size_t j = 0; j < point.size(); j++, n /= 2) must(zolt_poly_bind_low(ctx_.raw(), tmp, tmp, n, point[j].limbs.data()), "zolt_poly_bind_low");
    Fr out;
    must(zolt_dev_download(ctx_.raw(), out.limbs.data(), tmp, 32), "zolt_dev_download");
    zolt_dev_free(ctx_.raw(), tmp);
    return out;
  }

 private:
  Context& ctx_;
  size_t len_;
  void* d_ = nullptr;
};

// BatchOps.hornerEval (src/field/mod.zig:1164-1280): sum_i coeffs[i] x^i, O(n) host glue (Horner from the top coefficient)
struct BatchOps {
  static Fr hornerEval(const std::vector<Fr>& coeffs, const Fr& x) {
    Fr acc = Fr::zero();
    for (size_t i = coeffs.size(); i-- > 0;) acc = acc.mul(x).add(coeffs[i]);
    return acc;
  }
};

struct EqPolynomial {
  // mle(r, x) (src/poly/mod.zig:311-321): prod_i (r_i x_i + (1 - r_i)(1 - x_i)); O(n) host glue
  static Fr mle(const std::vector<Fr>& r, const std::vector<Fr>& x) {
    if (r.size() != x.size()) panic("EqPolynomial.mle: length mismatch", ZOLT_ERR_ARG);
    Fr acc = Fr::one();
    for (size_t i = 0; i < r.size(); i++) acc = acc.mul(r[i].mul(x[i]).add(Fr::one().sub(r[i]).mul(Fr::one().sub(x[i]))));
    return acc;
  }
  // evals(): big-endian table eq(r, .) (src/poly/mod.zig:240-290), computed on the device
  static std::vector<Fr> evals(Context& ctx, const std::vector<Fr>& r, const Fr* scale = nullptr) {
    const size_t n = size_t(1) << r.size();
    void* d = nullptr;
    must(zolt_dev_alloc(ctx.raw(), n * 32, &d), "zolt_dev_alloc");
    must(zolt_eq_evals(ctx.raw(), reinterpret_cast<const uint64_t*>(r.data()), (int)r.size(),
                       scale ? scale->limbs.data() : nullptr, d),
         "zolt_eq_evals");
    std::vector<Fr> out(n);
    must(zolt_dev_download(ctx.raw(), out.data(), d, n * 32), "zolt_dev_download");
    zolt_dev_free(ctx.raw(), d);
    return out;
  }
};

// Host part of GruenSplitEqPolynomial: the running scalar, the bound-variable counter and the cubic
// assembly (src/poly/split_eq.zig:213-248, :353-434).  The prefix tables live on the device.
class GruenSplitEqPolynomial {
  std::vector<Fr> tau_;  // declared first: current_index is initialised from it

 public:
  explicit GruenSplitEqPolynomial(std::vector<Fr> tau, const Fr* scaling = nullptr)
      : tau_(std::move(tau)), current_index(tau_.size()), current_scalar(scaling ? *scaling : Fr::one()) {}
  void bind(const Fr& r) {
    if (current_index == 0) return;
    const Fr& t = tau_[current_index - 1];
    const Fr eq = t.mul(r).add(Fr::one().sub(t).mul(Fr::one().sub(r)));
    current_scalar = current_scalar.mul(eq);
    current_index--;
  }
  std::array<Fr, 4> computeCubicRoundPoly(const Fr& q_constant, const Fr& q_quadratic_coeff, const Fr& previous_claim) const {
    if (current_index == 0) return {previous_claim, Fr::zero(), Fr::zero(), Fr::zero()};
    const Fr& t = tau_[current_index - 1];
    const Fr eq0 = current_scalar.mul(Fr::one().sub(t)), eq1 = current_scalar.mul(t);
    const Fr slope = eq1.sub(eq0);
    const Fr l2 = eq0.add(slope.mul(Fr::fromU64(2))), l3 = eq0.add(slope.mul(Fr::fromU64(3)));
    const Fr q1 = eq1.isZero() ? Fr::zero() : previous_claim.sub(eq0.mul(q_constant)).mul(eq1.inverse());
    const Fr e2 = q_quadratic_coeff.add(q_quadratic_coeff);
    const Fr q2 = q1.add(q1).sub(q_constant).add(e2);
    const Fr q3 = q2.add(q1).sub(q_constant).add(e2).add(e2);
    return {eq0.mul(q_constant), eq1.mul(q1), l2.mul(q2), l3.mul(q3)};
  }
  const std::vector<Fr>& tau() const { return tau_; }
  size_t current_index;
  Fr current_scalar;
};

// SumcheckInstance (src/zkvm/batched_sumcheck.zig:34-74): evaluations at X = 0,1,2,3, strictly
// alternating computeRoundPoly -> bindChallenge.
class SumcheckInstance {
 public:
  virtual ~SumcheckInstance() = default;
  virtual size_t numRounds() const = 0;
  virtual size_t degree() const = 0;
  virtual std::array<Fr, 4> computeRoundPoly(size_t round) = 0;
  virtual void bindChallenge(const Fr& challenge) = 0;
};

// k dense tables combined by one of the ZOLT_SC_* combiners; covers ValEvaluationProver,
// ValFinalProver, OutputSumcheckProver, JoltSpartanInterface and the config-3 Spartan shape.
class DenseSumcheckProver : public SumcheckInstance {
 public:
  DenseSumcheckProver(Context& ctx, int combiner, int pairing, const std::vector<const uint64_t*>& tables, int log_n,
                      int n_extra = 0)
      : log_n_(log_n), combiner_(combiner) {
    must(zolt_sc_create(ctx.raw(), combiner, pairing, log_n, (int)tables.size() - n_extra, n_extra, tables.data(), &sc_),
         "zolt_sc_create");
    must(zolt_sc_num_evals(sc_, &ne_), "zolt_sc_num_evals");
  }
  // tables already resident on the device (borrowed read-only by the instance)
  struct DeviceTables {};
  DenseSumcheckProver(Context& ctx, DeviceTables, int combiner, int pairing, const std::vector<const void*>& d_tables,
                      int log_n, int n_extra = 0)
      : log_n_(log_n), combiner_(combiner) {
    must(zolt_sc_create_dev(ctx.raw(), combiner, pairing, log_n, (int)d_tables.size() - n_extra, n_extra, d_tables.data(), &sc_),
         "zolt_sc_create_dev");
    must(zolt_sc_num_evals(sc_, &ne_), "zolt_sc_num_evals");
  }
  ~DenseSumcheckProver() override { zolt_sc_free(sc_); }
  size_t numRounds() const override { return (size_t)log_n_; }
  size_t degree() const override {
    return combiner_ == ZOLT_SC_SUM ? 1 : (combiner_ == ZOLT_SC_PROD2 ? 2 : 3);
  }
  std::array<Fr, 4> computeRoundPoly(size_t) override {
    std::array<Fr, 4> e{};
    uint64_t buf[16] = {0};
    must(zolt_sc_round_evals(sc_, buf), "zolt_sc_round_evals");
    for (int i = 0; i < ne_; i++) e[i] = Fr::fromRaw(buf + 4 * i);
    if (combiner_ == ZOLT_SC_SUM) {
      // linear round polynomial: p(2) = 2 p(1) - p(0), p(3) = 3 p(1) - 2 p(0) (jolt_r1cs.zig:441-444)
      const Fr d = e[1].sub(e[0]);
      e[2] = e[1].add(d);
      e[3] = e[2].add(d);
    }
    return e;
  }
  void bindChallenge(const Fr& r) override { must(zolt_sc_bind(sc_, r.limbs.data()), "zolt_sc_bind"); }
  // SumcheckInstance.input_claim (src/zkvm/batched_sumcheck.zig:34-74): with it the first round also takes p(1) from
  // p(0) + p(1) = claim instead of summing it (one accumulator less in the round-0 kernel); same evaluations
  void setInputClaim(const Fr& claim) { must(zolt_sc_set_claim(sc_, claim.limbs.data()), "zolt_sc_set_claim"); }
  std::vector<Fr> table(int t) {
    size_t n = 0;
    must(zolt_sc_len(sc_, &n), "zolt_sc_len");
    std::vector<Fr> out(n);
    must(zolt_sc_download(sc_, t, reinterpret_cast<uint64_t*>(out.data())), "zolt_sc_download");
    return out;
  }
  zolt_sc* raw() const { return sc_; }

 private:
  zolt_sc* sc_ = nullptr;
  int log_n_, combiner_, ne_ = 0;
};

// ProductVirtualRemainderProver / StreamingOuterProver linear phase
// (src/zkvm/spartan/product_remainder.zig:269-378): Gruen split-eq degree 3.
class GruenProductProver : public SumcheckInstance {
 public:
  GruenProductProver(Context& ctx, const std::vector<const uint64_t*>& tables, int log_n, const std::vector<Fr>& tau,
                     const Fr& initial_claim, bool with_c = false)
      : log_n_(log_n), split_eq_(tau), current_claim(initial_claim) {
    must(zolt_sc_create(ctx.raw(), with_c ? ZOLT_SC_GRUEN_AB_MINUS_C : ZOLT_SC_GRUEN_AB, ZOLT_PAIR_ADJACENT, log_n,
                        (int)tables.size(), 0, tables.data(), &sc_),
         "zolt_sc_create");
    must(zolt_sc_set_split_eq(sc_, reinterpret_cast<const uint64_t*>(tau.data()), (int)tau.size()), "zolt_sc_set_split_eq");
  }
  // tables already resident on the device (borrowed read-only by the instance)
  GruenProductProver(Context& ctx, DenseSumcheckProver::DeviceTables, const std::vector<const void*>& d_tables, int log_n,
                     const std::vector<Fr>& tau, const Fr& initial_claim, bool with_c = false)
      : log_n_(log_n), split_eq_(tau), current_claim(initial_claim) {
    must(zolt_sc_create_dev(ctx.raw(), with_c ? ZOLT_SC_GRUEN_AB_MINUS_C : ZOLT_SC_GRUEN_AB, ZOLT_PAIR_ADJACENT, log_n,
                            (int)d_tables.size(), 0, d_tables.data(), &sc_),
         "zolt_sc_create_dev");
    must(zolt_sc_set_split_eq(sc_, reinterpret_cast<const uint64_t*>(tau.data()), (int)tau.size()), "zolt_sc_set_split_eq");
  }
  ~GruenProductProver() override { zolt_sc_free(sc_); }
  size_t numRounds() const override { return (size_t)log_n_; }
  size_t degree() const override { return 3; }
  std::array<Fr, 4> computeRoundPoly(size_t) override {
    uint64_t buf[8];
    must(zolt_sc_round_evals(sc_, buf), "zolt_sc_round_evals");
    last_ = split_eq_.computeCubicRoundPoly(Fr::fromRaw(buf), Fr::fromRaw(buf + 4), current_claim);
    return last_;
  }
  void bindChallenge(const Fr& r) override {
    current_claim = evaluateCubicAtPoint(last_, r);  // updateClaim (product_remainder.zig:381-384)
    must(zolt_sc_bind(sc_, r.limbs.data()), "zolt_sc_bind");
    split_eq_.bind(r);
  }
  zolt_sc* raw() const { return sc_; }
  Fr current_claim;

 private:
  zolt_sc* sc_ = nullptr;
  int log_n_;
  GruenSplitEqPolynomial split_eq_;
  std::array<Fr, 4> last_{};
};

// Stage4GruenProver (src/zkvm/spartan/stage4_gruen_prover.zig): registers read/write checking,
//   sum_{k,j} eq(r_cycle, j) * [ ra(k,j) val(k,j) + wa(k,j) (val(k,j) + inc(j)) ],   ra = gamma rs1_ra + gamma^2 rs2_ra,
// over K = 128 registers x T cycles, LOG_K + log T rounds in three phases (:538-556): cycle variables through the Gruen
// split eq, the 7 address variables with the eq not bound, the remaining cycle variables against the merged dense eq.
// The prover is entered with the dense tables initWithPhaseConfig (:139-366) leaves (witness generation from the
// ExecutionTrace is the caller's side of the boundary); everything O(K T) runs on the device:
//   phase 1   ZOLT_SC_STAGE4_GRUEN over the flat [k][j] tables (rows compact, so the reference's pairs (k T + 2i, k T + 2i + 1)
//             are the flat adjacent pairs), inc replicated over k, weights E_out E_in over the cycle bits; the cubic is
//             gruenPolyDeg3 (gruen_eq.zig:214-262) on the host, O(1);
//   phases 2, 3   ONE ZOLT_SC_STAGE4_DENSE instance over the tables transposed to [j][k], merged eq (gruen_eq.zig:172-203)
//             and inc repeated over k: the k bit is now the lowest, address rounds bind eq and inc to themselves, and once
//             k is exhausted the same flat pairs are the cycle pairs of phase 3.
// Supported configurations are the ones the reference completes: 1 <= phase1 <= log T (merged_eq is only created inside a
// phase-1 bind, :1083-1086) and phase2 == LOG_K (bindPolynomials' third branch never binds an address variable,
// :1123-1158).  computeRoundEvals / bindChallenge / getFinalClaims keep the reference's signatures (:1165-1234).
class Stage4GruenProver {
 public:
  static constexpr int LOG_K = 7;
  static constexpr size_t K = size_t(1) << LOG_K;
  struct Stage3Claims {
    Fr rd_write_value, rs1_value, rs2_value;
  };
  struct FinalClaims {
    Fr val_claim, rs1_ra_claim, rs2_ra_claim, rd_wa_claim, inc_claim;
  };

  // d_val, d_rd_wa, d_rs1_ra, d_rs2_ra: K * T elements each, [k * T + j]; d_inc: T elements; all device-resident and only
  // read.  r_cycle in round order (little-endian), like the reference's argument.
  Stage4GruenProver(Context& ctx, int log_T, const void* d_val, const void* d_rd_wa, const void* d_rs1_ra, const void* d_rs2_ra,
                    const void* d_inc, const Fr& gamma, const std::vector<Fr>& r_cycle, const Stage3Claims* stage3_claims,
                    const Fr& batching_coeff, int phase1_num_rounds = -1, int phase2_num_rounds = LOG_K)
      : ctx_(ctx), log_T_(log_T), gamma_(gamma), batching_coeff_(batching_coeff),
        p1_(phase1_num_rounds < 0 ? log_T / 2 : phase1_num_rounds), p2_(phase2_num_rounds) {
    if (log_T < 1 || (int)r_cycle.size() != log_T || p1_ < 1 || p1_ > log_T || p2_ != LOG_K) panic("Stage4GruenProver: unsupported configuration", ZOLT_ERR_ARG);
    if (stage3_claims) {
      have_claims_ = true;
      claims_ = *stage3_claims;
    }
    w_.assign(r_cycle.rbegin(), r_cycle.rend());  // big-endian copy for the split eq (:259-281)
    current_index_ = w_.size();
    current_scalar_ = Fr::one();
    const size_t T = size_t(1) << log_T, n = K * T;
    d_in_[0] = d_val;
    d_in_[1] = d_rd_wa;
    d_in_[2] = d_rs1_ra;
    d_in_[3] = d_rs2_ra;
    d_inc_in_ = d_inc;
    // ra = gamma rs1_ra + gamma^2 rs2_ra (:211, :224); inc replicated over the rows
    d_ra_ = alloc(n);
    const Fr g2[2] = {gamma, gamma.mul(gamma)};
    must(zolt_fr_vec_dev(ctx.raw(), ZOLT_VEC_LIN2, d_ra_, d_rs1_ra, d_rs2_ra, nullptr, nullptr, reinterpret_cast<const uint64_t*>(g2), n),
         "zolt_fr_vec_dev");
    d_inc_rep_ = alloc(n);
    must(zolt_fr_index_map_dev(ctx.raw(), d_inc_rep_, d_inc, n, 1, T, 1, 0, 0, 0), "zolt_fr_index_map_dev");
    const void* tabs[6] = {d_ra_, d_rd_wa, d_val, d_inc_rep_, d_rs1_ra, d_rs2_ra};
    must(zolt_sc_create_dev(ctx.raw(), ZOLT_SC_STAGE4_GRUEN, ZOLT_PAIR_ADJACENT, LOG_K + log_T, 4, 2, tabs, &sc1_), "zolt_sc_create_dev");
    must(zolt_sc_set_split_eq(sc1_, reinterpret_cast<const uint64_t*>(w_.data()), log_T), "zolt_sc_set_split_eq");
  }
  ~Stage4GruenProver() {
    if (sc1_) zolt_sc_free(sc1_);
    if (sc2_) zolt_sc_free(sc2_);
    for (auto& b : owned_) zolt_dev_release_cached(ctx_.raw(), b.first, b.second);
  }
  Stage4GruenProver(const Stage4GruenProver&) = delete;
  Stage4GruenProver& operator=(const Stage4GruenProver&) = delete;

  size_t numRounds() const { return (size_t)(LOG_K + log_T_); }

  // computeInputClaim (:385-430): from the Stage 3 claims when given, else summed over the tables
  Fr computeInputClaim() {
    if (have_claims_) return claims_.rd_write_value.add(gamma_.mul(claims_.rs1_value.add(gamma_.mul(claims_.rs2_value))));
    const size_t T = size_t(1) << log_T_, n = K * T;
    void* d_eq = alloc(T);
    must(zolt_eq_evals(ctx_.raw(), reinterpret_cast<const uint64_t*>(w_.data()), log_T_, nullptr, d_eq), "zolt_eq_evals");
    void* d_eq_rep = alloc(n);
    must(zolt_fr_index_map_dev(ctx_.raw(), d_eq_rep, d_eq, n, 1, T, 1, 0, 0, 0), "zolt_fr_index_map_dev");
    void* d_term = alloc(n);
    must(zolt_dev_copy(ctx_.raw(), d_term, d_inc_rep_, n * sizeof(Fr)), "zolt_dev_copy");
    must(zolt_fr_vec_dev(ctx_.raw(), ZOLT_VEC_STAGE4_TERM, d_term, d_ra_, d_in_[0], d_in_[1], d_eq_rep, nullptr, n), "zolt_fr_vec_dev");
    Fr claim;
    must(zolt_poly_sum(ctx_.raw(), d_term, n, claim.limbs.data()), "zolt_poly_sum");
    release(d_term);
    release(d_eq_rep);
    release(d_eq);
    return claim;
  }

  // computeRoundEvals (:1165-1190): the round polynomial's coefficients, handed back as evaluations at 0..3
  std::array<Fr, 4> computeRoundEvals(size_t round, const Fr& current_claim) {
    const std::array<Fr, 4> c = computeRoundPolynomialGruen(round, current_claim);
    auto k = [](uint64_t v) { return Fr::fromU64(v); };
    return {c[0], c[0].add(c[1]).add(c[2]).add(c[3]), c[0].add(c[1].mul(k(2))).add(c[2].mul(k(4))).add(c[3].mul(k(8))),
            c[0].add(c[1].mul(k(3))).add(c[2].mul(k(9))).add(c[3].mul(k(27)))};
  }
  // computeRoundPolynomialGruen (:538-556): coefficients [c0, c1, c2, c3]
  std::array<Fr, 4> computeRoundPolynomialGruen(size_t round, const Fr& previous_claim) {
    uint64_t buf[16] = {0};
    if ((int)round < p1_) {
      must(zolt_sc_round_evals(sc1_, buf), "zolt_sc_round_evals");  // [q(0), q_X2]
      return gruenPolyDeg3(Fr::fromRaw(buf), Fr::fromRaw(buf + 4), previous_claim);
    }
    must(zolt_sc_round_evals(sc2_, buf), "zolt_sc_round_evals");  // [p(0), p(2), p(3)], summed directly like the reference
    const Fr e0 = Fr::fromRaw(buf), e2 = Fr::fromRaw(buf + 4), e3 = Fr::fromRaw(buf + 8);
    const Fr e1 = previous_claim.sub(e0);
    static const Fr two_inv = Fr::fromU64(2).inverse();
    if ((int)round < p1_ + p2_) {
      // phase 2 (:764-851): eq and inc do not depend on the address bit, the polynomial has degree 2
      const Fr c2 = e0.sub(e1.mul(Fr::fromU64(2))).add(e2).mul(two_inv);
      return {e0, e1.sub(e0).sub(c2), c2, Fr::zero()};
    }
    return coeffsFromEvals({e0, e1, e2, e3});  // phase 3 (:854-962)
  }

  // bindChallenge / bindPolynomials (:1047-1160, :1192-1217)
  void bindChallenge(size_t round, const Fr& challenge) {
    if ((int)round < p1_) {
      must(zolt_sc_bind(sc1_, challenge.limbs.data()), "zolt_sc_bind");
      // GruenSplitEqPolynomial.bind (gruen_eq.zig:120-160)
      const Fr& w_i = w_[current_index_ - 1];
      const Fr prod = w_i.mul(challenge);
      current_scalar_ = current_scalar_.mul(Fr::one().sub(w_i).sub(challenge).add(prod).add(prod));
      current_index_--;
      if ((int)round == p1_ - 1) enterPhase2();
      return;
    }
    must(zolt_sc_bind(sc2_, challenge.limbs.data()), "zolt_sc_bind");
  }

  // getFinalClaims (:1219-1234), after the last bind
  FinalClaims getFinalClaims() {
    FinalClaims f;
    // sc2 tables: eq, ra, wa, val, inc, rs1_ra, rs2_ra
    must(zolt_sc_download(sc2_, 3, f.val_claim.limbs.data()), "zolt_sc_download");
    must(zolt_sc_download(sc2_, 5, f.rs1_ra_claim.limbs.data()), "zolt_sc_download");
    must(zolt_sc_download(sc2_, 6, f.rs2_ra_claim.limbs.data()), "zolt_sc_download");
    must(zolt_sc_download(sc2_, 2, f.rd_wa_claim.limbs.data()), "zolt_sc_download");
    must(zolt_sc_download(sc2_, 4, f.inc_claim.limbs.data()), "zolt_sc_download");
    return f;
  }
  // merged_eq[0] after the last bind: eq(r_cycle, challenges) (:1203-1215)
  Fr finalEq() {
    Fr e;
    must(zolt_sc_download(sc2_, 0, e.limbs.data()), "zolt_sc_download");
    return e;
  }

  struct Result {
    std::vector<std::array<Fr, 4>> round_polys;  // UNBATCHED coefficients, like Stage4Result.round_polys
    std::vector<Fr> challenges;
    FinalClaims claims;
  };
  // prove (:433-536) against any transcript with the reference's two calls: appendScalar(F), challengeScalar() -> F
  template <class TranscriptT>
  Result prove(TranscriptT& transcript) {
    Result res;
    Fr current_claim = computeInputClaim().mul(batching_coeff_);
    for (size_t round = 0; round < numRounds(); round++) {
      const std::array<Fr, 4> c = computeRoundPolynomialGruen(round, current_claim);
      transcript.appendScalar(c[0].mul(batching_coeff_));
      transcript.appendScalar(c[2].mul(batching_coeff_));
      transcript.appendScalar(c[3].mul(batching_coeff_));
      const Fr r = transcript.challengeScalar();
      res.challenges.push_back(r);
      current_claim = c[0].add(r.mul(c[1].add(r.mul(c[2].add(r.mul(c[3])))))).mul(batching_coeff_);  // RoundPoly.evaluateAt (:20-27)
      bindChallenge(round, r);
      res.round_polys.push_back(c);
    }
    res.claims = getFinalClaims();
    return res;
  }

  // gruenPolyDeg3 (gruen_eq.zig:214-262) on the host state of the split eq
  std::array<Fr, 4> gruenPolyDeg3(const Fr& q_constant, const Fr& q_quadratic_coeff, const Fr& s_0_plus_s_1) const {
    const Fr eq1 = current_scalar_.mul(w_[current_index_ - 1]);
    const Fr eq0 = current_scalar_.sub(eq1);
    const Fr m = eq1.sub(eq0);
    const Fr eq2 = eq1.add(m), eq3 = eq2.add(m);
    const Fr cubic0 = eq0.mul(q_constant);
    const Fr cubic1 = s_0_plus_s_1.sub(cubic0);
    if (eq1.isZero()) panic("Stage4GruenProver: eq_eval_1 is zero", ZOLT_ERR_ARG);  // the reference panics too (:241)
    const Fr q1 = cubic1.mul(eq1.inverse());
    const Fr e2 = q_quadratic_coeff.add(q_quadratic_coeff);
    const Fr q2 = q1.add(q1).sub(q_constant).add(e2);
    const Fr q3 = q2.add(q1).sub(q_constant).add(e2).add(e2);
    return coeffsFromEvals({cubic0, cubic1, eq2.mul(q2), eq3.mul(q3)});
  }
  // fromEvals / coeffsFromEvals (gruen_eq.zig:265-291, stage4_gruen_prover.zig:1017-1039)
  static std::array<Fr, 4> coeffsFromEvals(const std::array<Fr, 4>& e) {
    static const Fr six_inv = Fr::fromU64(6).inverse(), two_inv = Fr::fromU64(2).inverse();
    auto k = [](uint64_t v) { return Fr::fromU64(v); };
    const Fr c3 = e[0].neg().add(e[1].mul(k(3))).sub(e[2].mul(k(3))).add(e[3]).mul(six_inv);
    const Fr c2 = e[0].mul(k(2)).sub(e[1].mul(k(5))).add(e[2].mul(k(4))).sub(e[3]).mul(two_inv);
    return {e[0], e[1].sub(e[0]).sub(c2).sub(c3), c2, c3};
  }

 private:
  // temporaries come from the context's buffer cache: a second proof of the same shape allocates nothing
  void* alloc(size_t elems) {
    void* p = nullptr;
    const size_t bytes = (elems ? elems : 1) * sizeof(Fr);
    must(zolt_dev_alloc_cached(ctx_.raw(), bytes, &p), "zolt_dev_alloc_cached");
    owned_.emplace_back(p, bytes);
    return p;
  }
  void release(void* p) {
    for (size_t i = 0; i < owned_.size(); i++)
      if (owned_[i].first == p) {
        zolt_dev_release_cached(ctx_.raw(), p, owned_[i].second);
        owned_.erase(owned_.begin() + (long)i);
        return;
      }
  }
  // After the last phase-1 bind: merge the split eq (gruen_eq.zig:172-203), transpose [k][j] -> [j][k], open the dense instance
  void enterPhase2() {
    const int log_Tc = log_T_ - p1_;
    const size_t Tc = size_t(1) << log_Tc, n = K * Tc;
    void* d_eq = alloc(Tc);
    must(zolt_eq_evals(ctx_.raw(), reinterpret_cast<const uint64_t*>(w_.data()), (int)current_index_, current_scalar_.limbs.data(), d_eq),
         "zolt_eq_evals");
    void* t[7];
    for (auto& p : t) p = alloc(n);
    must(zolt_fr_index_map_dev(ctx_.raw(), t[0], d_eq, n, K, Tc, 1, 0, 0, 0), "zolt_fr_index_map_dev");  // eq[j] repeated over k
    // sc1 tables: ra, wa, val, inc_rep, rs1_ra, rs2_ra  ->  sc2 tables: eq, ra, wa, val, inc, rs1_ra, rs2_ra
    static const int from[6] = {0, 1, 2, 3, 4, 5};
    for (int i = 0; i < 6; i++) {
      const void* src = nullptr;
      must(zolt_sc_table_dev(sc1_, from[i], &src), "zolt_sc_table_dev");
      must(zolt_fr_index_map_dev(ctx_.raw(), t[1 + i], src, n, 1, K, Tc, K, Tc, 1), "zolt_fr_index_map_dev");  // dst[j K + k] = src[k Tc + j]
    }
    must(zolt_sc_create_dev(ctx_.raw(), ZOLT_SC_STAGE4_DENSE, ZOLT_PAIR_ADJACENT, LOG_K + log_Tc, 5, 2, t, &sc2_), "zolt_sc_create_dev");
    zolt_sc_free(sc1_);  // joins its kernels first: the copies above are ordered behind them on the context's stream
    sc1_ = nullptr;
  }

  Context& ctx_;
  int log_T_;
  Fr gamma_, batching_coeff_;
  int p1_, p2_;
  bool have_claims_ = false;
  Stage3Claims claims_{};
  std::vector<Fr> w_;
  size_t current_index_ = 0;
  Fr current_scalar_;
  const void* d_in_[4] = {};
  const void* d_inc_in_ = nullptr;
  void* d_ra_ = nullptr;
  void* d_inc_rep_ = nullptr;
  zolt_sc* sc1_ = nullptr;
  zolt_sc* sc2_ = nullptr;
  std::vector<std::pair<void*, size_t>> owned_;
};

// ---- Stage 3 (src/zkvm/spartan/stage3_prover.zig): Shift, InstructionInput and RegistersClaimReduction provers ------------
// The provers are entered with the trace COLUMNS they read out of cycle_witnesses (device tables of 2^log_T elements, only
// read).  Everything O(T) runs on the device:
//   * Q buffers: zolt_fr_fold_high_dev (column folded over the high half of the cycle index against the suffix table);
//   * phase 1 of the prefix-suffix provers: ONE ZOLT_SC_PROD2 instance over the P buffers / the Q buffers laid end to end
//     (adjacent pairs never straddle two buffers, low-to-high binds keep the blocks apart);
//   * the witness columns bound in lock-step by a ZOLT_SC_SUM instance with extra tables -- the reference rebuilds them at
//     the transition as sum_i eq(r_prefix, i) w[j P + i] (:1626-1707), which is the same field element as the chain of binds;
//   * phase 2: ZOLT_SC_PROD2 instances (eq x batched value) with the columns as bound-only extras for finalClaims;
//   * InstructionInputProver: two ZOLT_SC_EQ_DOT2 instances (right and left operand) under w = eq_outer + gamma^2 eq_product.
// Batched values (linear combinations of columns) are built with ZOLT_VEC_LIN2: binding is linear, so the bound batched
// table is the batched bound tables, exactly.
namespace stage3 {

class DevPool {
 public:
  explicit DevPool(Context& ctx) : ctx_(ctx) {}
  ~DevPool() {
    for (auto& b : owned_) zolt_dev_release_cached(ctx_.raw(), b.first, b.second);
  }
  DevPool(const DevPool&) = delete;
  DevPool& operator=(const DevPool&) = delete;
  Fr* alloc(size_t elems) {
    void* p = nullptr;
    const size_t bytes = (elems ? elems : 1) * sizeof(Fr);
    must(zolt_dev_alloc_cached(ctx_.raw(), bytes, &p), "zolt_dev_alloc_cached");
    owned_.emplace_back(p, bytes);
    return static_cast<Fr*>(p);
  }
  // hand a temporary back to the context's cache before the prover ends (work queued on it is ordered on the same stream)
  void release(void* p) {
    for (size_t i = 0; i < owned_.size(); i++)
      if (owned_[i].first == p) {
        zolt_dev_release_cached(ctx_.raw(), p, owned_[i].second);
        owned_.erase(owned_.begin() + (long)i);
        return;
      }
  }
  // out = s0 a + s1 b
  void lin2(Fr* out, const void* a, const Fr& s0, const void* b, const Fr& s1, size_t n) {
    const Fr s[2] = {s0, s1};
    must(zolt_fr_vec_dev(ctx_.raw(), ZOLT_VEC_LIN2, out, a, b, nullptr, nullptr, reinterpret_cast<const uint64_t*>(s), n), "zolt_fr_vec_dev");
  }
  void fill(Fr* out, const Fr& v, size_t n) {
    const Fr s[2] = {v, Fr::zero()};
    must(zolt_fr_vec_dev(ctx_.raw(), ZOLT_VEC_AFFINE_INDEX, out, nullptr, nullptr, nullptr, nullptr, reinterpret_cast<const uint64_t*>(s), n),
         "zolt_fr_vec_dev");
  }
  void eq(Fr* out, const std::vector<Fr>& r) {
    if (r.empty()) return fill(out, Fr::one(), 1);
    must(zolt_eq_evals(ctx_.raw(), reinterpret_cast<const uint64_t*>(r.data()), (int)r.size(), nullptr, out), "zolt_eq_evals");
  }
  void eqPlusOne(Fr* out, const std::vector<Fr>& r) {
    if (r.empty()) return fill(out, Fr::zero(), 1);  // the empty sum of EqPlusOnePolynomial.mle
    must(zolt_eq_plus_one_evals(ctx_.raw(), reinterpret_cast<const uint64_t*>(r.data()), (int)r.size(), out), "zolt_eq_plus_one_evals");
  }
  void fold(Fr* out, const void* in, const void* w, size_t n_lo, size_t n_hi, const Fr* scale = nullptr) {
    must(zolt_fr_fold_high_dev(ctx_.raw(), out, in, w, n_lo, n_hi, scale ? scale->limbs.data() : nullptr), "zolt_fr_fold_high_dev");
  }
  void put(Fr* d, const Fr& v) { must(zolt_dev_upload(ctx_.raw(), d, &v, sizeof(Fr)), "zolt_dev_upload"); }
  Context& ctx() { return ctx_; }

 private:
  Context& ctx_;
  std::vector<std::pair<void*, size_t>> owned_;
};

inline std::array<Fr, 4> scEvals(zolt_sc* sc) {
  uint64_t buf[16] = {0};
  must(zolt_sc_round_evals(sc, buf), "zolt_sc_round_evals");
  return {Fr::fromRaw(buf), Fr::fromRaw(buf + 4), Fr::fromRaw(buf + 8), Fr::fromRaw(buf + 12)};
}
inline void scBind(zolt_sc* sc, const Fr& r) { must(zolt_sc_bind(sc, r.limbs.data()), "zolt_sc_bind"); }
inline Fr scEntry(Context& ctx, zolt_sc* sc, int table, size_t index = 0) {
  const void* d = nullptr;
  must(zolt_sc_table_dev(sc, table, &d), "zolt_sc_table_dev");
  Fr v;
  must(zolt_dev_download(ctx.raw(), &v, static_cast<const Fr*>(d) + index, sizeof(Fr)), "zolt_dev_download");
  return v;
}
inline const void* scTable(zolt_sc* sc, int table) {
  const void* d = nullptr;
  must(zolt_sc_table_dev(sc, table, &d), "zolt_sc_table_dev");
  return d;
}
inline Fr product(const std::vector<Fr>& r) {
  Fr m = Fr::one();
  for (const Fr& x : r) m = m.mul(x);
  return m;
}

// InstructionInputProver (:1921-2150): degree 3, evaluations [p(0), p(1), p(2), p(3)]
class InstructionInputProver {
 public:
  struct Columns {
    const void *left_is_rs1, *rs1_value, *left_is_pc, *unexpanded_pc, *right_is_rs2, *rs2_value, *right_is_imm, *imm;
  };
  struct FinalClaims {
    Fr left_is_rs1, rs1_value, left_is_pc, unexpanded_pc, right_is_rs2, rs2_value, right_is_imm, imm;
  };
  InstructionInputProver(Context& ctx, int log_T, const Columns& c, const std::vector<Fr>& r_outer, const std::vector<Fr>& r_product,
                         const Fr& gamma, const Fr& gamma_sqr)
      : pool_(ctx), gamma_(gamma), cols_(c), T_(size_t(1) << log_T) {
    if (log_T < 1 || (int)r_outer.size() != log_T || (int)r_product.size() != log_T) panic("InstructionInputProver: bad point", ZOLT_ERR_ARG);
    const size_t T = T_;
    Fr* eo = pool_.alloc(T);
    Fr* ep = pool_.alloc(T);
    pool_.eq(eo, r_outer);
    pool_.eq(ep, r_product);
    w_ = pool_.alloc(T);
    pool_.lin2(w_, eo, Fr::one(), ep, gamma_sqr, T);
    pool_.release(eo);
    pool_.release(ep);
    const void* right[5] = {w_, c.right_is_rs2, c.rs2_value, c.right_is_imm, c.imm};
    const void* left[5] = {w_, c.left_is_rs1, c.rs1_value, c.left_is_pc, c.unexpanded_pc};
    must(zolt_sc_create_dev(ctx.raw(), ZOLT_SC_EQ_DOT2, ZOLT_PAIR_ADJACENT, log_T, 5, 0, right, &right_), "zolt_sc_create_dev");
    must(zolt_sc_create_dev(ctx.raw(), ZOLT_SC_EQ_DOT2, ZOLT_PAIR_ADJACENT, log_T, 5, 0, left, &left_), "zolt_sc_create_dev");
  }
  ~InstructionInputProver() {
    if (right_) zolt_sc_free(right_);
    if (left_) zolt_sc_free(left_);
  }
  // sum_i w_i (right_i + gamma left_i): the full_sum of :262-276 (what the opening claims must add up to)
  // (in the prover it comes from the opening claims, computeInstructionInputClaim :808-825; call before the first bind)
  Fr inputClaim() {
    auto mul = [&](Fr* out, const void* a, const void* b) {
      must(zolt_fr_vec_dev(pool_.ctx().raw(), ZOLT_VEC_MUL, out, a, b, nullptr, nullptr, nullptr, T_), "zolt_fr_vec_dev");
    };
    Fr* x = pool_.alloc(T_);
    Fr* y = pool_.alloc(T_);
    Fr* right = pool_.alloc(T_);
    Fr* left = pool_.alloc(T_);
    mul(x, cols_.right_is_rs2, cols_.rs2_value);
    mul(y, cols_.right_is_imm, cols_.imm);
    pool_.lin2(right, x, Fr::one(), y, Fr::one(), T_);
    mul(x, cols_.left_is_rs1, cols_.rs1_value);
    mul(y, cols_.left_is_pc, cols_.unexpanded_pc);
    pool_.lin2(left, x, Fr::one(), y, Fr::one(), T_);
    pool_.lin2(x, right, Fr::one(), left, gamma_, T_);
    Fr claim;
    must(zolt_fr_inner_product_dev(pool_.ctx().raw(), w_, x, T_, claim.limbs.data()), "zolt_fr_inner_product_dev");
    for (Fr* t : {x, y, right, left}) pool_.release(t);
    return claim;
  }
  std::array<Fr, 4> computeRoundEvals(const Fr& previous_claim) {
    const std::array<Fr, 4> r = scEvals(right_), l = scEvals(left_);  // [p(0), p(2), p(3)] of each side
    const Fr p0 = r[0].add(gamma_.mul(l[0]));
    return {p0, previous_claim.sub(p0), r[1].add(gamma_.mul(l[1])), r[2].add(gamma_.mul(l[2]))};
  }
  void bind(const Fr& r_j) {
    scBind(right_, r_j);
    scBind(left_, r_j);
  }
  FinalClaims finalClaims() {
    return {scEntry(pool_.ctx(), left_, 1), scEntry(pool_.ctx(), left_, 2), scEntry(pool_.ctx(), left_, 3), scEntry(pool_.ctx(), left_, 4),
            scEntry(pool_.ctx(), right_, 1), scEntry(pool_.ctx(), right_, 2), scEntry(pool_.ctx(), right_, 3), scEntry(pool_.ctx(), right_, 4)};
  }

 private:
  DevPool pool_;
  Fr gamma_;
  Columns cols_;
  size_t T_;
  Fr* w_ = nullptr;
  zolt_sc* right_ = nullptr;
  zolt_sc* left_ = nullptr;
};

// ShiftPrefixSuffixProver (:928-1920): degree 2, evaluations [p(0), p(1), p(2)]
class ShiftPrefixSuffixProver {
 public:
  struct Columns {
    const void *unexpanded_pc, *pc, *is_virtual, *is_first_in_sequence, *is_noop;
  };
  struct FinalClaims {
    Fr unexpanded_pc, pc, is_virtual, is_first_in_sequence, is_noop;
  };
  ShiftPrefixSuffixProver(Context& ctx, int log_T, const Columns& c, const std::vector<Fr>& r_outer, const std::vector<Fr>& r_product,
                          const std::vector<Fr>& gamma_powers)
      : pool_(ctx), gp_(gamma_powers) {
    if (log_T < 2 || (int)r_outer.size() != log_T || (int)r_product.size() != log_T || gamma_powers.size() < 5)
      panic("ShiftPrefixSuffixProver: bad arguments", ZOLT_ERR_ARG);
    const int split = log_T / 2;
    prefix_vars_ = log_T - split;
    suffix_vars_ = split;
    const size_t ps = size_t(1) << prefix_vars_, ss = size_t(1) << suffix_vars_, T = size_t(1) << log_T;
    r_hi_[0].assign(r_outer.begin(), r_outer.begin() + split);
    r_hi_[1].assign(r_product.begin(), r_product.begin() + split);
    const std::vector<Fr> lo_o(r_outer.begin() + split, r_outer.end()), lo_p(r_product.begin() + split, r_product.end());
    // P buffers end to end: eq+1(r_lo, .), is_max(r_lo) delta_0, for r_outer then r_product (:1006-1030)
    Fr* P = pool_.alloc(4 * ps);
    pool_.eqPlusOne(P, lo_o);
    pool_.fill(P + ps, Fr::zero(), ps);
    pool_.put(P + ps, product(lo_o));
    pool_.eqPlusOne(P + 2 * ps, lo_p);
    pool_.fill(P + 3 * ps, Fr::zero(), ps);
    pool_.put(P + 3 * ps, product(lo_p));
    // suffix tables eq(r_hi, .), eq+1(r_hi, .) (:1046-1047)
    for (int k = 0; k < 2; k++) {
      suf_[2 * k] = pool_.alloc(ss);
      suf_[2 * k + 1] = pool_.alloc(ss);
      pool_.eq(suf_[2 * k], r_hi_[k]);
      pool_.eqPlusOne(suf_[2 * k + 1], r_hi_[k]);
    }
    // v = upc + gamma pc + gamma^2 virt + gamma^3 first, nn = 1 - noop (:1086-1098)
    Fr* t1 = pool_.alloc(T);
    Fr* t2 = pool_.alloc(T);
    Fr* v = pool_.alloc(T);
    pool_.lin2(t1, c.unexpanded_pc, Fr::one(), c.pc, gp_[1], T);
    pool_.lin2(t2, c.is_virtual, gp_[2], c.is_first_in_sequence, gp_[3], T);
    pool_.lin2(v, t1, Fr::one(), t2, Fr::one(), T);
    pool_.fill(t1, Fr::one(), T);
    pool_.lin2(t2, t1, Fr::one(), c.is_noop, Fr::one().neg(), T);  // nn
    Fr* Q = pool_.alloc(4 * ps);
    pool_.fold(Q, v, suf_[0], ps, ss);
    pool_.fold(Q + ps, v, suf_[1], ps, ss);
    pool_.fold(Q + 2 * ps, t2, suf_[2], ps, ss, &gp_[4]);
    pool_.fold(Q + 3 * ps, t2, suf_[3], ps, ss, &gp_[4]);
    for (Fr* t : {t1, t2, v}) pool_.release(t);  // T-sized temporaries: only the sqrt(T)-sized buffers live on
    const void* pq[2] = {P, Q};
    must(zolt_sc_create_dev(ctx.raw(), ZOLT_SC_PROD2, ZOLT_PAIR_ADJACENT, prefix_vars_ + 2, 2, 0, pq, &pq_), "zolt_sc_create_dev");
    const void* wit[5] = {c.unexpanded_pc, c.pc, c.is_virtual, c.is_first_in_sequence, c.is_noop};
    must(zolt_sc_create_dev(ctx.raw(), ZOLT_SC_SUM, ZOLT_PAIR_ADJACENT, log_T, 1, 4, wit, &wit_), "zolt_sc_create_dev");
  }
  ~ShiftPrefixSuffixProver() {
    for (zolt_sc* s : {a_, b_, pq_, wit_})  // the phase-2 instances borrow wit_'s tables: they go first
      if (s) zolt_sc_free(s);
  }
  // grand_sum(P * Q) (:1131-1139)
  Fr inputClaim() {
    const std::array<Fr, 4> e = scEvals(pq_);
    return e[0].add(e[1]);
  }
  std::array<Fr, 3> computeRoundEvals(const Fr& previous_claim) {
    if (!in_phase2_) {
      const std::array<Fr, 4> e = scEvals(pq_);
      return {e[0], e[1], e[2]};
    }
    const std::array<Fr, 4> a = scEvals(a_), b = scEvals(b_);
    const Fr p0 = a[0].add(b[0]);
    return {p0, previous_claim.sub(p0), a[2].add(b[2])};
  }
  void bind(const Fr& r_j) {
    if (in_phase2_) {
      scBind(a_, r_j);
      scBind(b_, r_j);
      return;
    }
    scEvals(pq_);
    scBind(pq_, r_j);
    scEvals(wit_);
    scBind(wit_, r_j);
    if (++bound_ == prefix_vars_) transitionToPhase2();
  }
  FinalClaims finalClaims() { return {scEntry(pool_.ctx(), a_, 2), scEntry(pool_.ctx(), a_, 3), scEntry(pool_.ctx(), a_, 4), scEntry(pool_.ctx(), a_, 5), scEntry(pool_.ctx(), b_, 2)}; }

 private:
  void transitionToPhase2() {
    in_phase2_ = true;
    const size_t ss = size_t(1) << suffix_vars_;
    // the bound P buffers ARE evaluateMle(prefix, r_prefix) (:1571-1603)
    const Fr p0o = scEntry(pool_.ctx(), pq_, 0, 0), p1o = scEntry(pool_.ctx(), pq_, 0, 1), p0p = scEntry(pool_.ctx(), pq_, 0, 2), p1p = scEntry(pool_.ctx(), pq_, 0, 3);
    Fr* eo = pool_.alloc(ss);
    Fr* ep = pool_.alloc(ss);
    pool_.lin2(eo, suf_[0], p0o, suf_[1], p1o, ss);
    pool_.lin2(ep, suf_[2], p0p, suf_[3], p1p, ss);
    const void* w[5];
    for (int k = 0; k < 5; k++) w[k] = scTable(wit_, k);
    Fr* t1 = pool_.alloc(ss);
    Fr* t2 = pool_.alloc(ss);
    Fr* val = pool_.alloc(ss);
    pool_.lin2(t1, w[0], Fr::one(), w[1], gp_[1], ss);
    pool_.lin2(t2, w[2], gp_[2], w[3], gp_[3], ss);
    pool_.lin2(val, t1, Fr::one(), t2, Fr::one(), ss);
    Fr* nn = pool_.alloc(ss);
    pool_.fill(t1, Fr::one(), ss);
    pool_.lin2(nn, t1, gp_[4], w[4], gp_[4].neg(), ss);  // gamma^4 (1 - noop)
    const void* ta[6] = {eo, val, w[0], w[1], w[2], w[3]};
    const void* tb[3] = {ep, nn, w[4]};
    must(zolt_sc_create_dev(pool_.ctx().raw(), ZOLT_SC_PROD2, ZOLT_PAIR_ADJACENT, suffix_vars_, 2, 4, ta, &a_), "zolt_sc_create_dev");
    must(zolt_sc_create_dev(pool_.ctx().raw(), ZOLT_SC_PROD2, ZOLT_PAIR_ADJACENT, suffix_vars_, 2, 1, tb, &b_), "zolt_sc_create_dev");
  }

  DevPool pool_;
  std::vector<Fr> gp_;
  std::vector<Fr> r_hi_[2];
  int prefix_vars_ = 0, suffix_vars_ = 0, bound_ = 0;
  bool in_phase2_ = false;
  Fr* suf_[4] = {};
  zolt_sc *pq_ = nullptr, *wit_ = nullptr, *a_ = nullptr, *b_ = nullptr;
};

// RegistersPrefixSuffixProver (:2156-2496): degree 2, evaluations [p(0), p(1), p(2)]
class RegistersPrefixSuffixProver {
 public:
  struct Columns {
    const void *rd_write_value, *rs1_value, *rs2_value;
  };
  struct FinalClaims {
    Fr rd_write_value, rs1_value, rs2_value;
  };
  RegistersPrefixSuffixProver(Context& ctx, int log_T, const Columns& c, const std::vector<Fr>& r_spartan, const Fr& gamma,
                              const Fr& gamma_sqr)
      : pool_(ctx), gamma_(gamma), gamma_sqr_(gamma_sqr) {
    if (log_T < 2 || (int)r_spartan.size() != log_T) panic("RegistersPrefixSuffixProver: bad arguments", ZOLT_ERR_ARG);
    const int split = log_T / 2;
    prefix_vars_ = log_T - split;
    suffix_vars_ = split;
    r_hi_.assign(r_spartan.begin(), r_spartan.begin() + split);
    r_lo_.assign(r_spartan.begin() + split, r_spartan.end());
    const size_t ps = size_t(1) << prefix_vars_, ss = size_t(1) << suffix_vars_, T = size_t(1) << log_T;
    Fr* P = pool_.alloc(ps);
    pool_.eq(P, r_lo_);
    suffix_ = pool_.alloc(ss);
    pool_.eq(suffix_, r_hi_);
    Fr* v = pool_.alloc(T);
    Fr* t = pool_.alloc(T);
    pool_.lin2(t, c.rd_write_value, Fr::one(), c.rs1_value, gamma, T);
    pool_.lin2(v, t, Fr::one(), c.rs2_value, gamma_sqr, T);
    Fr* Q = pool_.alloc(ps);
    pool_.fold(Q, v, suffix_, ps, ss);
    pool_.release(v);
    pool_.release(t);
    const void* pq[2] = {P, Q};
    must(zolt_sc_create_dev(ctx.raw(), ZOLT_SC_PROD2, ZOLT_PAIR_ADJACENT, prefix_vars_, 2, 0, pq, &pq_), "zolt_sc_create_dev");
    const void* wit[3] = {c.rd_write_value, c.rs1_value, c.rs2_value};
    must(zolt_sc_create_dev(ctx.raw(), ZOLT_SC_SUM, ZOLT_PAIR_ADJACENT, log_T, 1, 2, wit, &wit_), "zolt_sc_create_dev");
  }
  ~RegistersPrefixSuffixProver() {
    for (zolt_sc* s : {eqv_, pq_, wit_})  // eqv_ borrows wit_'s tables: it goes first
      if (s) zolt_sc_free(s);
  }
  Fr inputClaim() {
    const std::array<Fr, 4> e = scEvals(pq_);
    return e[0].add(e[1]);
  }
  std::array<Fr, 3> computeRoundEvals(const Fr& previous_claim) {
    const std::array<Fr, 4> e = scEvals(in_phase2_ ? eqv_ : pq_);
    return {e[0], previous_claim.sub(e[0]), e[2]};
  }
  void bind(const Fr& r_j) {
    if (in_phase2_) return scBind(eqv_, r_j);
    scEvals(pq_);
    scBind(pq_, r_j);
    scEvals(wit_);
    scBind(wit_, r_j);
    prefix_challenges_.push_back(r_j);
    if ((int)prefix_challenges_.size() == prefix_vars_) transitionToPhase2();
  }
  FinalClaims finalClaims() { return {scEntry(pool_.ctx(), eqv_, 2), scEntry(pool_.ctx(), eqv_, 3), scEntry(pool_.ctx(), eqv_, 4)}; }

 private:
  void transitionToPhase2() {
    in_phase2_ = true;
    const size_t ss = size_t(1) << suffix_vars_;
    // phase2_eq[j] = eq(r_hi, j) * eq(r_lo, reversed prefix challenges) (:2427-2465)
    const std::vector<Fr> reversed(prefix_challenges_.rbegin(), prefix_challenges_.rend());
    const Fr eq_prefix_eval = EqPolynomial::mle(r_lo_, reversed);
    Fr* eq = pool_.alloc(ss);
    pool_.lin2(eq, suffix_, eq_prefix_eval, suffix_, Fr::zero(), ss);
    const void* w[3];
    for (int k = 0; k < 3; k++) w[k] = scTable(wit_, k);
    Fr* t = pool_.alloc(ss);
    Fr* v = pool_.alloc(ss);
    pool_.lin2(t, w[0], Fr::one(), w[1], gamma_, ss);
    pool_.lin2(v, t, Fr::one(), w[2], gamma_sqr_, ss);
    const void* tabs[5] = {eq, v, w[0], w[1], w[2]};
    must(zolt_sc_create_dev(pool_.ctx().raw(), ZOLT_SC_PROD2, ZOLT_PAIR_ADJACENT, suffix_vars_, 2, 3, tabs, &eqv_), "zolt_sc_create_dev");
  }

  DevPool pool_;
  Fr gamma_, gamma_sqr_;
  std::vector<Fr> r_hi_, r_lo_, prefix_challenges_;
  int prefix_vars_ = 0, suffix_vars_ = 0;
  bool in_phase2_ = false;
  Fr* suffix_ = nullptr;
  zolt_sc *pq_ = nullptr, *wit_ = nullptr, *eqv_ = nullptr;
};

// the quadratic through (0, e0) (1, e1) (2, e2) at r (evalsToCoeffs + evaluatePolyAtPoint, :846-926)
inline Fr quadAt(const std::array<Fr, 3>& e, const Fr& r) {
  static const Fr two_inv = Fr::fromU64(2).inverse();
  const Fr c2 = e[2].sub(e[1]).sub(e[1]).add(e[0]).mul(two_inv);
  const Fr c1 = e[1].sub(e[0]).sub(c2);
  return e[0].add(r.mul(c1.add(r.mul(c2))));
}

// One round of generateStage3Proof's loop (:393-560): the three provers' evaluations and the batched cubic
struct RoundEvals {
  std::array<Fr, 3> shift;
  std::array<Fr, 4> instr;
  std::array<Fr, 3> reg;
  std::array<Fr, 4> combined;
};
struct Provers {
  ShiftPrefixSuffixProver& shift;
  InstructionInputProver& instr;
  RegistersPrefixSuffixProver& reg;
  std::array<Fr, 3> claims;  // current_shift_claim, current_instr_claim, current_reg_claim
  RoundEvals computeRound(const std::array<Fr, 3>& batching_coeffs) {
    RoundEvals e;
    e.shift = shift.computeRoundEvals(claims[0]);
    e.instr = instr.computeRoundEvals(claims[1]);
    e.reg = reg.computeRoundEvals(claims[2]);
    const Fr three = Fr::fromU64(3);
    const Fr s3 = e.shift[2].mul(three).sub(e.shift[1].mul(three)).add(e.shift[0]);
    const Fr r3 = e.reg[2].mul(three).sub(e.reg[1].mul(three)).add(e.reg[0]);
    for (int i = 0; i < 4; i++) {
      const Fr sv = i < 3 ? e.shift[i] : s3, rv = i < 3 ? e.reg[i] : r3;
      e.combined[i] = sv.mul(batching_coeffs[0]).add(e.instr[i].mul(batching_coeffs[1])).add(rv.mul(batching_coeffs[2]));
    }
    return e;
  }
  void bind(const RoundEvals& e, const Fr& r_j) {
    claims = {quadAt(e.shift, r_j), evaluateCubicAtPoint(e.instr, r_j), quadAt(e.reg, r_j)};
    shift.bind(r_j);
    instr.bind(r_j);
    reg.bind(r_j);
  }
};

}  // namespace stage3

// ---- LassoProver (src/zkvm/lasso/prover.zig:82-470) -------------------------------------------------------------------------
// eq_evals lives on the device.  Address rounds: zolt_lasso_addr_round applies the previous challenge and returns this round's
// partition sums in ONE pass over the table (the claim the reference recomputes after every bind, :402-407, is p(r) of the
// polynomial just sent -- the same field element -- so it costs nothing here).  Cycle rounds: the ZOLT_SC_SUM half-split
// instance over the same table (sum of the first / second half, fold (1 - r) lo + r hi = bindFirst).  expanding_v.get(0),
// the prover's final_eval, is prod (1 - r_i) over the address challenges; the 2^log_K-entry table is not materialised.
// The prefix-suffix decompositions and the split eq's bindOuter / bindInner never reach the prover's outputs and have no mirror.
struct LassoParams {
  Fr gamma;
  size_t log_T = 0, log_K = 0;
  std::vector<Fr> r_reduction;
};
class LassoProver {
 public:
  // d_lookup_indices: num_lookups x 16 bytes (u128, little-endian), device-resident and only read
  LassoProver(Context& ctx, const void* d_lookup_indices, size_t num_lookups, const LassoParams& params)
      : pool_(ctx), params_(params), d_idx_(d_lookup_indices), num_lookups_(num_lookups) {
    const size_t log_T = params.log_T, padded = size_t(1) << log_T;
    if (params.r_reduction.size() != log_T || num_lookups > padded || params.log_K > 128 || log_T > 40)
      panic("LassoProver: bad parameters", ZOLT_ERR_ARG);
    // eq_evals[j] = E_out[j >> inner] * E_in[j & mask] (:166-176), variable i of each half <-> bit i of its index
    // (split_eq.zig:120-142): as ONE big-endian table, the point is (w_out reversed, w_in reversed)
    const size_t outer = log_T / 2;
    std::vector<Fr> be(params.r_reduction.rend() - (long)outer, params.r_reduction.rend());
    be.insert(be.end(), params.r_reduction.rbegin(), params.r_reduction.rbegin() + (long)(log_T - outer));
    eq_ = pool_.alloc(padded);
    pool_.eq(eq_, be);
    if (num_lookups < padded) pool_.fill(eq_ + num_lookups, Fr::zero(), padded - num_lookups);  // cycles without lookups (:171-174)
    must(zolt_poly_sum(ctx.raw(), eq_, padded, current_claim.limbs.data()), "zolt_poly_sum");
  }
  ~LassoProver() {
    if (cycle_) zolt_sc_free(cycle_);
  }
  LassoProver(const LassoProver&) = delete;
  LassoProver& operator=(const LassoProver&) = delete;

  Fr computeInitialClaim() const { return current_claim; }
  bool isAddressPhase() const { return round < params_.log_K; }
  bool isComplete() const { return round >= params_.log_K + params_.log_T; }

  // coefficients [c0, c1, c2] of the round polynomial (UniPoly, :232-364)
  std::array<Fr, 3> computeRoundPolynomial() {
    if (have_poly_) return poly_;
    Fr s0, s1;
    if (isAddressPhase()) {
      uint64_t sums[8];
      must(zolt_lasso_addr_round(pool_.ctx().raw(), eq_, d_idx_, num_lookups_, pending_bit_, pending_bit_ >= 0 ? pending_r_.limbs.data() : nullptr,
                                 (int)round, sums),
           "zolt_lasso_addr_round");
      pending_bit_ = -1;
      s0 = Fr::fromRaw(sums);
      s1 = Fr::fromRaw(sums + 4);
    } else {
      enterCyclePhase();
      if (!cycle_) {  // log_T == 0: a single entry, nothing left to fold (:322-331)
        poly_ = {current_claim, Fr::zero(), Fr::zero()};
        have_poly_ = true;
        return poly_;
      }
      const std::array<Fr, 4> e = stage3::scEvals(cycle_);
      s0 = e[0];
      s1 = e[1];
    }
    poly_ = {s0, s1.sub(s0), Fr::zero()};
    have_poly_ = true;
    return poly_;
  }

  void receiveChallenge(const Fr& challenge) {
    const std::array<Fr, 3> p = computeRoundPolynomial();
    challenges.push_back(challenge);
    if (isAddressPhase()) {
      pending_bit_ = (int)round;  // applied by the next pass over the table
      pending_r_ = challenge;
      expanding_v0_ = expanding_v0_.mul(Fr::one().sub(challenge));
    } else if (cycle_) {
      stage3::scBind(cycle_, challenge);
    }
    // the sum of the updated table (:402-407, :441-446) = p(challenge)
    current_claim = p[0].add(p[1].mul(challenge)).add(p[2].mul(challenge).mul(challenge));
    have_poly_ = false;
    round++;
  }
  Fr getFinalEval() const { return expanding_v0_; }

  Fr current_claim;
  size_t round = 0;
  std::vector<Fr> challenges;

 private:
  void enterCyclePhase() {
    if (cycle_entered_) return;
    cycle_entered_ = true;
    if (pending_bit_ >= 0) {  // the last address challenge is still to be applied
      uint64_t sums[8];
      must(zolt_lasso_addr_round(pool_.ctx().raw(), eq_, d_idx_, num_lookups_, pending_bit_, pending_r_.limbs.data(), -1, sums),
           "zolt_lasso_addr_round");
      pending_bit_ = -1;
    }
    if (params_.log_T == 0) return;
    const void* t[1] = {eq_};
    must(zolt_sc_create_dev(pool_.ctx().raw(), ZOLT_SC_SUM, ZOLT_PAIR_HALF, (int)params_.log_T, 1, 0, t, &cycle_), "zolt_sc_create_dev");
  }

  stage3::DevPool pool_;
  LassoParams params_;
  const void* d_idx_;
  size_t num_lookups_;
  Fr* eq_ = nullptr;
  zolt_sc* cycle_ = nullptr;
  bool cycle_entered_ = false, have_poly_ = false;
  std::array<Fr, 3> poly_{};
  int pending_bit_ = -1;
  Fr pending_r_, expanding_v0_ = Fr::one();
};

// deriveChallenge (:537-551): the reference's stand-in for Fiat-Shamir inside runLassoProver
inline Fr lassoDeriveChallenge(const std::array<Fr, 3>& coeffs, size_t round) {
  uint64_t h = 0x9e3779b97f4a7c15ull;
  h ^= (uint64_t)round;
  h *= 0xff51afd7ed558ccdull;
  for (const Fr& c : coeffs)
    for (uint64_t limb : c.limbs) {
      h ^= limb;
      h *= 0xc4ceb9fe1a85ec53ull;
    }
  h ^= h >> 33;
  return Fr::fromU64(h);
}
struct LassoProof {
  std::vector<std::array<Fr, 3>> round_polys;
  Fr final_eval;
  std::vector<Fr> challenges;
};
// runLassoProver (:495-535)
inline LassoProof runLassoProver(Context& ctx, const void* d_lookup_indices, size_t num_lookups, const LassoParams& params) {
  LassoProver prover(ctx, d_lookup_indices, num_lookups, params);
  LassoProof proof;
  size_t round = 0;
  while (!prover.isComplete()) {
    proof.round_polys.push_back(prover.computeRoundPolynomial());
    prover.receiveChallenge(lassoDeriveChallenge(proof.round_polys.back(), round));
    round++;
  }
  proof.final_eval = prover.getFinalEval();
  proof.challenges = prover.challenges;
  return proof;
}

struct SumcheckProof {
  std::vector<std::array<Fr, 4>> round_polys;  // evaluations at 0..3
  std::vector<Fr> challenges;
};

// The stage loop of MultiStageProver (src/zkvm/prover.zig:405-428): per round absorb the first
// degree+1 evaluations as "round_poly_j", squeeze `challenge_label`, bind.
// Appends inst's rounds to `proof` (the transcript carries over from whatever ran before).
inline void proveSumcheckInto(SumcheckProof& proof, SumcheckInstance& inst, Transcript& transcript, std::string_view challenge_label,
                              size_t max_rounds = ~size_t(0)) {
  const size_t n_absorb = inst.degree() + 1;
  for (size_t round = 0; round < inst.numRounds() && round < max_rounds; round++) {
    const auto evals = inst.computeRoundPoly(round);
    for (size_t j = 0; j < n_absorb; j++) {
      char label[24];
      const int ln = std::snprintf(label, sizeof label, "round_poly_%zu", j);
      transcript.appendScalar(std::string_view(label, (size_t)ln), evals[j]);
    }
    const Fr r = transcript.challengeScalar(challenge_label);
    inst.bindChallenge(r);
    proof.round_polys.push_back(evals);
    proof.challenges.push_back(r);
  }
}

// batched_sumcheck.BatchedSumcheckProver (src/zkvm/batched_sumcheck.zig:77-262): random linear combination of several
// instances with different numbers of rounds.  Pure O(#instances) glue per round around the instances' device kernels.
// The reference samples the batching coefficients from its Blake2b transcript (:150-156), which is outside this tier;
// here the caller supplies them (and the per-round challenges come from whatever transcript drives the proof).
class BatchedSumcheckProver {
 public:
  void addInstance(SumcheckInstance* inst, const Fr& input_claim) {  // :117-122
    instances_.push_back(inst);
    claims_.push_back(input_claim);
    if (inst->numRounds() > max_rounds_) max_rounds_ = inst->numRounds();
  }
  // coefficients + initial batched claim  sum_i coeff_i * claim_i * 2^(max_rounds - rounds_i)  (:158-171)
  void setBatchingCoeffs(const std::vector<Fr>& coeffs) {
    coeffs_ = coeffs;
    current_claim = Fr::zero();
    for (size_t i = 0; i < instances_.size(); i++)
      current_claim = current_claim.add(scaled(claims_[i], max_rounds_ - instances_[i]->numRounds()).mul(coeffs_[i]));
  }
  size_t maxNumRounds() const { return max_rounds_; }
  // combined evaluations at X = 0..3 of this round (:189-220)
  std::array<Fr, 4> computeRoundEvals() {
    std::array<Fr, 4> combined{Fr::zero(), Fr::zero(), Fr::zero(), Fr::zero()};
    for (size_t i = 0; i < instances_.size(); i++) {
      const size_t start = max_rounds_ - instances_[i]->numRounds();
      if (current_round_ >= start) {
        const auto evals = instances_[i]->computeRoundPoly(current_round_ - start);
        for (int j = 0; j < 4; j++) combined[j] = combined[j].add(evals[j].mul(coeffs_[i]));
      } else {
        // not started yet: constant polynomial, the input claim scaled by the rounds still to wait (:206-218)
        const Fr w = scaled(claims_[i], start - current_round_).mul(coeffs_[i]);
        for (int j = 0; j < 4; j++) combined[j] = combined[j].add(w);
      }
    }
    return combined;
  }
  std::array<Fr, 3> computeRoundPolynomial() { return UniPoly::evalsToCompressed(computeRoundEvals()); }  // :223
  void updateClaim(const std::array<Fr, 4>& round_evals, const Fr& challenge) {  // :242-245
    current_claim = evaluateCubicAtPoint(round_evals, challenge);
  }
  void bindChallenge(const Fr& challenge) {  // :227-239
    for (auto* inst : instances_)
      if (current_round_ >= max_rounds_ - inst->numRounds()) inst->bindChallenge(challenge);
    current_round_++;
  }
  Fr current_claim = Fr::zero();

 private:
  static Fr scaled(Fr c, size_t doublings) {
    for (size_t k = 0; k < doublings; k++) c = c.add(c);
    return c;
  }
  std::vector<SumcheckInstance*> instances_;
  std::vector<Fr> claims_, coeffs_;
  size_t max_rounds_ = 0, current_round_ = 0;
};

// One rank of a dense LowToHigh sumcheck whose tables are sharded over `world` GPUs in contiguous hypercube
// blocks.  Every rank runs this with its own block; the round kernels exchange their sums through
// the NVLink mailboxes, so each rank sees the global round polynomials, runs the same Keccak transcript and
// derives the same challenges.  When the local block is down to ZOLT_SC_TAIL_ENTRIES entries per table the blocks are
// all-gathered over NVLink and the remaining rounds run on every rank alike, without exchanges.  Returns the whole proof; final_out
// receives the k fully bound table values.
inline SumcheckProof proveSumcheckSharded(Context& ctx, zolt_peer_box* box, int world, int combiner,
                                          const std::vector<const void*>& d_tables, int local_log, Transcript& transcript,
                                          std::string_view challenge_label, std::vector<Fr>* final_out,
                                          const Fr* input_claim = nullptr) {
  SumcheckProof proof;
  const int k = (int)d_tables.size();
  DenseSumcheckProver local(ctx, DenseSumcheckProver::DeviceTables{}, combiner, ZOLT_PAIR_ADJACENT, d_tables, local_log);
  if (input_claim) local.setInputClaim(*input_claim);  // the GLOBAL claim: every rank's round kernel folds all ranks' sums
  if (world == 1) {
    proveSumcheckInto(proof, local, transcript, challenge_label);
    if (final_out)
      for (int t = 0; t < k; t++) final_out->push_back(local.table(t)[0]);
    return proof;
  }
  must(zolt_sc_set_peers(local.raw(), box), "zolt_sc_set_peers");
  // exchange rounds while the block is larger than the tail threshold ...
  int tail_local_log = 0;
  while ((size_t(2) << tail_local_log) <= (size_t)ZOLT_SC_TAIL_ENTRIES && tail_local_log < local_log) tail_local_log++;
  proveSumcheckInto(proof, local, transcript, challenge_label, (size_t)(local_log - tail_local_log));
  // ... then one all-gather of the blocks and the remaining rounds on every rank alike
  std::vector<const void*> tail_tabs((size_t)k);
  size_t tail_n = 0;
  must(zolt_sc_gather_tail(local.raw(), tail_tabs.data(), &tail_n), "zolt_sc_gather_tail");
  int tail_log = 0;
  while ((size_t(1) << tail_log) < tail_n) tail_log++;
  DenseSumcheckProver tail(ctx, DenseSumcheckProver::DeviceTables{}, combiner, ZOLT_PAIR_ADJACENT, tail_tabs, tail_log);
  proveSumcheckInto(proof, tail, transcript, challenge_label);
  if (final_out)
    for (int t = 0; t < k; t++) final_out->push_back(tail.table(t)[0]);
  return proof;
}

// The dense LowToHigh sumcheck of a single-process prover over the devices of a zolt_multi: host tables are split into
// contiguous hypercube blocks, block r is uploaded to device r, and one host thread per device runs proveSumcheckSharded
// (every thread keeps its own copy of the transcript, they all derive the same challenges).  Returns rank 0's proof.
inline SumcheckProof proveSumcheckMultiDevice(zolt_multi* multi, int combiner, const std::vector<const uint64_t*>& tables, int log_n,
                                              std::string_view domain, std::string_view challenge_label, std::vector<Fr>* final_out,
                                              const Fr* input_claim = nullptr) {
  int world = 0;
  must(zolt_multi_size(multi, &world), "zolt_multi_size");
  int shift = 0;
  while ((1 << shift) < world) shift++;
  if ((1 << shift) != world || shift > log_n) panic("proveSumcheckMultiDevice: device count must be a power of two <= the table length", ZOLT_ERR_ARG);
  const int local_log = log_n - shift;
  const size_t local = size_t(1) << local_log;
  const int k = (int)tables.size();
  std::vector<SumcheckProof> proofs((size_t)world);
  std::vector<std::vector<Fr>> finals((size_t)world);
  std::vector<std::vector<void*>> bufs((size_t)world, std::vector<void*>((size_t)k, nullptr));
  // stage 1 (no exchange yet): upload the blocks and warm the pools; ranks sharing a device must not allocate while a peer's
  // kernel is already waiting for them
  for (int r = 0; r < world; r++) {
    zolt_ctx* c = nullptr;
    must(zolt_multi_rank(multi, r, &c, nullptr), "zolt_multi_rank");
    for (int t = 0; t < k; t++) {
      must(zolt_dev_alloc(c, local * 32, &bufs[(size_t)r][(size_t)t]), "zolt_dev_alloc");
      must(zolt_dev_upload(c, bufs[(size_t)r][(size_t)t], tables[(size_t)t] + 4 * local * (size_t)r, local * 32), "zolt_dev_upload");
    }
    Context ctx(c);
    std::vector<const void*> d(bufs[(size_t)r].begin(), bufs[(size_t)r].end());
    DenseSumcheckProver warm(ctx, DenseSumcheckProver::DeviceTables{}, combiner, ZOLT_PAIR_ADJACENT, d, local_log);
    Transcript tr(domain);
    SumcheckProof p;
    proveSumcheckInto(p, warm, tr, challenge_label);
    must(zolt_gpu_sync(c), "zolt_gpu_sync");
  }
  std::vector<std::thread> pool;
  for (int r = 0; r < world; r++) {
    pool.emplace_back([&, r]() {
      zolt_ctx* c = nullptr;
      zolt_peer_box* b = nullptr;
      must(zolt_multi_rank(multi, r, &c, &b), "zolt_multi_rank");
      Context ctx(c);
      std::vector<const void*> d(bufs[(size_t)r].begin(), bufs[(size_t)r].end());
      Transcript tr(domain);
      proofs[(size_t)r] = proveSumcheckSharded(ctx, b, world, combiner, d, local_log, tr, challenge_label, &finals[(size_t)r], input_claim);
      must(zolt_gpu_sync(c), "zolt_gpu_sync");
    });
  }
  for (auto& t : pool) t.join();
  for (int r = 0; r < world; r++) {
    zolt_ctx* c = nullptr;
    must(zolt_multi_rank(multi, r, &c, nullptr), "zolt_multi_rank");
    for (void* p : bufs[(size_t)r]) zolt_dev_free(c, p);
  }
  if (final_out) *final_out = finals[0];
  return proofs[0];
}

inline SumcheckProof proveSumcheck(SumcheckInstance& inst, Transcript& transcript, std::string_view challenge_label) {
  SumcheckProof proof;
  const size_t n_absorb = inst.degree() + 1;
  for (size_t round = 0; round < inst.numRounds(); round++) {
    const auto evals = inst.computeRoundPoly(round);
    for (size_t j = 0; j < n_absorb; j++) {
      char label[24];
      const int ln = std::snprintf(label, sizeof label, "round_poly_%zu", j);
      transcript.appendScalar(std::string_view(label, (size_t)ln), evals[j]);
    }
    const Fr r = transcript.challengeScalar(challenge_label);
    inst.bindChallenge(r);
    proof.round_polys.push_back(evals);
    proof.challenges.push_back(r);
  }
  return proof;
}

}  // namespace zolt::host
