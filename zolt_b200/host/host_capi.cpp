// C exports of the C++ host mirror (zolt_host.hpp) so the Python harness (tests/, bench.py) can drive
// the same host code a Zig/C++ caller would: the Keccak transcript round loop around the GPU kernels.
#include "zolt_host.hpp"

using namespace zolt::host;

extern "C" {

// ---- transcript ------------------------------------------------------------------------------------
void* zolt_host_transcript_new(const char* domain, size_t n) { return new Transcript(std::string_view(domain, n)); }
void zolt_host_transcript_free(void* t) { delete static_cast<Transcript*>(t); }
void zolt_host_transcript_append_bytes(void* t, const void* data, size_t n) { static_cast<Transcript*>(t)->appendBytes(data, n); }
void zolt_host_transcript_append_scalar(void* t, const char* label, size_t ln, const uint64_t* scalar) {
  static_cast<Transcript*>(t)->appendScalar(std::string_view(label, ln), Fr::fromRaw(scalar));
}
void zolt_host_transcript_challenge_scalar(void* t, const char* label, size_t ln, uint64_t* out) {
  static_cast<Transcript*>(t)->challengeScalar(std::string_view(label, ln)).store(out);
}

// ---- host field glue (parity hook for fr_host.hpp) -----------------------------------------------------
// op: 0 mul, 1 add, 2 sub, 3 inverse, 4 fromBytes (a = raw 256-bit LE), 5 toBytes, 6 neg
void zolt_host_fr_op(int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    Fr x = Fr::fromRaw(a + 4 * i), y = b ? Fr::fromRaw(b + 4 * i) : Fr::zero(), r;
    switch (op) {
      case 0: r = x.mul(y); break;
      case 1: r = x.add(y); break;
      case 2: r = x.sub(y); break;
      case 3: r = x.isZero() ? x : x.inverse(); break;
      case 4: r = Fr::fromBytes(reinterpret_cast<const uint8_t*>(a + 4 * i)); break;
      case 5: x.toBytes(reinterpret_cast<uint8_t*>(out + 4 * i)); continue;
      default: r = x.neg(); break;
    }
    r.store(out + 4 * i);
  }
}
void zolt_host_evals_to_compressed(const uint64_t* evals, uint64_t* out3) {
  auto c = UniPoly::evalsToCompressed({Fr::fromRaw(evals), Fr::fromRaw(evals + 4), Fr::fromRaw(evals + 8), Fr::fromRaw(evals + 12)});
  for (int i = 0; i < 3; i++) c[i].store(out3 + 4 * i);
}
void zolt_host_cubic_eval(const uint64_t* evals, const uint64_t* x, uint64_t* out) {
  evaluateCubicAtPoint({Fr::fromRaw(evals), Fr::fromRaw(evals + 4), Fr::fromRaw(evals + 8), Fr::fromRaw(evals + 12)}, Fr::fromRaw(x)).store(out);
}

// ---- round loops ------------------------------------------------------------------------------------------
static void store_proof(const SumcheckProof& proof, uint64_t* evals_out, uint64_t* challenges_out) {
  for (size_t r = 0; r < proof.round_polys.size(); r++) {
    for (int j = 0; j < 4; j++) proof.round_polys[r][j].store(evals_out + 16 * r + 4 * j);
    proof.challenges[r].store(challenges_out + 4 * r);
  }
}

// Dense sumcheck over host tables with the Keccak transcript: upload, then per round
// {GPU evals -> absorb -> squeeze -> fused GPU bind+evals}.  final_out: n_tables elements.
int zolt_host_sumcheck_prove(zolt_ctx* ctx, int combiner, int pairing, int log_n, int n_tables, int n_extra,
                             const uint64_t* const* tables, const char* domain, size_t dn, const char* label, size_t ln,
                             uint64_t* evals_out, uint64_t* challenges_out, uint64_t* final_out) {
  Context c(ctx);
  std::vector<const uint64_t*> tabs(tables, tables + n_tables);
  DenseSumcheckProver prover(c, combiner, pairing, tabs, log_n, n_extra);
  Transcript tr(std::string_view(domain, dn));
  SumcheckProof proof = proveSumcheck(prover, tr, std::string_view(label, ln));
  store_proof(proof, evals_out, challenges_out);
  if (final_out)
    for (int t = 0; t < n_tables; t++) prover.table(t)[0].store(final_out + 4 * t);
  return ZOLT_OK;
}

// zolt_host_sumcheck_prove with the instance's input claim (may be NULL)
int zolt_host_sumcheck_prove_claim(zolt_ctx* ctx, int combiner, int pairing, int log_n, int n_tables, int n_extra,
                                   const uint64_t* const* tables, const uint64_t* claim, const char* domain, size_t dn,
                                   const char* label, size_t ln, uint64_t* evals_out, uint64_t* challenges_out, uint64_t* final_out) {
  Context c(ctx);
  std::vector<const uint64_t*> tabs(tables, tables + n_tables);
  DenseSumcheckProver prover(c, combiner, pairing, tabs, log_n, n_extra);
  if (claim) prover.setInputClaim(Fr::fromRaw(claim));
  Transcript tr(std::string_view(domain, dn));
  SumcheckProof proof = proveSumcheck(prover, tr, std::string_view(label, ln));
  store_proof(proof, evals_out, challenges_out);
  if (final_out)
    for (int t = 0; t < n_tables; t++) prover.table(t)[0].store(final_out + 4 * t);
  return ZOLT_OK;
}

// Same round loop over tables that are already resident on the device.
int zolt_host_sumcheck_prove_dev(zolt_ctx* ctx, int combiner, int pairing, int log_n, int n_tables, int n_extra,
                                 const void* const* d_tables, const char* domain, size_t dn, const char* label, size_t ln,
                                 uint64_t* evals_out, uint64_t* challenges_out, uint64_t* final_out) {
  Context c(ctx);
  std::vector<const void*> tabs(d_tables, d_tables + n_tables);
  DenseSumcheckProver prover(c, DenseSumcheckProver::DeviceTables{}, combiner, pairing, tabs, log_n, n_extra);
  Transcript tr(std::string_view(domain, dn));
  SumcheckProof proof = proveSumcheck(prover, tr, std::string_view(label, ln));
  store_proof(proof, evals_out, challenges_out);
  if (final_out)
    for (int t = 0; t < n_tables; t++) prover.table(t)[0].store(final_out + 4 * t);
  return ZOLT_OK;
}

// zolt_host_sumcheck_prove_dev with the instance's input claim (may be NULL): round 0 then derives p(1) from it
int zolt_host_sumcheck_prove_dev_claim(zolt_ctx* ctx, int combiner, int pairing, int log_n, int n_tables, int n_extra,
                                       const void* const* d_tables, const uint64_t* claim, const char* domain, size_t dn,
                                       const char* label, size_t ln, uint64_t* evals_out, uint64_t* challenges_out,
                                       uint64_t* final_out) {
  Context c(ctx);
  std::vector<const void*> tabs(d_tables, d_tables + n_tables);
  DenseSumcheckProver prover(c, DenseSumcheckProver::DeviceTables{}, combiner, pairing, tabs, log_n, n_extra);
  if (claim) prover.setInputClaim(Fr::fromRaw(claim));
  Transcript tr(std::string_view(domain, dn));
  SumcheckProof proof = proveSumcheck(prover, tr, std::string_view(label, ln));
  store_proof(proof, evals_out, challenges_out);
  if (final_out)
    for (int t = 0; t < n_tables; t++) prover.table(t)[0].store(final_out + 4 * t);
  return ZOLT_OK;
}

// Gruen split-eq sumcheck over tables resident on the device
int zolt_host_sumcheck_prove_gruen_dev(zolt_ctx* ctx, int with_c, int log_n, int n_tables, const void* const* d_tables,
                                       const uint64_t* tau, const uint64_t* claim, const char* domain, size_t dn,
                                       const char* label, size_t ln, uint64_t* evals_out, uint64_t* challenges_out,
                                       uint64_t* final_claim_out) {
  Context c(ctx);
  std::vector<const void*> tabs(d_tables, d_tables + n_tables);
  std::vector<Fr> tv(log_n);
  for (int i = 0; i < log_n; i++) tv[i] = Fr::fromRaw(tau + 4 * i);
  GruenProductProver prover(c, DenseSumcheckProver::DeviceTables{}, tabs, log_n, tv, Fr::fromRaw(claim), with_c != 0);
  Transcript tr(std::string_view(domain, dn));
  SumcheckProof proof = proveSumcheck(prover, tr, std::string_view(label, ln));
  store_proof(proof, evals_out, challenges_out);
  if (final_claim_out) prover.current_claim.store(final_claim_out);
  return ZOLT_OK;
}

// One process, several devices: dense sumcheck over the devices of a zolt_multi (one host thread per device)
int zolt_host_sumcheck_prove_multi(zolt_multi* multi, int combiner, int log_n, int n_tables, const uint64_t* const* tables,
                                   const uint64_t* claim, const char* domain, size_t dn, const char* label, size_t ln,
                                   uint64_t* evals_out, uint64_t* challenges_out, uint64_t* final_out) {
  std::vector<const uint64_t*> tabs(tables, tables + n_tables);
  std::vector<Fr> fin;
  Fr cl;
  if (claim) cl = Fr::fromRaw(claim);
  SumcheckProof proof = proveSumcheckMultiDevice(multi, combiner, tabs, log_n, std::string_view(domain, dn), std::string_view(label, ln), &fin,
                                                 claim ? &cl : nullptr);
  store_proof(proof, evals_out, challenges_out);
  if (final_out)
    for (int t = 0; t < n_tables; t++) fin[t].store(final_out + 4 * t);
  return ZOLT_OK;
}

// Stage4GruenProver over device-resident tables.  challenges != NULL: the caller's round loop (proof_converter.zig:2112-2282 --
// computeRoundEvals(round, claim) / claim = cubic at the challenge / bindChallenge), evals_out = rounds x 4 evaluations.
// challenges == NULL: Stage4GruenProver.prove against the Keccak transcript of this repository (appendScalar ->
// "stage4_coeff", challengeScalar -> "stage4_r"), evals_out = rounds x 4 UNBATCHED coefficients, challenges_out filled.
// input_claim may be NULL (computeInputClaim over the tables).  finals_out: val, rs1_ra, rs2_ra, rd_wa, inc, eq (6 x 4).
namespace {
struct KeccakAdapter {
  Transcript& t;
  void appendScalar(const Fr& s) { t.appendScalar("stage4_coeff", s); }
  Fr challengeScalar() { return t.challengeScalar("stage4_r"); }
};
}  // namespace
int zolt_host_stage4_run(zolt_ctx* ctx, int log_T, const void* d_val, const void* d_rd_wa, const void* d_rs1_ra, const void* d_rs2_ra,
                         const void* d_inc, const uint64_t* gamma, const uint64_t* r_cycle, int phase1_rounds,
                         const uint64_t* input_claim, const uint64_t* batching_coeff, const uint64_t* challenges, const char* domain,
                         size_t dn, uint64_t* evals_out, uint64_t* challenges_out, uint64_t* finals_out, uint64_t* claim_out) {
  Context c(ctx);
  std::vector<Fr> rc(log_T);
  for (int i = 0; i < log_T; i++) rc[i] = Fr::fromRaw(r_cycle + 4 * i);
  const Fr coeff = batching_coeff ? Fr::fromRaw(batching_coeff) : Fr::one();
  Stage4GruenProver prover(c, log_T, d_val, d_rd_wa, d_rs1_ra, d_rs2_ra, d_inc, Fr::fromRaw(gamma), rc, nullptr, coeff, phase1_rounds);
  const size_t rounds = prover.numRounds();
  if (challenges) {
    Fr claim = input_claim ? Fr::fromRaw(input_claim) : prover.computeInputClaim();
    if (claim_out) claim.store(claim_out);
    for (size_t r = 0; r < rounds; r++) {
      const std::array<Fr, 4> e = prover.computeRoundEvals(r, claim);
      for (int j = 0; j < 4; j++) e[j].store(evals_out + 16 * r + 4 * j);
      const Fr ch = Fr::fromRaw(challenges + 4 * r);
      claim = evaluateCubicAtPoint(e, ch);
      prover.bindChallenge(r, ch);
    }
    if (claim_out) claim.store(claim_out + 4);
  } else {
    Transcript tr(std::string_view(domain, dn));
    KeccakAdapter ad{tr};
    if (claim_out) prover.computeInputClaim().store(claim_out);
    Stage4GruenProver::Result res = prover.prove(ad);
    for (size_t r = 0; r < rounds; r++) {
      for (int j = 0; j < 4; j++) res.round_polys[r][j].store(evals_out + 16 * r + 4 * j);
      if (challenges_out) res.challenges[r].store(challenges_out + 4 * r);
    }
  }
  if (finals_out) {
    const Stage4GruenProver::FinalClaims f = prover.getFinalClaims();
    f.val_claim.store(finals_out);
    f.rs1_ra_claim.store(finals_out + 4);
    f.rs2_ra_claim.store(finals_out + 8);
    f.rd_wa_claim.store(finals_out + 12);
    f.inc_claim.store(finals_out + 16);
    prover.finalEq().store(finals_out + 20);
  }
  return ZOLT_OK;
}

// Stage 3 (stage3_prover.zig): the three provers over device-resident trace columns, driven by the caller's challenges like
// generateStage3Proof's loop (:393-560).  cols: 16 device pointers -- shift (unexpanded_pc, pc, is_virtual,
// is_first_in_sequence, is_noop), instr (left_is_rs1, rs1_value, left_is_pc, unexpanded_pc, right_is_rs2, rs2_value,
// right_is_imm, imm), reg (rd_write_value, rs1_value, rs2_value).  gammas: shift gamma_powers[0..4], instr gamma, reg gamma.
// evals_out: rounds x 14 field elements (shift 3, instr 4, reg 3, combined 4); claims_out: the 3 input claims then the 3
// claims after the last round; finals_out: the 16 final claims in the order of cols.
int zolt_host_stage3_run(zolt_ctx* ctx, int log_T, const void* const* cols, const uint64_t* r_outer, const uint64_t* r_product,
                         const uint64_t* gammas, const uint64_t* batching, const uint64_t* challenges, uint64_t* evals_out,
                         uint64_t* claims_out, uint64_t* finals_out) {
  using namespace stage3;
  Context c(ctx);
  std::vector<Fr> ro(log_T), rp(log_T), gp(5);
  for (int i = 0; i < log_T; i++) {
    ro[i] = Fr::fromRaw(r_outer + 4 * i);
    rp[i] = Fr::fromRaw(r_product + 4 * i);
  }
  for (int i = 0; i < 5; i++) gp[i] = Fr::fromRaw(gammas + 4 * i);
  const Fr ig = Fr::fromRaw(gammas + 20), rg = Fr::fromRaw(gammas + 24);
  ShiftPrefixSuffixProver shift(c, log_T, {cols[0], cols[1], cols[2], cols[3], cols[4]}, ro, rp, gp);
  InstructionInputProver instr(c, log_T, {cols[5], cols[6], cols[7], cols[8], cols[9], cols[10], cols[11], cols[12]}, ro, rp, ig, ig.mul(ig));
  RegistersPrefixSuffixProver reg(c, log_T, {cols[13], cols[14], cols[15]}, ro, rg, rg.mul(rg));
  Provers all{shift, instr, reg, {shift.inputClaim(), instr.inputClaim(), reg.inputClaim()}};
  for (int i = 0; i < 3; i++) all.claims[i].store(claims_out + 4 * i);
  const std::array<Fr, 3> bc = {Fr::fromRaw(batching), Fr::fromRaw(batching + 4), Fr::fromRaw(batching + 8)};
  for (int round = 0; round < log_T; round++) {
    const RoundEvals e = all.computeRound(bc);
    uint64_t* o = evals_out + (size_t)round * 14 * 4;
    for (int i = 0; i < 3; i++) e.shift[i].store(o + 4 * i);
    for (int i = 0; i < 4; i++) e.instr[i].store(o + 12 + 4 * i);
    for (int i = 0; i < 3; i++) e.reg[i].store(o + 28 + 4 * i);
    for (int i = 0; i < 4; i++) e.combined[i].store(o + 40 + 4 * i);
    all.bind(e, Fr::fromRaw(challenges + 4 * round));
  }
  for (int i = 0; i < 3; i++) all.claims[i].store(claims_out + 12 + 4 * i);
  const ShiftPrefixSuffixProver::FinalClaims sf = shift.finalClaims();
  const InstructionInputProver::FinalClaims inf = instr.finalClaims();
  const RegistersPrefixSuffixProver::FinalClaims rf = reg.finalClaims();
  const Fr fin[16] = {sf.unexpanded_pc, sf.pc, sf.is_virtual, sf.is_first_in_sequence, sf.is_noop,
                      inf.left_is_rs1, inf.rs1_value, inf.left_is_pc, inf.unexpanded_pc, inf.right_is_rs2, inf.rs2_value, inf.right_is_imm, inf.imm,
                      rf.rd_write_value, rf.rs1_value, rf.rs2_value};
  for (int i = 0; i < 16; i++) fin[i].store(finals_out + 4 * i);
  return ZOLT_OK;
}

// LassoProver over device-resident lookup indices (n x u128).  challenges != NULL: the caller's loop (computeRoundPolynomial /
// receiveChallenge, "lasso prover claim tracking"); NULL: runLassoProver with the reference's deriveChallenge, challenges_out
// filled.  coeffs_out: rounds x 3; claims_out: rounds + 1 (the claim before every round, then the last); final_out: final_eval.
int zolt_host_lasso_run(zolt_ctx* ctx, const void* d_indices, size_t num_lookups, int log_T, int log_K, const uint64_t* r_reduction,
                        const uint64_t* challenges, uint64_t* coeffs_out, uint64_t* claims_out, uint64_t* challenges_out,
                        uint64_t* final_out) {
  Context c(ctx);
  LassoParams params;
  params.gamma = Fr::zero();
  params.log_T = (size_t)log_T;
  params.log_K = (size_t)log_K;
  for (int i = 0; i < log_T; i++) params.r_reduction.push_back(Fr::fromRaw(r_reduction + 4 * i));
  LassoProver prover(c, d_indices, num_lookups, params);
  size_t round = 0;
  while (!prover.isComplete()) {
    prover.current_claim.store(claims_out + 4 * round);
    const std::array<Fr, 3> p = prover.computeRoundPolynomial();
    for (int j = 0; j < 3; j++) p[j].store(coeffs_out + 12 * round + 4 * j);
    const Fr ch = challenges ? Fr::fromRaw(challenges + 4 * round) : lassoDeriveChallenge(p, round);
    if (challenges_out) ch.store(challenges_out + 4 * round);
    prover.receiveChallenge(ch);
    round++;
  }
  prover.current_claim.store(claims_out + 4 * round);
  prover.getFinalEval().store(final_out);
  return ZOLT_OK;
}

// SRS container parsers: kind 0 raw binary (srs.zig:256-306), 1 snarkjs PTAU (:733-900), 2 Jolt Dory SRS (dory.zig:740-826).
// Returns the srs::Error code (0 = ok); *bases_out is a zolt_bases handle the caller frees.
int zolt_host_srs_load(zolt_ctx* ctx, int kind, const uint8_t* data, size_t len, zolt_bases** bases_out, size_t* count,
                       uint64_t* max_num_vars, size_t* bad_index) {
  Context c(ctx);
  srs::G1Srs s;
  const srs::Error e = kind == 0 ? srs::loadFromRawBinary(c, data, len, &s) : kind == 1 ? srs::loadFromPtau(c, data, len, &s)
                                                                                        : srs::loadDorySrs(c, data, len, &s);
  *bases_out = s.powers;
  *count = s.count;
  *max_num_vars = s.max_num_vars;
  *bad_index = s.bad_index;
  return (int)e;
}

// host glue parity hooks: EqPolynomial.mle, BatchOps.hornerEval, DensePolynomial.evaluate
void zolt_host_eq_mle(const uint64_t* r, const uint64_t* x, size_t n, uint64_t* out) {
  std::vector<Fr> rv(n), xv(n);
  for (size_t i = 0; i < n; i++) { rv[i] = Fr::fromRaw(r + 4 * i); xv[i] = Fr::fromRaw(x + 4 * i); }
  EqPolynomial::mle(rv, xv).store(out);
}
void zolt_host_horner_eval(const uint64_t* coeffs, size_t n, const uint64_t* x, uint64_t* out) {
  std::vector<Fr> cv(n);
  for (size_t i = 0; i < n; i++) cv[i] = Fr::fromRaw(coeffs + 4 * i);
  BatchOps::hornerEval(cv, Fr::fromRaw(x)).store(out);
}
int zolt_host_dense_evaluate(zolt_ctx* ctx, const uint64_t* evals, size_t n, const uint64_t* point, size_t n_vars, uint64_t* out) {
  Context c(ctx);
  std::vector<Fr> ev(n), pt(n_vars);
  for (size_t i = 0; i < n; i++) ev[i] = Fr::fromRaw(evals + 4 * i);
  for (size_t i = 0; i < n_vars; i++) pt[i] = Fr::fromRaw(point + 4 * i);
  DensePolynomial poly(c, ev);
  poly.evaluate(pt).store(out);
  return ZOLT_OK;
}

// One rank of the sharded dense sumcheck (contiguous hypercube blocks, NVLink mailbox exchange inside the round
// kernels).  evals_out / challenges_out hold local_log + log2(world) rounds.
int zolt_host_sumcheck_prove_sharded_claim(zolt_ctx* ctx, zolt_peer_box* box, int world, int combiner, int local_log, int n_tables,
                                           const void* const* d_tables, const uint64_t* claim, const char* domain, size_t dn,
                                           const char* label, size_t ln, uint64_t* evals_out, uint64_t* challenges_out,
                                           uint64_t* final_out) {
  Context c(ctx);
  std::vector<const void*> tabs(d_tables, d_tables + n_tables);
  Transcript tr(std::string_view(domain, dn));
  std::vector<Fr> fin;
  Fr cl;
  if (claim) cl = Fr::fromRaw(claim);
  SumcheckProof proof = proveSumcheckSharded(c, box, world, combiner, tabs, local_log, tr, std::string_view(label, ln), &fin, claim ? &cl : nullptr);
  store_proof(proof, evals_out, challenges_out);
  if (final_out)
    for (int t = 0; t < n_tables; t++) fin[t].store(final_out + 4 * t);
  return ZOLT_OK;
}

int zolt_host_sumcheck_prove_sharded(zolt_ctx* ctx, zolt_peer_box* box, int world, int combiner, int local_log, int n_tables,
                                     const void* const* d_tables, const char* domain, size_t dn, const char* label, size_t ln,
                                     uint64_t* evals_out, uint64_t* challenges_out, uint64_t* final_out) {
  Context c(ctx);
  std::vector<const void*> tabs(d_tables, d_tables + n_tables);
  Transcript tr(std::string_view(domain, dn));
  std::vector<Fr> fin;
  SumcheckProof proof = proveSumcheckSharded(c, box, world, combiner, tabs, local_log, tr, std::string_view(label, ln), &fin);
  store_proof(proof, evals_out, challenges_out);
  if (final_out)
    for (int t = 0; t < n_tables; t++) fin[t].store(final_out + 4 * t);
  return ZOLT_OK;
}

// BatchedSumcheckProver over n_inst dense instances (host tables, adjacent pairs) with caller-supplied batching
// coefficients, input claims and per-round challenges.  Outputs per round the compressed polynomial [c0, c2, c3]
// (generateBatchedProof, batched_sumcheck.zig:330-420: c1 is recovered from the running claim, the claim is updated from
// the re-expanded evaluations) and the final claim.
int zolt_host_batched_sumcheck(zolt_ctx* ctx, int n_inst, const int* combiners, const int* log_ns, const int* n_tables,
                               const uint64_t* const* tables /* all instances, concatenated */, const uint64_t* claims,
                               const uint64_t* coeffs, const uint64_t* challenges, uint64_t* compressed_out /* rounds*3*4 */,
                               uint64_t* final_claim_out) {
  Context c(ctx);
  std::vector<std::unique_ptr<DenseSumcheckProver>> provers;
  BatchedSumcheckProver batch;
  size_t t0 = 0;
  for (int i = 0; i < n_inst; i++) {
    std::vector<const uint64_t*> tabs(tables + t0, tables + t0 + n_tables[i]);
    t0 += (size_t)n_tables[i];
    provers.push_back(std::make_unique<DenseSumcheckProver>(c, combiners[i], ZOLT_PAIR_ADJACENT, tabs, log_ns[i]));
    batch.addInstance(provers.back().get(), Fr::fromRaw(claims + 4 * i));
  }
  std::vector<Fr> cf;
  for (int i = 0; i < n_inst; i++) cf.push_back(Fr::fromRaw(coeffs + 4 * i));
  batch.setBatchingCoeffs(cf);
  for (size_t round = 0; round < batch.maxNumRounds(); round++) {
    const auto comp = batch.computeRoundPolynomial();
    for (int j = 0; j < 3; j++) comp[j].store(compressed_out + 12 * round + 4 * j);
    const Fr r = Fr::fromRaw(challenges + 4 * round);
    const Fr c0 = comp[0], c2 = comp[1], c3 = comp[2];
    const Fr c1 = batch.current_claim.sub(c0).sub(c0).sub(c2).sub(c3);  // :389
    const Fr s1 = c0.add(c1).add(c2).add(c3);
    const Fr s2 = c0.add(c1.mul(Fr::fromU64(2))).add(c2.mul(Fr::fromU64(4))).add(c3.mul(Fr::fromU64(8)));
    const Fr s3 = c0.add(c1.mul(Fr::fromU64(3))).add(c2.mul(Fr::fromU64(9))).add(c3.mul(Fr::fromU64(27)));
    batch.updateClaim({c0, s1, s2, s3}, r);
    batch.bindChallenge(r);
  }
  batch.current_claim.store(final_claim_out);
  return ZOLT_OK;
}

// Gruen split-eq degree-3 sumcheck (ProductVirtualRemainderProver shape); claim = initial claim.
int zolt_host_sumcheck_prove_gruen(zolt_ctx* ctx, int with_c, int log_n, int n_tables, const uint64_t* const* tables,
                                   const uint64_t* tau, const uint64_t* claim, const char* domain, size_t dn,
                                   const char* label, size_t ln, uint64_t* evals_out, uint64_t* challenges_out,
                                   uint64_t* final_claim_out) {
  Context c(ctx);
  std::vector<const uint64_t*> tabs(tables, tables + n_tables);
  std::vector<Fr> tv(log_n);
  for (int i = 0; i < log_n; i++) tv[i] = Fr::fromRaw(tau + 4 * i);
  GruenProductProver prover(c, tabs, log_n, tv, Fr::fromRaw(claim), with_c != 0);
  Transcript tr(std::string_view(domain, dn));
  SumcheckProof proof = proveSumcheck(prover, tr, std::string_view(label, ln));
  store_proof(proof, evals_out, challenges_out);
  if (final_claim_out) prover.current_claim.store(final_claim_out);
  return ZOLT_OK;
}

// HyperKZG.commit through the mirror (params = device bases handle)
int zolt_host_hyperkzg_commit(zolt_ctx* ctx, zolt_bases* params, const uint64_t* evals, size_t n, uint64_t* out_xy,
                              uint8_t* out_inf) {
  size_t len = 0;
  must(zolt_msm_bases_len(params, &len), "zolt_msm_bases_len");
  const size_t m = n < len ? n : len;
  if (m == 0) {
    std::memset(out_xy, 0, 64);
    *out_inf = 1;
    return ZOLT_OK;
  }
  return zolt_msm(ctx, params, 0, evals, m, out_xy, out_inf);
}

}  // extern "C"
