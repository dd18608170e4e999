"""ORACLE (test infrastructure only).  Big-integer restatement of Zolt's Stage 4 registers read/write-checking prover in
its Gruen form, loop for loop.  Values are STANDARD integers mod r.

Reference anchors (relative to /root/reference):
  GruenSplitEqPolynomial (Jolt-compatible variant)   src/zkvm/spartan/gruen_eq.zig:19-345
      evalsCached :296-334, E_in_current / E_out_current :102-112, bind :120-160, merge :172-203, gruenPolyDeg3 :214-262,
      fromEvals :265-291
  Stage4GruenProver                                  src/zkvm/spartan/stage4_gruen_prover.zig
      computeInputClaim :385-430, prove :433-536, phase1ComputeMessage :561-757, phase2ComputeMessage :764-851,
      phase3ComputeMessage :854-1015, coeffsFromEvals :1017-1039, bindPolynomials :1047-1160, computeRoundEvals :1165-1190,
      getFinalClaims :1219-1234
The witness generation from an ExecutionTrace (initWithPhaseConfig :139-366) is not restated: the prover is entered with its
dense K x T tables (val, rd_wa, rs1_ra, rs2_ra), inc (T) and the phase configuration, which is the state that function leaves.
Tables are indexed [k * T + j] with the ORIGINAL stride T throughout, like the reference.
Pinned value by value where the reference's captured run allows it: logs/zolt.log prints E_out[0..3], E_in[0..3] and two
E_out * E_in weights of the Stage 4 split eq over r_cycle = that run's Stage 3 challenges, and the cubic gruenPolyDeg3 returned
in its first two rounds from the logged q(0), q_X2 and claim (the second after bind) (tests/golden/zolt_log_kats.json,
tests/test_oracle_stage4.py); the witness tables of the run are not in the log, the rest is pinned through the protocol's
identities and the reference's unit expectations.
"""
from __future__ import annotations

from .bn254 import FR_MOD as P
from .poly import eq_evals

LOG_K = 7
K = 1 << LOG_K


class GruenEq:
    """gruen_eq.zig GruenSplitEqPolynomial; w is big-endian (w[0] = MSB)."""

    def __init__(self, w: list[int], scaling: int = 1):
        n = len(w)
        assert n > 0
        self.w = [x % P for x in w]
        self.current_index = n
        self.current_scalar = scaling % P
        m = n // 2
        self.E_out_vec = self._evals_cached(self.w[:m])
        self.E_in_vec = self._evals_cached(self.w[m:n - 1] if n > m + 1 else [])
        self.E_in_active = len(self.E_in_vec)
        self.E_out_active = len(self.E_out_vec)

    @staticmethod
    def _evals_cached(w):
        tables = [[1]]
        for k, wk in enumerate(w):
            prev = tables[k]
            cur = [0] * (2 * len(prev))
            for i in range(len(cur) - 1, 0, -2):
                s = prev[i // 2]
                cur[i] = s * wk % P
                cur[i - 1] = (s - cur[i]) % P
            tables.append(cur)
        return tables

    def E_in_current(self):
        return [1] if self.E_in_active == 0 else self.E_in_vec[self.E_in_active - 1]

    def E_out_current(self):
        return [1] if self.E_out_active == 0 else self.E_out_vec[self.E_out_active - 1]

    def bind(self, r: int):
        w_i = self.w[self.current_index - 1]
        prod = w_i * r % P
        eq_val = (1 - w_i - r + prod + prod) % P
        self.current_scalar = self.current_scalar * eq_val % P
        self.current_index -= 1
        m = len(self.w) // 2
        if m < self.current_index and self.E_in_active > 1:
            self.E_in_active -= 1
        elif 0 < self.current_index and self.E_out_active > 1:
            self.E_out_active -= 1

    def merge(self):
        if self.current_index == 0:
            return [self.current_scalar]
        return eq_evals(self.w[: self.current_index], self.current_scalar)

    def gruen_poly_deg3(self, q_constant: int, q_quadratic: int, s01: int):
        eq1 = self.current_scalar * self.w[self.current_index - 1] % P
        eq0 = (self.current_scalar - eq1) % P
        m = (eq1 - eq0) % P
        eq2 = (eq1 + m) % P
        eq3 = (eq2 + m) % P
        c0e = eq0 * q_constant % P
        c1e = (s01 - c0e) % P
        q1 = c1e * pow(eq1, -1, P) % P
        e2 = 2 * q_quadratic % P
        q2 = (q1 + q1 - q_constant + e2) % P
        q3 = (q2 + q1 - q_constant + e2 + e2) % P
        return coeffs_from_evals([c0e, c1e, eq2 * q2 % P, eq3 * q3 % P])


def coeffs_from_evals(e):
    """fromEvals / coeffsFromEvals: [p(0..3)] -> [c0, c1, c2, c3]."""
    inv6, inv2 = pow(6, -1, P), pow(2, -1, P)
    c3 = (-e[0] + 3 * e[1] - 3 * e[2] + e[3]) * inv6 % P
    c2 = (2 * e[0] - 5 * e[1] + 4 * e[2] - e[3]) * inv2 % P
    c1 = (e[1] - e[0] - c2 - c3) % P
    return [e[0] % P, c1, c2, c3]


def eval_coeffs(c, x):
    return (c[0] + c[1] * x + c[2] * x * x + c[3] * x * x * x) % P


class Stage4Prover:
    def __init__(self, val, rd_wa, rs1_ra, rs2_ra, inc, gamma, r_cycle, phase1_rounds=None, phase2_rounds=None,
                 stage3_claims=None, batching_coeff=1):
        self.T = len(inc)
        self.log_T = self.T.bit_length() - 1
        assert len(r_cycle) == self.log_T and len(val) == K * self.T
        self.gamma, self.gamma_sq = gamma % P, gamma * gamma % P
        self.val, self.wa = list(val), list(rd_wa)
        self.rs1, self.rs2 = list(rs1_ra), list(rs2_ra)
        self.ra = [(self.gamma * a + self.gamma_sq * b) % P for a, b in zip(rs1_ra, rs2_ra)]  # :211, :224
        self.inc = list(inc)
        self.gruen = GruenEq(list(reversed(r_cycle)))  # LE round order -> BE (:259-281)
        self.stage3_claims = stage3_claims
        self.batching_coeff = batching_coeff % P
        self.p1 = self.log_T // 2 if phase1_rounds is None else phase1_rounds
        self.p2 = LOG_K if phase2_rounds is None else phase2_rounds
        # Configurations the reference itself completes: merged_eq only exists after a phase-1 round (:1083-1086, so
        # phase1 >= 1), and bindPolynomials' third branch never binds an address variable (:1123-1158: with phase2 < LOG_K
        # the leftover address rounds of phase 3 would leave the tables untouched), so phase 2 must take all LOG_K of them.
        assert 1 <= self.p1 <= self.log_T and self.p2 == LOG_K
        self.num_rounds = LOG_K + self.log_T
        self.cur_T, self.cur_K = self.T, K
        self.merged_eq = None

    # ---- claims ------------------------------------------------------------------------------------------------------
    def compute_input_claim(self):
        if self.stage3_claims is not None:
            rd_wv, rs1_v, rs2_v = self.stage3_claims
            return (rd_wv + self.gamma * (rs1_v + self.gamma * rs2_v)) % P
        g = self.gruen
        E_out, E_in = g.E_out_current(), g.E_in_current()
        bits = (len(E_in).bit_length() - 1) if len(E_in) > 0 else 0
        w_last = g.w[g.current_index - 1]
        claim = 0
        T = self.T
        for k in range(K):
            for j in range(T):
                x_last, ju = j & 1, j >> 1
                x_in, x_out = ju & ((1 << bits) - 1), ju >> bits
                eo = E_out[x_out] if x_out < len(E_out) else 1
                ei = E_in[x_in] if x_in < len(E_in) else 1
                lin = (g.current_scalar - g.current_scalar * w_last) % P if x_last == 0 else g.current_scalar * w_last % P
                idx = k * T + j
                comb = (self.ra[idx] * self.val[idx] + self.wa[idx] * (self.val[idx] + self.inc[j])) % P
                claim = (claim + eo * ei % P * lin % P * comb) % P
        return claim

    # ---- round messages ------------------------------------------------------------------------------------------------
    def round_coeffs(self, rnd: int, claim: int):
        if rnd < self.p1:
            return self._phase1(claim)
        if rnd < self.p1 + self.p2:
            return self._phase2(claim)
        return self._phase3(claim)

    def _phase1(self, claim):
        g = self.gruen
        E_in, E_out = g.E_in_current(), g.E_out_current()
        bits = (len(E_in).bit_length() - 1) if len(E_in) > 1 else 0
        mask = (1 << bits) - 1
        T = self.T
        q0 = qx2 = 0
        for i in range(self.cur_T // 2):
            je, jo = 2 * i, 2 * i + 1
            x_in, x_out = (i & mask) if bits > 0 else 0, i >> bits
            e = (E_out[x_out] if x_out < len(E_out) else 1) * (E_in[x_in] if x_in < len(E_in) else 1) % P
            inc0, incs = self.inc[je], (self.inc[jo] - self.inc[je]) % P
            for k in range(self.cur_K):
                a, b = k * T + je, k * T + jo
                ras, was, vs = self.ra[b] - self.ra[a], self.wa[b] - self.wa[a], self.val[b] - self.val[a]
                c0 = (self.ra[a] * self.val[a] + self.wa[a] * (self.val[a] + inc0)) % P
                cx2 = (ras * vs + was * (vs + incs)) % P
                q0 = (q0 + e * c0) % P
                qx2 = (qx2 + e * cx2) % P
        return g.gruen_poly_deg3(q0, qx2, claim)

    def _body(self, ra, wa, val, inc):
        return (ra * val + wa * (val + inc)) % P

    def _phase2(self, claim):
        T = self.T
        e0 = e2 = 0
        for i in range(self.cur_K // 2):
            ke, ko = 2 * i, 2 * i + 1
            for j in range(self.cur_T):
                a, b = ke * T + j, ko * T + j
                inc_j, eq_j = self.inc[j], self.merged_eq[j]
                e0 = (e0 + eq_j * self._body(self.ra[a], self.wa[a], self.val[a], inc_j)) % P
                ra2 = self.ra[a] + 2 * (self.ra[b] - self.ra[a])
                wa2 = self.wa[a] + 2 * (self.wa[b] - self.wa[a])
                v2 = self.val[a] + 2 * (self.val[b] - self.val[a])
                e2 = (e2 + eq_j * self._body(ra2, wa2, v2, inc_j)) % P
        return self._deg2(e0, e2, claim)

    @staticmethod
    def _deg2(e0, e2, claim):
        e1 = (claim - e0) % P
        c2 = (e0 - 2 * e1 + e2) * pow(2, -1, P) % P
        c1 = (e1 - e0 - c2) % P
        return [e0 % P, c1, c2, 0]

    def _phase3(self, claim):
        T = self.T
        if self.cur_T > 1:
            e0 = e2 = e3 = 0
            for j in range(self.cur_T // 2):
                je, jo = 2 * j, 2 * j + 1
                incs = self.inc[jo] - self.inc[je]
                eqs = self.merged_eq[jo] - self.merged_eq[je]
                inc2, inc3 = self.inc[je] + 2 * incs, self.inc[je] + 3 * incs
                eq2, eq3 = self.merged_eq[je] + 2 * eqs, self.merged_eq[je] + 3 * eqs
                i0 = i2 = i3 = 0
                for k in range(self.cur_K):
                    a, b = k * T + je, k * T + jo
                    ras, was, vs = self.ra[b] - self.ra[a], self.wa[b] - self.wa[a], self.val[b] - self.val[a]
                    i0 += self._body(self.ra[a], self.wa[a], self.val[a], self.inc[je])
                    i2 += self._body(self.ra[a] + 2 * ras, self.wa[a] + 2 * was, self.val[a] + 2 * vs, inc2)
                    i3 += self._body(self.ra[a] + 3 * ras, self.wa[a] + 3 * was, self.val[a] + 3 * vs, inc3)
                e0 = (e0 + self.merged_eq[je] * i0) % P
                e2 = (e2 + eq2 * i2) % P
                e3 = (e3 + eq3 * i3) % P
            return coeffs_from_evals([e0, (claim - e0) % P, e2, e3])
        inc_e, eq_e = self.inc[0], self.merged_eq[0]
        e0 = e2 = 0
        for i in range(self.cur_K // 2):
            a, b = 2 * i * T, (2 * i + 1) * T
            e0 += self._body(self.ra[a], self.wa[a], self.val[a], inc_e)
            e2 += self._body(self.ra[a] + 2 * (self.ra[b] - self.ra[a]), self.wa[a] + 2 * (self.wa[b] - self.wa[a]),
                             self.val[a] + 2 * (self.val[b] - self.val[a]), inc_e)
        return self._deg2(eq_e * e0 % P, eq_e * e2 % P, claim)

    def round_evals(self, rnd: int, claim: int):
        """computeRoundEvals :1165-1190."""
        c = self.round_coeffs(rnd, claim)
        return [eval_coeffs(c, x) for x in range(4)]

    # ---- binding -------------------------------------------------------------------------------------------------------
    def bind(self, rnd: int, r: int):
        T = self.T
        tabs = (self.val, self.wa, self.ra, self.rs1, self.rs2)

        def lerp(lo, hi):
            return (lo * (1 - r) + hi * r) % P

        if rnd < self.p1 or rnd >= self.p1 + self.p2:
            h = self.cur_T // 2
            for t in tabs:
                for k in range(self.cur_K):
                    for i in range(h):
                        t[k * T + i] = lerp(t[k * T + 2 * i], t[k * T + 2 * i + 1])
            for i in range(h):
                self.inc[i] = lerp(self.inc[2 * i], self.inc[2 * i + 1])
            if rnd < self.p1:
                self.cur_T = h
                self.gruen.bind(r)
                if rnd == self.p1 - 1:
                    self.merged_eq = self.gruen.merge()
            else:
                for i in range(h):
                    self.merged_eq[i] = lerp(self.merged_eq[2 * i], self.merged_eq[2 * i + 1])
                self.cur_T = h
            return
        h = self.cur_K // 2
        for t in tabs:
            for i in range(h):
                for j in range(self.cur_T):
                    t[i * T + j] = lerp(t[2 * i * T + j], t[(2 * i + 1) * T + j])
        self.cur_K = h

    def final_claims(self):
        return {"val": self.val[0], "rs1_ra": self.rs1[0], "rs2_ra": self.rs2[0], "rd_wa": self.wa[0], "inc": self.inc[0]}

    # ---- whole protocol with caller-supplied challenges (proof_converter.zig:2112-2282 call pattern) --------------------------
    def run(self, challenges, input_claim=None):
        claim = self.compute_input_claim() if input_claim is None else input_claim
        evals = []
        for rnd, r in enumerate(challenges):
            e = self.round_evals(rnd, claim)
            evals.append(e)
            claim = eval_coeffs(coeffs_from_evals(e), r)
            self.bind(rnd, r)
        return evals, claim
