"""ORACLE (test infrastructure only).  Big-integer restatement of Zolt's Stage 3 provers, loop for loop.  Values are STANDARD
integers mod r.

Reference anchors (relative to /root/reference):
  EqPlusOnePolynomial.mle                      src/poly/mod.zig:387-421 (unit expectations :890-960)
  Stage3Prover.generateStage3Proof (round loop) src/zkvm/spartan/stage3_prover.zig:113-760
  ShiftPrefixSuffixProver                       :928-1920   init :979-1112 (P / Q buffers), computeRoundEvalsPhase1 :1351-1397,
      computeRoundEvalsPhase2 :1399-1456, bind :1458-1472, bindPhase1 :1479-1504, transitionToPhase2 :1506-1707,
      bindPhase2 :1782-1818, evaluateMle :1820-1838, computeEqEvals :1840-1858, computeEqPlusOneEvals :1878-1895
  InstructionInputProver                        :1921-2150  computeRoundEvals :2029-2105, bind :2107-2126
  RegistersPrefixSuffixProver                   :2156-2496  init :2189-2314, phases :2334-2386, bind :2388-2481
The provers are entered with the trace COLUMNS they read out of cycle_witnesses (the R1CS witness generation is out of
scope): lists of length trace_len = 2^n.  Parity is pinned by the reference's unit expectations for eq+1, by the
sumcheck identities the reference's own debug lines assert (p(0) + p(1) = claim every round, grand_sum(P*Q) = direct_sum,
the final-claim relations of :620-700) and by what the reference's own captured run printed for this stage (logs/zolt.log,
committed as tests/golden/zolt_log_kats.json: the eight STAGE3_ROUND_k tuples chain through the cubic evaluation used here,
shift_p0 + shift_p1 of round 0 equals input_claim[0], gamma_powers[4] = gamma_powers[1]^4, and eq+1(r_outer, k) for
k = 0..4, 255 plus eq(r_outer, 255) equal the printed values for r_outer = that run's Stage 1 challenges; trace_len 256 -> prefix_size 16,
suffix_size 16).  The run's witness columns are not in the log, so the provers' per-round values are pinned through these
identities, not value by value.
"""
from __future__ import annotations

from .bn254 import FR_MOD as P
from .poly import eq_evals, eq_mle


def eq_plus_one_mle(x: list[int], y: list[int]) -> int:
    """src/poly/mod.zig:387-421, both points big-endian."""
    l = len(x)
    assert len(y) == l
    result = 0
    for k in range(l):
        lower = 1
        for i in range(k):
            idx = l - 1 - i
            lower = lower * (x[idx] * ((1 - y[idx]) % P) % P) % P
        kth = l - 1 - k
        kth_prod = ((1 - x[kth]) % P) * y[kth] % P
        higher = 1
        for i in range(k + 1, l):
            idx = l - 1 - i
            higher = higher * ((x[idx] * y[idx] + ((1 - x[idx]) % P) * ((1 - y[idx]) % P)) % P) % P
        result = (result + lower * kth_prod % P * higher) % P
    return result


def _bits_be(j: int, n: int) -> list[int]:
    return [(j >> (n - 1 - k)) & 1 for k in range(n)]


def eq_plus_one_evals(r: list[int]) -> list[int]:
    """computeEqPlusOneEvals (:1878-1895): one mle per hypercube point."""
    n = len(r)
    return [eq_plus_one_mle(r, _bits_be(j, n)) for j in range(1 << n)]


def eq_evals_pointwise(r: list[int]) -> list[int]:
    """computeEqEvals (:1840-1858)."""
    n = len(r)
    return [eq_mle(r, _bits_be(j, n)) for j in range(1 << n)]


def _bind(t: list[int], r: int) -> list[int]:
    return [(t[2 * i] + r * ((t[2 * i + 1] - t[2 * i]) % P)) % P for i in range(len(t) // 2)]


def evaluate_mle_le(coeffs: list[int], point: list[int]) -> int:
    """evaluateMle (:1820-1838): point[0] binds the LSB."""
    if len(coeffs) == 1 or not point:
        return coeffs[0]
    t = list(coeffs)
    for r in point:
        t = _bind(t, r)
    return t[0]


class InstructionInputProver:
    COLS = ("left_is_rs1", "rs1_value", "left_is_pc", "unexpanded_pc", "right_is_rs2", "rs2_value", "right_is_imm", "imm")

    def __init__(self, cols: dict[str, list[int]], r_outer, r_product, gamma):
        self.t = {k: [v % P for v in cols[k]] for k in self.COLS}
        self.eq_outer = eq_evals(r_outer)
        self.eq_product = eq_evals(r_product)
        self.gamma = gamma % P
        self.gamma_sqr = gamma * gamma % P
        self.size = len(self.eq_outer)

    def compute_round_evals(self, previous_claim: int) -> list[int]:
        t, g, g2 = self.t, self.gamma, self.gamma_sqr
        evals = [0, 0, 0]
        for j in range(self.size // 2):
            def at(name, pt):
                a0, a1 = name[2 * j], name[2 * j + 1]
                if pt == 0:
                    return a0
                a2 = (a1 + a1 - a0) % P
                return a2 if pt == 2 else (a2 + a1 - a0) % P
            for slot, pt in enumerate((0, 2, 3)):
                left = (at(t["left_is_rs1"], pt) * at(t["rs1_value"], pt) + at(t["left_is_pc"], pt) * at(t["unexpanded_pc"], pt)) % P
                right = (at(t["right_is_rs2"], pt) * at(t["rs2_value"], pt) + at(t["right_is_imm"], pt) * at(t["imm"], pt)) % P
                w = (at(self.eq_outer, pt) + g2 * at(self.eq_product, pt)) % P
                evals[slot] = (evals[slot] + w * ((right + g * left) % P)) % P
        return [evals[0], (previous_claim - evals[0]) % P, evals[1], evals[2]]

    def bind(self, r: int):
        for k in self.COLS:
            self.t[k] = _bind(self.t[k], r)
        self.eq_outer = _bind(self.eq_outer, r)
        self.eq_product = _bind(self.eq_product, r)
        self.size //= 2

    def input_claim(self) -> int:
        t = self.t
        s = 0
        for i in range(self.size):
            left = (t["left_is_rs1"][i] * t["rs1_value"][i] + t["left_is_pc"][i] * t["unexpanded_pc"][i]) % P
            right = (t["right_is_rs2"][i] * t["rs2_value"][i] + t["right_is_imm"][i] * t["imm"][i]) % P
            s = (s + (self.eq_outer[i] + self.gamma_sqr * self.eq_product[i]) % P * ((right + self.gamma * left) % P)) % P
        return s

    def final_claims(self) -> list[int]:
        return [self.t[k][0] for k in self.COLS]


class ShiftPrefixSuffixProver:
    COLS = ("unexpanded_pc", "pc", "is_virtual", "is_first_in_sequence", "is_noop")

    def __init__(self, cols: dict[str, list[int]], r_outer, r_product, gamma_powers):
        n_vars = len(r_outer)
        split = n_vars // 2
        self.r_outer, self.r_product = list(r_outer), list(r_product)
        self.gp = [g % P for g in gamma_powers]
        self.w = {k: [v % P for v in cols[k]] for k in self.COLS}
        self.trace_len = len(self.w["pc"])
        self.prefix_n_vars = n_vars - split
        self.suffix_n_vars = split
        ps, ss = 1 << self.prefix_n_vars, 1 << self.suffix_n_vars
        r_o_hi, r_o_lo = self.r_outer[:split], self.r_outer[split:]
        r_p_hi, r_p_lo = self.r_product[:split], self.r_product[split:]

        def is_max(r):
            m = 1
            for x in r:
                m = m * x % P
            return m
        self.P = [eq_plus_one_evals(r_o_lo), [is_max(r_o_lo)] + [0] * (ps - 1), eq_plus_one_evals(r_p_lo), [is_max(r_p_lo)] + [0] * (ps - 1)]
        suf = [eq_evals_pointwise(r_o_hi), eq_plus_one_evals(r_o_hi), eq_evals_pointwise(r_p_hi), eq_plus_one_evals(r_p_hi)]
        w, gp = self.w, self.gp
        self.Q = [[0] * ps for _ in range(4)]
        for x_lo in range(ps):
            acc = [0, 0, 0, 0]
            for x_hi in range(ss):
                x = x_lo + (x_hi << self.prefix_n_vars)
                if x >= self.trace_len:
                    continue
                v = (w["unexpanded_pc"][x] + gp[1] * w["pc"][x] + gp[2] * w["is_virtual"][x] + gp[3] * w["is_first_in_sequence"][x]) % P
                nn = (1 - w["is_noop"][x]) % P
                acc[0] = (acc[0] + v * suf[0][x_hi]) % P
                acc[1] = (acc[1] + v * suf[1][x_hi]) % P
                acc[2] = (acc[2] + nn * suf[2][x_hi]) % P
                acc[3] = (acc[3] + nn * suf[3][x_hi]) % P
            self.Q[0][x_lo], self.Q[1][x_lo] = acc[0], acc[1]
            self.Q[2][x_lo], self.Q[3][x_lo] = acc[2] * gp[4] % P, acc[3] * gp[4] % P
        self.prefix_size = ps
        self.challenges: list[int] = []
        self.in_phase2 = False
        self.eq1_outer = self.eq1_prod = None

    def input_claim(self) -> int:
        """grand_sum(P*Q) (:1131-1139)."""
        return sum(p * q for Pb, Qb in zip(self.P, self.Q) for p, q in zip(Pb, Qb)) % P

    def compute_round_evals(self, previous_claim: int) -> list[int]:
        if not self.in_phase2:
            e = [0, 0, 0]
            for Pb, Qb in zip(self.P, self.Q):
                for i in range(self.prefix_size // 2):
                    p0, p1, q0, q1 = Pb[2 * i], Pb[2 * i + 1], Qb[2 * i], Qb[2 * i + 1]
                    p2, q2 = (p1 + p1 - p0) % P, (q1 + q1 - q0) % P
                    e[0] = (e[0] + p0 * q0) % P
                    e[1] = (e[1] + p1 * q1) % P
                    e[2] = (e[2] + p2 * q2) % P
            return e
        w, gp = self.w, self.gp
        e0 = e2 = 0
        for j in range(len(self.eq1_outer) // 2):
            def pair(t):
                a0, a1 = t[2 * j], t[2 * j + 1]
                return a0, (a1 + a1 - a0) % P
            eo, ep = pair(self.eq1_outer), pair(self.eq1_prod)
            cols = {k: pair(w[k]) for k in self.COLS}
            for slot in (0, 1):
                val = (cols["unexpanded_pc"][slot] + gp[1] * cols["pc"][slot] + gp[2] * cols["is_virtual"][slot]
                       + gp[3] * cols["is_first_in_sequence"][slot]) % P
                f = (eo[slot] * val + gp[4] * ((1 - cols["is_noop"][slot]) % P) % P * ep[slot]) % P
                if slot == 0:
                    e0 = (e0 + f) % P
                else:
                    e2 = (e2 + f) % P
        return [e0, (previous_claim - e0) % P, e2]

    def bind(self, r: int):
        if self.in_phase2:
            for k in self.COLS:
                self.w[k] = _bind(self.w[k], r)
            self.eq1_outer = _bind(self.eq1_outer, r)
            self.eq1_prod = _bind(self.eq1_prod, r)
            return
        last = self.prefix_size == 2
        self.P = [_bind(b, r) for b in self.P]
        self.Q = [_bind(b, r) for b in self.Q]
        self.prefix_size //= 2
        self.challenges.append(r)
        if last:
            self._transition()

    def _transition(self):
        self.in_phase2 = True
        r_le = self.challenges
        r_be = r_le[::-1]
        sv, ss = self.suffix_n_vars, 1 << self.suffix_n_vars
        r_o_hi, r_o_lo = self.r_outer[:sv], self.r_outer[sv:]
        r_p_hi, r_p_lo = self.r_product[:sv], self.r_product[sv:]

        def prefix_evals(r_lo):
            p0 = eq_plus_one_evals(r_lo)
            m = 1
            for x in r_lo:
                m = m * x % P
            p1 = [m] + [0] * (len(p0) - 1)
            return evaluate_mle_le(p0, r_le), evaluate_mle_le(p1, r_le)
        p0o, p1o = prefix_evals(r_o_lo)
        p0p, p1p = prefix_evals(r_p_lo)
        s0o, s1o = eq_evals_pointwise(r_o_hi), eq_plus_one_evals(r_o_hi)
        s0p, s1p = eq_evals_pointwise(r_p_hi), eq_plus_one_evals(r_p_hi)
        self.eq1_outer = [(p0o * s0o[j] + p1o * s1o[j]) % P for j in range(ss)]
        self.eq1_prod = [(p0p * s0p[j] + p1p * s1p[j]) % P for j in range(ss)]
        eqp = eq_evals_pointwise(r_be)
        pd = len(eqp)
        new = {k: [0] * ss for k in self.COLS}
        for j in range(ss):
            for i in range(pd):
                idx = j * pd + i
                if idx >= self.trace_len:
                    continue
                for k in self.COLS:
                    new[k][j] = (new[k][j] + eqp[i] * self.w[k][idx]) % P
        self.w = new

    def final_claims(self) -> list[int]:
        return [self.w[k][0] for k in self.COLS]


class RegistersPrefixSuffixProver:
    COLS = ("rd_write_value", "rs1_value", "rs2_value")

    def __init__(self, cols: dict[str, list[int]], r_spartan, gamma):
        n_vars = len(r_spartan)
        split = n_vars // 2
        self.r_hi, self.r_lo = list(r_spartan[:split]), list(r_spartan[split:])
        self.gamma, self.gamma_sqr = gamma % P, gamma * gamma % P
        self.w = {k: [v % P for v in cols[k]] for k in self.COLS}
        self.trace_len = len(self.w["rs1_value"])
        self.prefix_n_vars = n_vars - split
        ps, ss = 1 << self.prefix_n_vars, 1 << split
        self.P = eq_evals(self.r_lo)
        suffix = eq_evals(self.r_hi)
        self.Q = [0] * ps
        for x_lo in range(ps):
            acc = 0
            for x_hi in range(ss):
                x = x_lo + (x_hi << self.prefix_n_vars)
                if x >= self.trace_len:
                    continue
                v = (self.w["rd_write_value"][x] + self.gamma * self.w["rs1_value"][x] + self.gamma_sqr * self.w["rs2_value"][x]) % P
                acc = (acc + v * suffix[x_hi]) % P
            self.Q[x_lo] = acc
        self.prefix_size = ps
        self.in_phase2 = False
        self.prefix_challenges: list[int] = []
        self.eq = None

    def input_claim(self) -> int:
        return sum(p * q for p, q in zip(self.P, self.Q)) % P

    def compute_round_evals(self, previous_claim: int) -> list[int]:
        e0 = e2 = 0
        if not self.in_phase2:
            for i in range(self.prefix_size // 2):
                p0, p1, q0, q1 = self.P[2 * i], self.P[2 * i + 1], self.Q[2 * i], self.Q[2 * i + 1]
                e0 = (e0 + p0 * q0) % P
                e2 = (e2 + ((p1 + p1 - p0) % P) * ((q1 + q1 - q0) % P)) % P
        else:
            g, g2, w = self.gamma, self.gamma_sqr, self.w
            for j in range(len(self.eq) // 2):
                def pair(t):
                    a0, a1 = t[2 * j], t[2 * j + 1]
                    return a0, (a1 + a1 - a0) % P
                eq, rd, r1, r2 = pair(self.eq), pair(w["rd_write_value"]), pair(w["rs1_value"]), pair(w["rs2_value"])
                e0 = (e0 + eq[0] * ((rd[0] + g * r1[0] + g2 * r2[0]) % P)) % P
                e2 = (e2 + eq[1] * ((rd[1] + g * r1[1] + g2 * r2[1]) % P)) % P
        return [e0, (previous_claim - e0) % P, e2]

    def bind(self, r: int):
        if self.in_phase2:
            for k in self.COLS:
                self.w[k] = _bind(self.w[k], r)
            self.eq = _bind(self.eq, r)
            return
        last = self.prefix_size == 2
        self.P, self.Q = _bind(self.P, r), _bind(self.Q, r)
        self.prefix_size //= 2
        for k in self.COLS:
            self.w[k] = _bind(self.w[k], r)
        self.prefix_challenges.append(r)
        if last:
            self.in_phase2 = True
            eq_prefix_eval = eq_mle(self.r_lo, self.prefix_challenges[::-1])
            self.eq = [e * eq_prefix_eval % P for e in eq_evals(self.r_hi)]

    def final_claims(self) -> list[int]:
        return [self.w[k][0] for k in self.COLS]


def _quad_at(e: list[int], r: int) -> int:
    """evalsToCoeffs(.., 2) + evaluatePolyAtPoint (:846-926): the quadratic through (0,e0) (1,e1) (2,e2) at r."""
    inv2 = pow(2, -1, P)
    c0 = e[0]
    c2 = (e[2] - 2 * e[1] + e[0]) * inv2 % P
    c1 = (e[1] - e[0] - c2) % P
    return (c0 + r * (c1 + r * c2)) % P


def _cubic_at(e: list[int], r: int) -> int:
    from .poly import evaluate_cubic_at_point
    return evaluate_cubic_at_point(e, r)


def run_stage3(shift: ShiftPrefixSuffixProver, instr: InstructionInputProver, reg: RegistersPrefixSuffixProver,
               batching_coeffs: list[int], challenges: list[int]):
    """The round loop of generateStage3Proof (:393-560) with the verifier's challenges supplied: per round the three provers'
    evaluations, the batched cubic [p(0..3)], then the per-instance claims at the challenge and the binds."""
    claims = [shift.input_claim(), instr.input_claim(), reg.input_claim()]
    out = []
    for r in challenges:
        se = shift.compute_round_evals(claims[0])
        ie = instr.compute_round_evals(claims[1])
        re_ = reg.compute_round_evals(claims[2])
        s3 = (3 * se[2] - 3 * se[1] + se[0]) % P
        r3 = (3 * re_[2] - 3 * re_[1] + re_[0]) % P
        s4, r4 = se + [s3], re_ + [r3]
        combined = [(s4[i] * batching_coeffs[0] + ie[i] * batching_coeffs[1] + r4[i] * batching_coeffs[2]) % P for i in range(4)]
        out.append({"shift": se, "instr": ie, "reg": re_, "combined": combined})
        claims = [_quad_at(se, r), _cubic_at(ie, r), _quad_at(re_, r)]
        shift.bind(r)
        instr.bind(r)
        reg.bind(r)
    return out, claims
