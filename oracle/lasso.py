"""ORACLE (test infrastructure only).  Big-integer restatement of Zolt's Lasso lookup prover, loop for loop.  Values are
STANDARD integers mod r.

Reference anchors (relative to /root/reference):
  LassoProver                  src/zkvm/lasso/prover.zig:82-470   init :121-200, computeAddressRoundPoly :265-310,
      computeCycleRoundPoly :312-364, receiveChallenge :366-456, getFinalEval :458-463
  runLassoProver / deriveChallenge   :495-551 (the reference's own stand-in for Fiat-Shamir: a 64-bit mix of the coefficients)
  SplitEqPolynomial.buildTables / getEq   src/zkvm/lasso/split_eq.zig:120-167 (variable i of each half <-> bit i of its index)
  ExpandingTable.bind / get    src/zkvm/lasso/expanding_table.zig:83-124
The prefix-suffix decompositions and the split eq's bindOuter / bindInner the prover also advances every round never reach
its outputs (round polynomials, claims, final_eval, challenges) and are not restated.
Pinned by the reference's unit tests (prover.zig "lasso prover claim tracking" :620-690, integration_test.zig "lasso
end-to-end simple case" :68-120 and "lasso multiple rounds consistent" :282-338), whose inputs and assertions
tests/test_oracle_lasso.py replays, and by the 24 round tuples of the reference's own captured run (logs/zolt.log ->
tests/golden/zolt_log_kats.json): claim = p(0) + p(1), the re-summed claim after a challenge = p(challenge), c2 = 0.
"""
from __future__ import annotations

from .bn254 import FR_MOD as P

MASK64 = (1 << 64) - 1


def split_eq_tables(num_outer: int, num_inner: int, w: list[int]):
    """SplitEqPolynomial.buildTables (split_eq.zig:120-142)."""
    def build(ws):
        t = [1]
        for i, wi in enumerate(ws):
            half = 1 << i
            t = t + [0] * half
            for j in range(half):
                t[j + half] = t[j] * wi % P
                t[j] = t[j] * ((1 - wi) % P) % P
        return t
    return build(w[:num_outer]), build(w[num_outer:num_outer + num_inner])


def derive_challenge(coeffs_mont_limbs: list[list[int]], rnd: int) -> int:
    """deriveChallenge (prover.zig:537-551) over the MONTGOMERY limbs of the coefficients; returns the u64 fed to F.fromU64."""
    h = 0x9E3779B97F4A7C15
    h ^= rnd
    h = h * 0xFF51AFD7ED558CCD & MASK64
    for limbs in coeffs_mont_limbs:
        for limb in limbs:
            h ^= limb
            h = h * 0xC4CEB9FE1A85EC53 & MASK64
    h ^= h >> 33
    return h


class LassoProver:
    def __init__(self, lookup_indices: list[int], log_T: int, log_K: int, r_reduction: list[int], full_expanding_table: bool = True):
        # full_expanding_table = False keeps only entry 0 of the ExpandingTable (the one getFinalEval reads): the reference
        # allocates all 2^log_K entries, which no machine holds at the 64- / 128-bit index widths the indices allow
        self.full_expanding = full_expanding_table
        self.idx = list(lookup_indices)
        self.log_T, self.log_K = log_T, log_K
        outer = log_T // 2
        inner = log_T - outer
        E_out, E_in = split_eq_tables(outer, inner, [x % P for x in r_reduction])
        padded = 1 << log_T
        mask = (1 << inner) - 1
        self.eq_evals = [(E_out[j >> inner] * E_in[j & mask] % P) if j < len(self.idx) else 0 for j in range(padded)]
        self.eq_len = padded
        self.current_claim = sum(self.eq_evals) % P
        self.round = 0
        self.expanding = [1]
        self.challenges: list[int] = []

    def is_address_phase(self) -> bool:
        return self.round < self.log_K

    def is_complete(self) -> bool:
        return self.round >= self.log_K + self.log_T

    def compute_round_polynomial(self) -> list[int]:
        """coefficients [c0, c1, c2]"""
        if self.is_address_phase():
            s = [0, 0]
            for j, idx in enumerate(self.idx):
                b = (idx >> self.round) & 1
                s[b] = (s[b] + self.eq_evals[j]) % P
            return [s[0], (s[1] - s[0]) % P, 0]
        n = self.eq_len
        if n <= 1:
            return [self.eq_evals[0] if n > 0 else 0, 0, 0]
        half = n // 2
        s0 = sum(self.eq_evals[:half]) % P
        s1 = sum(self.eq_evals[half:n]) % P
        return [s0, (s1 - s0) % P, 0]

    def receive_challenge(self, c: int):
        c %= P
        self.challenges.append(c)
        if self.is_address_phase():
            om = (1 - c) % P
            for j, idx in enumerate(self.idx):
                self.eq_evals[j] = self.eq_evals[j] * (c if (idx >> self.round) & 1 else om) % P
            self.current_claim = sum(self.eq_evals) % P
            self.expanding = [v * f % P for v in self.expanding for f in (om, c)] if self.full_expanding else [self.expanding[0] * om % P]
        else:
            n = self.eq_len
            if n <= 1:
                self.round += 1
                return
            half = n // 2
            om = (1 - c) % P
            for j in range(half):
                e1 = self.eq_evals[j + half] if j + half < n else 0
                self.eq_evals[j] = (om * self.eq_evals[j] + c * e1) % P
            self.eq_len = half
            self.current_claim = sum(self.eq_evals[:half]) % P
        self.round += 1

    def final_eval(self) -> int:
        assert self.is_complete()
        return self.expanding[0]


def run_lasso_prover(lookup_indices, log_T, log_K, r_reduction, to_mont_limbs, full_expanding_table: bool = True):
    """runLassoProver (:495-535); to_mont_limbs(int) -> the four 64-bit Montgomery limbs the reference hashes."""
    pr = LassoProver(lookup_indices, log_T, log_K, r_reduction, full_expanding_table)
    polys = []
    rnd = 0
    while not pr.is_complete():
        c = pr.compute_round_polynomial()
        polys.append(c)
        pr.receive_challenge(derive_challenge([to_mont_limbs(x) for x in c], rnd))
        rnd += 1
    return polys, pr.final_eval(), pr.challenges
