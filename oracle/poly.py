"""ORACLE (test infrastructure only).  Big-integer restatement of Zolt's multilinear-polynomial
helpers.  All values are STANDARD integers mod r (not Montgomery); convert at the boundary with
oracle.bn254.ints_to_mont_array / mont_array_to_ints.

Reference anchors (relative to /root/reference):
  DensePolynomial.bindFirst / bindLow / evaluate   src/poly/mod.zig:128-149, :160-175, :73-92
  EqPolynomial.evalsSliceWithScaling / mle         src/poly/mod.zig:252-290, :311-321
  UniPoly.evaluate / interpolateDegree3 / compress src/poly/mod.zig:608-618, :632-685
  GruenSplitEqPolynomial                           src/poly/split_eq.zig:63-183, :213-248, :311-434
  evaluateCubicAtPointGeneric                      src/zkvm/batched_sumcheck.zig:437-457
"""
from __future__ import annotations

from .bn254 import FR_MOD as P


def bind_low(evals: list[int], r: int) -> list[int]:
    """new[i] = old[2i] + r*(old[2i+1] - old[2i])  (src/poly/mod.zig:165-172)."""
    return [(evals[2 * i] + r * (evals[2 * i + 1] - evals[2 * i])) % P for i in range(len(evals) // 2)]


def bind_first(evals: list[int], r: int) -> list[int]:
    """new[i] = (1-r)*old[i] + r*old[i+half]  (src/poly/mod.zig:137-142)."""
    h = len(evals) // 2
    return [((1 - r) * evals[i] + r * evals[i + h]) % P for i in range(h)]


def mle_evaluate(evals: list[int], point: list[int]) -> int:
    """DensePolynomial.evaluate (src/poly/mod.zig:73-92): index bit j <-> point[j]."""
    cur = list(evals)
    for x in point:  # binding x_0 first == bindLow
        cur = bind_low(cur, x)
    return cur[0] % P


def eq_evals(r: list[int], scale: int = 1) -> list[int]:
    """EqPolynomial.evalsSliceWithScaling (src/poly/mod.zig:252-290), literal loop; big-endian:
    r[0] <-> MSB of the index."""
    n = len(r)
    res = [0] * (1 << n)
    res[0] = scale % P
    size = 1
    for j in range(n - 1, -1, -1):
        rj = r[j]
        for i in range(size):
            x = res[i]
            y = x * rj % P
            res[i + size] = y
            res[i] = (x - y) % P
        size *= 2
    return res


def eq_mle(r: list[int], x: list[int]) -> int:
    """EqPolynomial.mle (src/poly/mod.zig:311-321)."""
    acc = 1
    for ri, xi in zip(r, x):
        acc = acc * (ri * xi + (1 - ri) * (1 - xi)) % P
    return acc


def unipoly_evaluate(coeffs: list[int], x: int) -> int:
    """Horner (src/poly/mod.zig:608-618)."""
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % P
    return acc


def interpolate_degree3(evals: list[int]) -> list[int]:
    """UniPoly.interpolateDegree3 (src/poly/mod.zig:632-677)."""
    p0, p1, p2, p3 = evals
    inv6 = pow(6, -1, P)
    inv2 = pow(2, -1, P)
    c0 = p0 % P
    c1 = (-11 * p0 + 18 * p1 - 9 * p2 + 2 * p3) * inv6 % P
    c2 = (2 * p0 - 5 * p1 + 4 * p2 - p3) * inv2 % P
    c3 = (-p0 + 3 * p1 - 3 * p2 + p3) * inv6 % P
    return [c0, c1, c2, c3]


def evals_to_compressed(evals: list[int]) -> list[int]:
    """UniPoly.evalsToCompressed (src/poly/mod.zig:682-685): [c0, c2, c3]."""
    c = interpolate_degree3(evals)
    return [c[0], c[2], c[3]]


def evaluate_cubic_at_point(evals: list[int], x: int) -> int:
    """evaluateCubicAtPointGeneric (src/zkvm/batched_sumcheck.zig:437-457): Lagrange at 0,1,2,3."""
    e0, e1, e2, e3 = evals
    L0 = (x - 1) * (x - 2) * (x - 3) * pow(-6, -1, P)
    L1 = x * (x - 2) * (x - 3) * pow(2, -1, P)
    L2 = x * (x - 1) * (x - 3) * pow(-2, -1, P)
    L3 = x * (x - 1) * (x - 2) * pow(6, -1, P)
    return (e0 * L0 + e1 * L1 + e2 * L2 + e3 * L3) % P


def compressed_claim_update(claim: int, c0: int, c2: int, c3: int, r: int) -> int:
    """generateBatchedProof's claim update (src/zkvm/batched_sumcheck.zig:386-402): recover c1 from
    s(0)+s(1) = claim, re-expand to s(0..3) and evaluate at r."""
    c1 = (claim - 2 * c0 - c2 - c3) % P
    s = [
        c0 % P,
        (c0 + c1 + c2 + c3) % P,
        (c0 + 2 * c1 + 4 * c2 + 8 * c3) % P,
        (c0 + 3 * c1 + 9 * c2 + 27 * c3) % P,
    ]
    return evaluate_cubic_at_point(s, r)


class GruenSplitEq:
    """GruenSplitEqPolynomial (src/poly/split_eq.zig:63-434), LowToHigh."""

    def __init__(self, tau: list[int], scaling: int = 1):
        self.tau = [t % P for t in tau]
        self.current_index = len(tau)
        self.current_scalar = scaling % P
        n = len(tau)
        self.m = n // 2
        num_x_out = self.m
        num_x_in = (n - 1 - self.m) if n > 1 else 0
        # prefix tables, big-endian: next[2i] = prev[i]*(1-tau_k), next[2i+1] = prev[i]*tau_k (:131-171)
        self.E_out_vec = [[1]]
        for k in range(num_x_out):
            prev = self.E_out_vec[k]
            t = self.tau[k]
            nxt = [0] * (2 * len(prev))
            for i, v in enumerate(prev):
                nxt[2 * i] = v * (1 - t) % P
                nxt[2 * i + 1] = v * t % P
            self.E_out_vec.append(nxt)
        self.E_in_vec = [[1]]
        for k in range(num_x_in):
            prev = self.E_in_vec[k]
            t = self.tau[self.m + k]
            nxt = [0] * (2 * len(prev))
            for i, v in enumerate(prev):
                nxt[2 * i] = v * (1 - t) % P
                nxt[2 * i + 1] = v * t % P
            self.E_in_vec.append(nxt)

    def bind(self, r: int):
        """src/poly/split_eq.zig:213-248."""
        if self.current_index == 0:
            return
        t = self.tau[self.current_index - 1]
        eq_val = (t * r + (1 - t) * (1 - r)) % P
        self.current_scalar = self.current_scalar * eq_val % P
        self.current_index -= 1
        if self.m < self.current_index:
            if len(self.E_in_vec) > 1:
                self.E_in_vec.pop()
        elif self.current_index > 0:
            if len(self.E_out_vec) > 1:
                self.E_out_vec.pop()

    def window_tables(self):
        """getWindowEqTables with window_size 1 (src/poly/split_eq.zig:311-343)."""
        num_unbound = self.current_index
        head_len = max(0, num_unbound - min(1, num_unbound))
        head_out_bits = min(head_len, self.m)
        head_in_bits = max(0, head_len - head_out_bits)
        E_out = self.E_out_vec[head_out_bits] if head_out_bits < len(self.E_out_vec) else self.E_out_vec[-1]
        E_in = self.E_in_vec[head_in_bits] if head_in_bits < len(self.E_in_vec) else self.E_in_vec[-1]
        return E_out, E_in, head_in_bits

    def cubic_round_poly(self, q_constant: int, q_quadratic: int, previous_claim: int) -> list[int]:
        """computeCubicRoundPoly (src/poly/split_eq.zig:353-434): s(0..3) = l(X) * q(X)."""
        if self.current_index == 0:
            return [previous_claim % P, 0, 0, 0]
        t = self.tau[self.current_index - 1]
        eq0 = self.current_scalar * (1 - t) % P
        eq1 = self.current_scalar * t % P
        slope = (eq1 - eq0) % P
        l = [eq0, eq1, (eq0 + 2 * slope) % P, (eq0 + 3 * slope) % P]
        c = q_constant % P
        e = q_quadratic % P
        q1 = 0 if l[1] == 0 else (previous_claim - l[0] * c) * pow(l[1], -1, P) % P
        q2 = (2 * q1 - c + 2 * e) % P
        q3 = (q2 + q1 - c + 4 * e) % P
        return [l[0] * c % P, l[1] * q1 % P, l[2] * q2 % P, l[3] * q3 % P]
