"""ORACLE (test infrastructure only).  Big-integer restatement of the per-round loops of Zolt's
sumcheck provers.  Tables are lists of STANDARD integers mod r.

Each `round_*` function restates one reference computeRoundPolynomial loop; `bind_tables` restates
the matching bind loops.  Reference anchors (relative to /root/reference):
  S1  Sumcheck.Prover.nextRound / receiveChallenge   src/subprotocols/mod.zig:69-122 (half split)
  S2  JoltSpartanInterface round / bind              src/zkvm/r1cs/jolt_r1cs.zig:413-489 (adjacent)
  S3b ValFinalProver                                 src/zkvm/ram/val_final.zig:150-206
  S4  ValEvaluationProver                            src/zkvm/ram/val_evaluation.zig:554-640
  S4b OutputSumcheckProver                           src/zkvm/ram/output_check.zig:375-479
  S5b ProductVirtualRemainderProver (Gruen)          src/zkvm/spartan/product_remainder.zig:269-378
  elementwise eq*(Az*Bz - Cz)                        src/zkvm/r1cs/jolt_r1cs.zig:352-360
"""
from __future__ import annotations

from .bn254 import FR_MOD as P
from .poly import GruenSplitEq, bind_first, bind_low

ADJACENT, HALF = 0, 1


def _pairs(table: list[int], pairing: int):
    n = len(table)
    h = n // 2
    if pairing == ADJACENT:
        return [(table[2 * i], table[2 * i + 1]) for i in range(h)]
    return [(table[i], table[i + h]) for i in range(h)]


def bind_tables(tables: list[list[int]], r: int, pairing: int) -> list[list[int]]:
    f = bind_low if pairing == ADJACENT else bind_first
    return [f(t, r) for t in tables]


def round_sum(tables, pairing):
    """[sum lo, sum hi]: S1 g0/g1 (subprotocols/mod.zig:84-93), S2 p0/p1 (jolt_r1cs.zig:436-439)."""
    lo = hi = 0
    for a, b in _pairs(tables[0], pairing):
        lo += a
        hi += b
    return [lo % P, hi % P]


def round_prod2(tables, pairing):
    """ValFinalProver (val_final.zig:164-184): evals at X=0..3 of sum a(X) b(X), f(X) = X f1 - (X-1) f0."""
    e = [0, 0, 0, 0]
    for (a0, a1), (b0, b1) in zip(_pairs(tables[0], pairing), _pairs(tables[1], pairing)):
        for x in range(4):
            e[x] += (x * a1 - (x - 1) * a0) * (x * b1 - (x - 1) * b0)
    return [v % P for v in e]


def round_prod3(tables, pairing):
    """ValEvaluationProver (val_evaluation.zig:567-601)."""
    e = [0, 0, 0, 0]
    pa, pb, pc = (_pairs(t, pairing) for t in tables[:3])
    for (a0, a1), (b0, b1), (c0, c1) in zip(pa, pb, pc):
        for x in range(4):
            e[x] += (x * a1 - (x - 1) * a0) * (x * b1 - (x - 1) * b0) % P * (x * c1 - (x - 1) * c0)
    return [v % P for v in e]


def round_eq_a_bmc(tables, pairing):
    """OutputSumcheckProver (output_check.zig:385-433): sum eq(X) io(X) (vf(X) - vio(X))."""
    e = [0, 0, 0, 0]
    pe, pa, pb, pc = (_pairs(t, pairing) for t in tables[:4])
    for (e0, e1), (a0, a1), (b0, b1), (c0, c1) in zip(pe, pa, pb, pc):
        v0, v1 = b0 - c0, b1 - c1
        for x in range(4):
            e[x] += (e0 + x * (e1 - e0)) * (a0 + x * (a1 - a0)) % P * (v0 + x * (v1 - v0))
    return [v % P for v in e]


def round_eq_ab_minus_c(tables, pairing):
    """Spartan-style sum eq(X) (az(X) bz(X) - cz(X)) (BASELINE config 3; elementwise form at
    jolt_r1cs.zig:352-360)."""
    e = [0, 0, 0, 0]
    pe, pa, pb, pc = (_pairs(t, pairing) for t in tables[:4])
    for (e0, e1), (a0, a1), (b0, b1), (c0, c1) in zip(pe, pa, pb, pc):
        for x in range(4):
            ex = e0 + x * (e1 - e0)
            ax = a0 + x * (a1 - a0)
            bx = b0 + x * (b1 - b0)
            cx = c0 + x * (c1 - c0)
            e[x] += ex * ((ax * bx - cx) % P)
    return [v % P for v in e]


def round_gruen_ab(tables, split_eq: GruenSplitEq, with_c: bool = False):
    """product_remainder.zig:283-322: nested x_out / x_in loops, inner sums weighted by E_in, outer by
    E_out.  Returns [t0, t_inf]."""
    left, right = tables[0], tables[1]
    cz = tables[2] if with_c else None
    E_out, E_in, _ = split_eq.window_tables()
    bits = (len(E_in).bit_length() - 1) if len(E_in) > 1 else 0
    groups = len(left) // 2
    t0 = tinf = 0
    for xo, eo in enumerate(E_out):
        i0 = ii = 0
        for xi, ei in enumerate(E_in):
            g = (xo << bits) | xi
            if g < groups:
                l0, l1 = left[2 * g], left[2 * g + 1]
                r0, r1 = right[2 * g], right[2 * g + 1]
                p0 = l0 * r0
                if cz is not None:
                    p0 -= cz[2 * g]
                i0 += p0 % P * ei
                ii += (l1 - l0) * (r1 - r0) % P * ei
        t0 += i0 % P * eo
        tinf += ii % P * eo
    return [t0 % P, tinf % P]


ROUND_FNS = {
    "sum": round_sum,
    "prod2": round_prod2,
    "prod3": round_prod3,
    "eq_a_bmc": round_eq_a_bmc,
    "eq_ab_minus_c": round_eq_ab_minus_c,
}


def run_dense(kind: str, tables, pairing: int, challenges, n_extra: int = 0):
    """Alternate round-eval / bind over all variables with the given challenges.
    Returns (list of per-round evals, list of per-round bound tables [after each bind])."""
    tabs = [list(t) for t in tables]
    fn = ROUND_FNS[kind]
    evals, bound = [], []
    for r in challenges:
        evals.append(fn(tabs, pairing))
        tabs = bind_tables(tabs, r, pairing)
        bound.append([list(t) for t in tabs])
    return evals, bound


def run_raf(ra, start_address: int, challenges, initial_claim: int):
    """RafEvaluationProver (src/zkvm/ram/raf_checking.zig:335-409, bind :162-174, claim update :420-445) restated as
    written: the unmap factor is implicit, unmap(x) = base + x * current_power + remaining(i) with
    base = start_address + sum_j bound_j * 8 * 2^j (:345-351); s0 = sum ra_lo * unmap(0), s2 = sum (2 ra_hi - ra_lo) * unmap(2),
    s1 = claim - s0, s3 = s0 - 3 s1 + 3 s2.  Returns (evals per round, bound ra tables per round, final claim)."""
    from . import poly as OP

    ra = [v % P for v in ra]
    log_k = len(ra).bit_length() - 1
    claim = initial_claim % P
    bound, evals_all, tables = [], [], []
    for rnd, r in enumerate(challenges):
        half = len(ra) // 2
        base = start_address % P
        power = 8
        for v in bound:
            base = (base + v * power) % P
            power *= 2
        current_power = power
        s0 = s2 = 0
        remaining_vars = log_k - rnd - 1
        for i in range(half):
            ra_lo, ra_hi = ra[2 * i], ra[2 * i + 1]
            ra_at_2 = (2 * ra_hi - ra_lo) % P
            remaining, rp, idx = 0, current_power * 2, i
            for _ in range(remaining_vars):
                if idx & 1:
                    remaining += rp
                idx >>= 1
                rp *= 2
            s0 = (s0 + ra_lo * (base + remaining)) % P
            s2 = (s2 + ra_at_2 * (base + current_power * 2 + remaining)) % P
        s1 = (claim - s0) % P
        s3 = (s0 - 3 * s1 + 3 * s2) % P
        ev = [s0, s1, s2, s3]
        evals_all.append(ev)
        ra = [((1 - r) * ra[2 * i] + r * ra[2 * i + 1]) % P for i in range(half)]
        tables.append(list(ra))
        bound.append(r)
        claim = OP.evaluate_cubic_at_point(ev, r)
    return evals_all, tables, claim


def run_batched(instances, coeffs, challenges):
    """BatchedSumcheckProver + generateBatchedProof (src/zkvm/batched_sumcheck.zig:158-171, :189-257, :330-420), adjacent
    pairs.  instances: list of (kind, tables).  Returns (compressed round polynomials [c0, c2, c3] per round, final claim,
    input claims).  An instance of n_i rounds starts at round max - n_i; before that it contributes its input claim scaled
    by 2^(rounds still to wait) as a constant polynomial (:206-218).
    Restated AS WRITTEN: the reference scales the waiting constant by 2^(start - round), one doubling more than the
    2^(start - round - 1) that would keep s(0) + s(1) equal to the instance's share of the claim (Jolt's convention); because
    only [c0, c2, c3] travel and c1 is recovered from the running claim (:389), the proof stays self-consistent round by
    round.  With equal numbers of rounds (no waiting instance) the final claim equals sum_i coeff_i * combiner_i(bound
    values), which tests/test_oracle_golden.py checks."""
    from . import poly as OP

    rounds = [len(ts[0]).bit_length() - 1 for _, ts in instances]
    mx = max(rounds)
    per_inst, claims = [], []
    for (kind, ts), n_r in zip(instances, rounds):
        ev, _ = run_dense(kind, ts, 0, challenges[mx - n_r:])
        ev = [list(e) + ([(2 * e[1] - e[0]) % P, (3 * e[1] - 2 * e[0]) % P] if len(e) == 2 else []) for e in ev]
        per_inst.append(ev)
        claims.append((ev[0][0] + ev[0][1]) % P)
    claim = sum(cf * cl * pow(2, mx - n_r, P) for cf, cl, n_r in zip(coeffs, claims, rounds)) % P
    out = []
    for rnd in range(mx):
        comb = [0, 0, 0, 0]
        for ev, cf, cl, n_r in zip(per_inst, coeffs, claims, rounds):
            start = mx - n_r
            for j in range(4):
                term = ev[rnd - start][j] if rnd >= start else cl * pow(2, start - rnd, P)
                comb[j] = (comb[j] + cf * term) % P
        c0, c2, c3 = OP.evals_to_compressed(comb)
        out.append([c0, c2, c3])
        claim = OP.compressed_claim_update(claim, c0, c2, c3, challenges[rnd])
    return out, claim, claims


def run_gruen(tables, tau, challenges, with_c: bool = False):
    tabs = [list(t) for t in tables]
    sq = GruenSplitEq(tau)
    evals, bound = [], []
    for r in challenges:
        evals.append(round_gruen_ab(tabs, sq, with_c))
        tabs = bind_tables(tabs, r, ADJACENT)
        sq.bind(r)
        bound.append([list(t) for t in tabs])
    return evals, bound
