"""ORACLE (test infrastructure only -- never imported by the product path).

Big-integer restatement of Zolt's BN254 field / curve / MSM semantics.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (zolt_b200/) talks to
the CUDA C-ABI library and fails loudly when that library is missing.

Why big-int math is a sound oracle: every reference field op returns the
canonical representative (< p) in Montgomery form (final conditional
subtract at reference src/field/mod.zig:301-305, :771-776) and the MSM returns
affine coordinates (src/msm/mod.zig:437), so any mathematically exact
computation yields identical limbs.

Reference anchors (all relative to /root/reference):
  constants           src/field/mod.zig:16-41 (Fr), :51-75 (Fq)
  Montgomery encode   src/field/mod.zig:617-652
  bytes               src/field/mod.zig:625-639, :685-695
  affine / Jacobian   src/msm/mod.zig:15-140, :145-329
  MSM.compute         src/msm/mod.zig:355-372, pippenger :375-438
  getWindow           src/msm/mod.zig:441-471, optimalWindowSize :475-484
  HyperKZG setup      src/poly/commitment/mod.zig:174-213, commit :239-255
"""
from __future__ import annotations

import numpy as np

# --- constants (src/field/mod.zig:16-21, :51-56) ---------------------------------
FR_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
FQ_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R_BITS = 256
FR_R = (1 << R_BITS) % FR_MOD
FQ_R = (1 << R_BITS) % FQ_MOD
FR_R2 = (FR_R * FR_R) % FR_MOD
FQ_R2 = (FQ_R * FQ_R) % FQ_MOD
FR_RINV = pow(FR_R, -1, FR_MOD)
FQ_RINV = pow(FQ_R, -1, FQ_MOD)
FR_INV64 = (-pow(FR_MOD, -1, 1 << 64)) % (1 << 64)  # src/field/mod.zig:41
FQ_INV64 = (-pow(FQ_MOD, -1, 1 << 64)) % (1 << 64)  # src/field/mod.zig:75
CURVE_B = 3  # y^2 = x^3 + 3, src/msm/mod.zig:106-115
G1_GEN = (1, 2)  # src/msm/mod.zig:43-49
MASK64 = (1 << 64) - 1


def int_to_limbs(v: int) -> list[int]:
    """256-bit integer -> 4 little-endian u64 limbs (layout of `limbs: [4]u64`)."""
    return [(v >> (64 * i)) & MASK64 for i in range(4)]


def limbs_to_int(limbs) -> int:
    return sum(int(l) << (64 * i) for i, l in enumerate(limbs))


# --- Montgomery encode / decode (src/field/mod.zig:617-652) ----------------------
def fr_to_mont(v: int) -> int:
    return (v * FR_R) % FR_MOD


def fr_from_mont(m: int) -> int:
    return (m * FR_RINV) % FR_MOD


def fq_to_mont(v: int) -> int:
    return (v * FQ_R) % FQ_MOD


def fq_from_mont(m: int) -> int:
    return (m * FQ_RINV) % FQ_MOD


def mont_mul(a: int, b: int, mod: int) -> int:
    """montgomeryMul on raw limbs-as-integers: a*b*R^-1 mod p.  Valid for any
    a < 2^256, b < p (CIOS bound); result canonical."""
    rinv = FR_RINV if mod == FR_MOD else FQ_RINV
    return (a * b * rinv) % mod


def fr_from_bytes(b: bytes) -> int:
    """BN254Scalar.fromBytes (src/field/mod.zig:625-639): LE 256-bit value times
    R^2 through montgomeryMul == (value mod r) in Montgomery form.  Returns the
    *standard* integer value (value mod r)."""
    buf = bytes(b[:32]).ljust(32, b"\0")
    return int.from_bytes(buf, "little") % FR_MOD


def fr_to_bytes(v: int) -> bytes:
    """BN254Scalar.toBytes (src/field/mod.zig:685-695): canonical LE."""
    return int(v % FR_MOD).to_bytes(32, "little")


# --- numpy interchange: arrays of shape (n, 4) uint64, Montgomery limbs ------------
def ints_to_mont_array(vals, mod: int) -> np.ndarray:
    """Standard integers -> (n,4) uint64 Montgomery limbs."""
    r = FR_R if mod == FR_MOD else FQ_R
    buf = b"".join(((int(v) * r) % mod).to_bytes(32, "little") for v in vals)
    return np.frombuffer(buf, dtype="<u8").reshape(-1, 4).copy()


def mont_array_to_ints(arr: np.ndarray, mod: int) -> list[int]:
    """(n,4) uint64 Montgomery limbs -> standard integers."""
    rinv = FR_RINV if mod == FR_MOD else FQ_RINV
    raw = np.ascontiguousarray(arr, dtype="<u8").tobytes()
    return [
        (int.from_bytes(raw[i : i + 32], "little") * rinv) % mod
        for i in range(0, len(raw), 32)
    ]


def raw_array_to_ints(arr: np.ndarray) -> list[int]:
    raw = np.ascontiguousarray(arr, dtype="<u8").tobytes()
    return [int.from_bytes(raw[i : i + 32], "little") for i in range(0, len(raw), 32)]


def ints_to_raw_array(vals) -> np.ndarray:
    buf = b"".join(int(v).to_bytes(32, "little") for v in vals)
    return np.frombuffer(buf, dtype="<u8").reshape(-1, 4).copy()


def random_fr(rng: np.random.Generator, n: int) -> list[int]:
    """Uniform in [0, r): 256-bit draw, reject >= r (SURVEY 8d workload recipe)."""
    out: list[int] = []
    while len(out) < n:
        need = n - len(out)
        raw = rng.integers(0, 1 << 64, size=(need + 8, 4), dtype=np.uint64)
        raw[:, 3] &= np.uint64((1 << 62) - 1)  # 254-bit draw keeps rejection rate low
        for v in raw_array_to_ints(raw):
            if v < FR_MOD:
                out.append(v)
                if len(out) == n:
                    break
    return out


# --- affine group law over Fq on standard integers --------------------------------
# Points are (x, y) tuples or None for the identity (AffinePoint.infinity).
def g1_is_on_curve(P) -> bool:
    if P is None:
        return True
    x, y = P
    return (y * y - x * x * x - CURVE_B) % FQ_MOD == 0


def g1_neg(P):
    if P is None:
        return None
    return (P[0], (-P[1]) % FQ_MOD)


def g1_add(P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % FQ_MOD == 0:
            return None
        lam = (3 * x1 * x1) * pow(2 * y1, -1, FQ_MOD) % FQ_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, FQ_MOD) % FQ_MOD
    x3 = (lam * lam - x1 - x2) % FQ_MOD
    y3 = (lam * (x1 - x3) - y1) % FQ_MOD
    return (x3, y3)


def g1_mul(P, k: int):
    """k*P by double-and-add (result-equivalent to scalarMul, src/msm/mod.zig:503-540)."""
    k %= FR_MOD
    acc = None
    add = P
    while k:
        if k & 1:
            acc = g1_add(acc, add)
        add = g1_add(add, add)
        k >>= 1
    return acc


# --- Jacobian arithmetic for a fast oracle MSM (no inversion per add) ----------------
def _jac_double(P):
    X, Y, Z = P
    if Z == 0:
        return P
    q = FQ_MOD
    A = X * X % q
    B = Y * Y % q
    C = B * B % q
    D = 2 * ((X + B) * (X + B) - A - C) % q
    E = 3 * A % q
    X3 = (E * E - 2 * D) % q
    Y3 = (E * (D - X3) - 8 * C) % q
    Z3 = 2 * Y * Z % q
    return (X3, Y3, Z3)


def _jac_add(P, Q):
    q = FQ_MOD
    X1, Y1, Z1 = P
    X2, Y2, Z2 = Q
    if Z1 == 0:
        return Q
    if Z2 == 0:
        return P
    Z1Z1 = Z1 * Z1 % q
    Z2Z2 = Z2 * Z2 % q
    U1 = X1 * Z2Z2 % q
    U2 = X2 * Z1Z1 % q
    S1 = Y1 * Z2 * Z2Z2 % q
    S2 = Y2 * Z1 * Z1Z1 % q
    H = (U2 - U1) % q
    rr = (S2 - S1) % q
    if H == 0:
        if rr == 0:
            return _jac_double(P)
        return (1, 1, 0)
    HH = H * H % q
    HHH = H * HH % q
    V = U1 * HH % q
    X3 = (rr * rr - HHH - 2 * V) % q
    Y3 = (rr * (V - X3) - S1 * HHH) % q
    Z3 = Z1 * Z2 * H % q
    return (X3, Y3, Z3)


def _jac_from_affine(P):
    return (1, 1, 0) if P is None else (P[0], P[1], 1)


def _jac_to_affine(P):
    X, Y, Z = P
    if Z == 0:
        return None
    zi = pow(Z, -1, FQ_MOD)
    zi2 = zi * zi % FQ_MOD
    return (X * zi2 % FQ_MOD, Y * zi2 * zi % FQ_MOD)


# --- MSM ---------------------------------------------------------------------------
def optimal_window_size(n: int) -> int:
    """src/msm/mod.zig:475-484 (pinned by the reference test at :827-837)."""
    for bound, c in ((8, 1), (32, 2), (128, 3), (512, 4), (2048, 5), (8192, 6), (32768, 7)):
        if n < bound:
            return c
    return 8


def get_window(scalar_std: int, window_idx: int, c: int) -> int:
    """getWindow (src/msm/mod.zig:441-471) on the canonical scalar: bits
    [w*c, w*c+c), zero beyond bit 255."""
    off = window_idx * c
    if off >= 256:
        return 0
    return (scalar_std >> off) & ((1 << c) - 1)  # scalar < 2^256: high bits read as 0


def msm_pippenger_literal(bases, scalars):
    """Literal restatement of pippengerMSM (src/msm/mod.zig:375-438): unsigned
    c-bit digits, windows MSB->LSB, c doublings, bucket pass, running-sum
    reduction.  bases: affine tuples / None; scalars: standard integers."""
    n = len(bases)
    c = optimal_window_size(n)
    num_windows = (256 + c - 1) // c
    num_buckets = (1 << c) - 1
    ident = (1, 1, 0)
    result = ident
    for w in range(num_windows - 1, -1, -1):
        if result[2] != 0:
            for _ in range(c):
                result = _jac_double(result)
        buckets = [ident] * num_buckets
        for B, s in zip(bases, scalars):
            if B is None:
                continue
            d = get_window(s, w, c)
            if d == 0:
                continue
            buckets[d - 1] = _jac_add(buckets[d - 1], _jac_from_affine(B))
        running = ident
        wsum = ident
        for b in range(num_buckets - 1, -1, -1):
            running = _jac_add(running, buckets[b])
            wsum = _jac_add(wsum, running)
        result = _jac_add(result, wsum)
    return _jac_to_affine(result)


def msm(bases, scalars):
    """MSM(F,Fp).compute result semantics (src/msm/mod.zig:355-372): identity on
    empty input; otherwise the exact group sum, returned affine (None = identity,
    which the reference encodes as x=0,y=0,infinity=true, :24-30)."""
    assert len(bases) == len(scalars)
    if len(bases) == 0:
        return None
    if len(bases) < 8:
        acc = (1, 1, 0)
        for B, s in zip(bases, scalars):
            acc = _jac_add(acc, _jac_from_affine(g1_mul(B, s)))
        return _jac_to_affine(acc)
    return msm_pippenger_literal(bases, scalars)


def msm_closed_form(base_logs, scalars):
    """If bases[i] = k_i * G, then MSM = (sum s_i k_i mod r) * G."""
    acc = 0
    for k, s in zip(base_logs, scalars):
        acc = (acc + k * s) % FR_MOD
    return g1_mul(G1_GEN, acc)


# --- HyperKZG mock SRS + commit -------------------------------------------------------
HYPERKZG_TAU = 0x12345678  # src/poly/commitment/mod.zig:189


def hyperkzg_setup(max_degree: int):
    """powers_of_tau_g1[i] = tau^i * G (src/poly/commitment/mod.zig:193-199)."""
    out = []
    t = 1
    for _ in range(max_degree):
        out.append(g1_mul(G1_GEN, t))
        t = t * HYPERKZG_TAU % FR_MOD
    return out


def hyperkzg_commit(powers, evals):
    """HyperKZG.commit (src/poly/commitment/mod.zig:239-255)."""
    n = min(len(evals), len(powers))
    if n == 0:
        return None
    return msm(powers[:n], evals[:n])


def hyperkzg_commit_closed_form(evals):
    acc = 0
    t = 1
    for e in evals:
        acc = (acc + e * t) % FR_MOD
        t = t * HYPERKZG_TAU % FR_MOD
    return g1_mul(G1_GEN, acc)


def hyperkzg_open(powers, evals, point, value):
    """HyperKZG.open (src/poly/commitment/mod.zig:261-324): per variable, commit the quotient hi - lo
    (:296-303), then fold new[j] = (1 - r) * low[j] + r * high[j] (:305-312), half-split order; stops early when
    the table runs out (:289).  Returns (quotient commitments, final_eval); num_vars == 0 keeps `value` (:270-276).
    `powers` may be None: commitments are then taken from the closed form (sum_j q_j tau^j) G of the mock SRS."""
    if not point:
        return [], value % FR_MOD
    cur = [e % FR_MOD for e in evals]
    quotients = []
    for r in point:
        half = len(cur) // 2
        if half == 0:
            break
        q = [(cur[j + half] - cur[j]) % FR_MOD for j in range(half)]
        quotients.append(hyperkzg_commit(powers, q) if powers is not None else hyperkzg_commit_closed_form(q))
        one_minus_r = (1 - r) % FR_MOD
        cur = [(cur[j] * one_minus_r + cur[j + half] * r) % FR_MOD for j in range(half)]
    return quotients, (cur[0] if cur else 0)


def evaluate_multilinear_as_written(evals, point):
    """HyperKZG.evaluateMultilinear (src/poly/commitment/mod.zig:788-817) AS WRITTEN: index bit j <-> point[j]; the direct
    sum only when there are at most 10 variables and 1024 entries, otherwise the "fallback" returns evals[0]."""
    if not evals:
        return 0
    if not point:
        return evals[0] % FR_MOD
    if len(point) <= 10 and len(evals) <= 1024:
        total = 0
        for idx, e in enumerate(evals):
            term, bits = e, idx
            for r in point:
                term = term * (r if bits & 1 else (1 - r)) % FR_MOD
                bits >>= 1
            total = (total + term) % FR_MOD
        return total
    return evals[0] % FR_MOD


def hyperkzg_batch_open(powers, polys, point):
    """HyperKZG.batchOpen (src/poly/commitment/mod.zig:607-732).  Returns (quotient commitments, evaluations, final_eval,
    batching challenge).  `powers` may be None (closed-form commitments of the mock SRS)."""
    if not polys:
        return [], [], 0, 1
    evaluations = [evaluate_multilinear_as_written(p, point) for p in polys]
    gamma = 0x9A8B7C6D
    for r in point:
        gamma = gamma * ((r + 11) % FR_MOD) % FR_MOD
    if gamma == 0:
        gamma = 1
    size = len(polys[0])
    combined = [0] * size
    gp = 1
    for p in polys:
        for j in range(min(size, len(p))):
            combined[j] = (combined[j] + gp * p[j]) % FR_MOD
        gp = gp * gamma % FR_MOD
    if not point:
        acc, gp = 0, 1
        for e in evaluations:
            acc = (acc + gp * e) % FR_MOD
            gp = gp * gamma % FR_MOD
        return [], evaluations, acc, gamma
    quotients, final = hyperkzg_open(powers, combined, point, 0)
    return quotients, evaluations, final, gamma


# --- affine point <-> C-ABI layout --------------------------------------------------
def points_to_arrays(points):
    """Affine tuples/None -> (xy (n,8) uint64 Montgomery Fq limbs, inf (n,) uint8).
    Identity is x=0,y=0,infinity=1 (src/msm/mod.zig:24-30)."""
    n = len(points)
    xs = [0 if P is None else P[0] for P in points]
    ys = [0 if P is None else P[1] for P in points]
    xy = np.empty((n, 8), dtype=np.uint64)
    if n:
        xy[:, :4] = ints_to_mont_array(xs, FQ_MOD)
        xy[:, 4:] = ints_to_mont_array(ys, FQ_MOD)
    inf = np.array([1 if P is None else 0 for P in points], dtype=np.uint8)
    return xy, inf


def arrays_to_points(xy: np.ndarray, inf: np.ndarray):
    xy = np.asarray(xy, dtype=np.uint64).reshape(-1, 8)
    xs = mont_array_to_ints(xy[:, :4], FQ_MOD)
    ys = mont_array_to_ints(xy[:, 4:], FQ_MOD)
    return [None if f else (x, y) for x, y, f in zip(xs, ys, np.asarray(inf).reshape(-1))]
