"""One process, several devices (zolt_gpu_init_multi -- the shape `zolt prove` has, src/zkvm/mod.zig:366-462): the packaged
point-range MSM and the sharded sumcheck with one host thread per device, against the oracle.  On a one-GPU box the same
device is listed twice / four times (ranks share it); with several GPUs visible the real device set is used as well."""
import numpy as np
import pytest

import zolt_b200 as Z
from oracle import bn254 as B
from oracle import cpu_ref as R
from tests.util import fr_arr, rand_fr
from zolt_b200 import host as H

pytestmark = pytest.mark.gpu


def _device_sets():
    import torch

    n = torch.cuda.device_count()
    sets = [[0, 0], [0, 0, 0, 0]]
    for w in (2, 4, 8):
        if n >= w:
            sets.append(list(range(w)))
    return sets


@pytest.mark.parametrize("devices", _device_sets() if True else [], ids=lambda d: "dev" + "".join(map(str, d)))
def test_multi_device_msm_and_sumcheck(built_lib, devices):
    world = len(devices)
    m = Z.MultiContext(devices)
    try:
        # MSM: rank r holds bases / takes scalars of the contiguous range r
        n = 1 << 12
        logs, scalars = rand_fr(901, n), rand_fr(902, n)
        per = n // world
        bases = []
        for r in range(world):
            ctx, _ = m.rank(r)
            bases.append(Z.Bases.fixed_base_mul(ctx, fr_arr(logs[r * per:(r + 1) * per])))
        xy, inf = m.msm(bases, fr_arr(scalars))
        assert B.arrays_to_points(xy.reshape(1, 8), [inf])[0] == B.msm_closed_form(logs, scalars)
        xy2, inf2 = m.msm(bases, fr_arr(scalars))  # again: workspaces are warm, same answer
        assert (xy2 == xy).all() and inf2 == inf
        for b in bases:
            b.free()
        # sumcheck: blocks of the hypercube per device, mailbox exchange inside the round kernels, resident tail, final gather
        for kind, comb, k, log_n in (("eq_ab_minus_c", Z.SC_EQ_AB_MINUS_C, 4, 14), ("prod2", Z.SC_PROD2, 2, 9)):
            tabs = [fr_arr(rand_fr(910 + 10 * log_n + t, 1 << log_n)) for t in range(k)]
            ev, ch, fin = H.sumcheck_prove_multi(m, comb, tabs)
            r_ev, r_ch, r_fin = R.sc_prove(kind, 0, tabs)
            ne = 3 if comb == Z.SC_PROD2 else 4
            assert (ch == r_ch).all() and (ev[:, :ne] == r_ev[:, :ne]).all(), kind
            assert all((fin[t] == r_fin[t][0]).all() for t in range(k)), kind
    finally:
        m.close()
