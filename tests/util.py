"""Shared helpers for the parity tests (oracle side)."""
import numpy as np

from oracle import bn254 as B

P = B.FR_MOD
Q = B.FQ_MOD


def fr_arr(vals):
    return B.ints_to_mont_array(vals, P)


def fr_ints(arr):
    return B.mont_array_to_ints(np.asarray(arr).reshape(-1, 4), P)


def rand_fr(seed, n):
    return B.random_fr(np.random.default_rng(seed), n)


def point_from_out(xy, inf):
    pts = B.arrays_to_points(np.asarray(xy).reshape(1, 8), np.array([inf], dtype=np.uint8))
    return pts[0]
