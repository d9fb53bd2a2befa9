"""Segmented substitution (adsb_segment_plan): host tables checked against the oracle's dgbtrs.

Pass A (every segment solved alone with its own columns of the factor) is done here with the oracle's
dgbtrs on the shifted arrays; the boundary chains and the correction use the tables libadsb200 builds.
The result must equal one dgbtrs over the whole line (include/ads/lin/band_solve.hpp:21-31)."""
import numpy as np
import pytest

from iga_ads_b200 import host
from oracle.oracle import Oracle


def segmented_solve(orc, lu, ipiv, kl, ku, t, B):
    """numpy model of the device path: pass A, states (finite-depth chains), pass B"""
    nrhs, n = B.shape
    S, KL, KD, bounds = t["S"], t["KL"], t["KD"], t["bounds"]
    xh = np.zeros_like(B)
    for s in range(S):
        a, b = int(bounds[s]), int(bounds[s + 1])
        ip = (ipiv[a:b] - a).astype(np.int32)
        xh[:, a:b] = orc.solve_factorized(np.ascontiguousarray(lu[a:b]), ip, kl, ku,
                                          np.ascontiguousarray(B[:, a:b])).reshape(nrhs, b - a)
    D = [xh[:, bounds[s + 1] - KL:bounds[s + 1]] @ t["E"][s].T for s in range(S)]
    din = [sum((D[s - d] @ t["Wf"][s, d - 1].T for d in range(1, t["DF"] + 1) if s - d >= 0),
               np.zeros((nrhs, KL))) for s in range(S)]
    X = [xh[:, bounds[s]:bounds[s] + KD] + din[s] @ t["XiF"][s].T for s in range(S)]
    tin = [sum((X[s + d] @ t["Vb"][s, d - 1].T for d in range(1, t["DB"] + 1) if s + d < S),
               np.zeros((nrhs, KD))) for s in range(S)]
    out = np.zeros_like(B)
    for s in range(S):
        a, b = int(bounds[s]), int(bounds[s + 1])
        cf = t["cf"][a:b]
        out[:, a:b] = xh[:, a:b] + tin[s] @ cf[:, :KD].T + din[s] @ cf[:, KD:].T
    return out


CASES = [  # p, elements, segments, kind, h, fix, align
    (1, 64, 4, 0, 0.0, 0, 1), (2, 64, 4, 0, 0.0, 0, 16), (3, 64, 4, 0, 0.0, 1, 1), (4, 64, 4, 0, 0.0, 0, 1),
    (5, 64, 4, 0, 0.0, 0, 1), (2, 512, 8, 0, 0.0, 0, 16), (3, 256, 8, 0, 0.0, 1, 16), (5, 96, 3, 0, 0.0, 1, 1),
    (3, 256, 4, 3, 1e-2 / 3, 0, 1), (2, 128, 4, 3, 0.5e-2, 0, 1), (3, 509, 7, 0, 0.0, 0, 18),
]


@pytest.mark.parametrize("p,ne,S,kind,h,fix,align", CASES)
def test_segmented_solve_equals_dgbtrs(p, ne, S, kind, h, fix, align):
    orc = Oracle()
    ab = orc.matrix_1d(kind, p, ne, h=h, fix=fix)
    lu, ipiv, info = orc.factorize(ab, p, p)
    assert info == 0
    n = lu.shape[0]
    bounds = host.segment_bounds(ipiv, p, S, align)
    assert bounds[0] == 0 and bounds[-1] == n and np.all(np.diff(bounds) > 0)
    t = host.segment_plan(lu, ipiv, p, p, bounds)
    rng = np.random.default_rng(p * 1000 + ne)
    B = rng.standard_normal((7, n))
    want = orc.solve_factorized(lu, ipiv, p, p, B.copy()).reshape(7, n)
    got = segmented_solve(orc, lu, ipiv, p, p, t, B)
    err = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert err < 5e-15, err


def test_depth_one_for_gram_slabs():
    """z-slabs of 64 planes of the headline problem: only the adjacent slab matters (chain depth 1)"""
    orc = Oracle()
    lu, ipiv, _ = orc.factorize(orc.matrix_1d(0, 2, 512), 2, 2)
    t = host.segment_plan(lu, ipiv, 2, 2, host.segment_bounds(ipiv, 2, 8, 16))
    assert (t["DF"], t["DB"]) == (1, 1)
    assert list(np.diff(t["bounds"])) == [64] * 7 + [66]


def test_rejects_bad_cuts():
    orc = Oracle()
    lu, ipiv, _ = orc.factorize(orc.matrix_1d(0, 5, 64), 5, 5)
    pv = ipiv - 1 - np.arange(len(ipiv))
    crossing = next(a for a in range(6, 60) if any(j + pv[j] >= a for j in range(a - 5, a)))
    with pytest.raises(Exception):
        host.segment_plan(lu, ipiv, 5, 5, np.array([0, crossing, len(ipiv)], dtype=np.int32))
    with pytest.raises(Exception):  # a segment shorter than the band
        host.segment_plan(lu, ipiv, 5, 5, np.array([0, 3, len(ipiv)], dtype=np.int32))
