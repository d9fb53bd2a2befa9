"""CPU tests of the product's host side (no GPU): the C-ABI library loads and exports what
include/adsb200.h declares, the host setup reproduces the reference's tables / matrices / factors
(golden vectors from the compiled reference), the chunked-substitution plan is algebraically the
dgbtrs recurrence (emulated in numpy here and compared with the oracle), and the device entry
points refuse to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import kats
import iga_ads_b200 as ads
from iga_ads_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "adsb200.h")).read()
    declared = set(re.findall(r"\b(adsb_[a-z0-9_]+)\s*\(", header))
    declared.discard("adsb_ctx")
    assert len(declared) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_lib.EXPORTS)
    assert _lib.load().adsb_abi_version() == 2


def test_gauss_tables_knots_matrices_factors_vs_golden(golden):
    g = golden["setup"]
    for q in range(2, 8):
        x, w = ads.gauss(q)
        assert np.array_equal(x, g[f"gauss_x_{q}"]) and np.array_equal(w, g[f"gauss_w_{q}"])
    for p, ne in ((1, 5), (2, 12), (3, 7), (4, 9), (5, 6)):
        t = ads.basis_tables(p, ne)
        for k in ("b", "x", "w", "J", "first_dof"):
            assert np.array_equal(t[k], g[f"tab_{p}_{ne}_{k}"]), (p, ne, k)
        assert np.array_equal(ads.knots(p, ne), g[f"tab_{p}_{ne}_knots"])
        for kind, h, fix in ((0, 0.0, 0), (0, 0.0, 1), (1, 0.0, 0), (2, 0.0, 0), (3, 0.005, 0), (3, 3.0, 1)):
            tag = f"mat_{p}_{ne}_{kind}_{fix}_{h}"
            m = ads.matrix_1d(kind, p, ne, h=h, fix=fix)
            assert np.array_equal(m, g[tag]), tag
            if kind in (0, 3):
                lu, piv = ads.band_factorize(m, p, p)
                assert np.array_equal(piv, g[tag + "_ipiv"]), tag
                np.testing.assert_allclose(lu, g[tag + "_lu"], rtol=1e-14, atol=1e-300)


def test_bspline_kats():
    # tests/ads/bspline/bspline_test.cpp:14-68, eval_test.cpp:24-78
    k = ads.knots(2, 4)
    assert np.array_equal(k, [0, 0, 0, 0.25, 0.5, 0.75, 1, 1, 1])
    for x, s in ((-1, 2), (2.0, 5), (0.0, 2), (1.0, 5), (0.25, 3), (0.75, 5), (0.1, 2), (0.9, 5)):
        assert ads.find_span(x, k, 2) == s
    k = ads.knots(2, 5)
    for i in range(101):
        x = i / 100
        d = ads.basis_ders(ads.find_span(x, k, 2), x, k, 2, 2)
        assert abs(d[0].sum() - 1) < 1e-12 and abs(d[1].sum()) < 1e-7 and abs(d[2].sum()) < 1e-7


def test_dimension_mirror_fix_and_factor(golden):
    g = golden["setup"]
    d = ads.dimension(ads.dim_config(3, 7))
    assert d.dofs() == 10
    d.fix_left()
    assert np.array_equal(d.M, g["mat_3_7_0_1_0.0"])
    lu, piv = d.factorize_matrix()
    assert np.array_equal(piv, g["mat_3_7_0_1_0.0_ipiv"])


def test_factorize_reference_kat_pivots_and_singular():
    lu, piv = ads.band_factorize(kats.to_band(kats.MX, 1, 1), 1, 1)
    assert piv[0] == 2  # tests/ads/solver_test.cpp: Mx needs a row interchange
    with pytest.raises(ads.AdsbError) as e:
        ads.band_factorize(kats.to_band(np.zeros((3, 3)), 1, 1), 1, 1)
    assert e.value.code == -3


# ------------------------------------------------------------------ the substitution plan
def get_plan(lu, ipiv, kl, ku):
    lib = _lib.load()
    lu = np.ascontiguousarray(lu)
    ipiv = np.ascontiguousarray(ipiv, dtype=np.int32)
    n, ldab = lu.shape
    dims = np.zeros(16, dtype=np.int32)
    none = None
    _lib.check(lib.adsb_sweep_plan(n, kl, ku, ldab, _lib.d_(lu), _lib.i_(ipiv), _lib.i_(dims), none, none, none,
                                   none, none, none, none, none))
    KL, KD, piv, CH, R, SC, ST, rows, LF, LB, LC, DF, DB, seq, MD, _ = (int(v) for v in dims)
    P = dict(KL=KL, KD=KD, piv=piv, CH=CH, R=R, SC=SC, ST=ST, rows=rows, DF=DF, DB=DB, seq=seq, MD=MD, n=n,
             pv=np.zeros(rows, dtype=np.int32), cfF=np.zeros((rows, LF)), cfB=np.zeros((rows, LB)),
             cfC=np.zeros((rows, LC)), T=np.zeros((SC, KL, KL)), Rm=np.zeros((SC, KD, KD)),
             W=np.zeros((SC, MD - 1, KL, KL)), V=np.zeros((SC, MD - 1, KD, KD)))
    _lib.check(lib.adsb_sweep_plan(n, kl, ku, ldab, _lib.d_(lu), _lib.i_(ipiv), _lib.i_(dims), _lib.i_(P["pv"]),
                                   _lib.d_(P["cfF"]), _lib.d_(P["cfB"]), _lib.d_(P["cfC"]), _lib.d_(P["T"]),
                                   _lib.d_(P["Rm"]), _lib.d_(P["W"]), _lib.d_(P["V"])))
    return P


def emulate_chunked_sweep(P, b, force_seq=False):
    """numpy transcription of sweep_kernel's phases for one line (tests only)."""
    KL, KD, CH, SC, n = P["KL"], P["KD"], P["CH"], P["SC"], P["n"]
    Lm, Ut, rinv = P["cfF"][:, :KL], P["cfB"][:, :KD], P["cfB"][:, KD]
    Psi, Xi = P["cfC"][:, :KD], P["cfC"][:, KD:KD + KL]
    bp = np.zeros(SC * CH + KL)
    bp[:n] = b
    v = np.zeros((SC, CH + KL))
    Dl = np.zeros((SC, KL))
    for c in range(SC):                                 # F1 + B1, local
        j0 = c * CH
        v[c] = bp[j0:j0 + CH + KL]
        o = v[c, CH:].copy()
        for i in range(CH):
            t = P["pv"][j0 + i]
            if t:
                v[c, i], v[c, i + t] = v[c, i + t], v[c, i]
            for r in range(1, KL + 1):
                v[c, i + r] -= Lm[j0 + i, r - 1] * v[c, i]
        Dl[c] = v[c, CH:] - o
        for i in range(CH - 1, -1, -1):
            acc = v[c, i]
            for k in range(KD, 0, -1):
                if i + k < CH:
                    acc -= Ut[j0 + i, k - 1] * v[c, i + k]
            v[c, i] = acc * rinv[j0 + i]
    seq = bool(P["seq"]) or force_seq
    delta = np.zeros((SC, KL))                          # S1
    if seq:
        for c in range(SC - 1):
            delta[c + 1] = Dl[c] + P["T"][c] @ delta[c]
    else:
        for c in range(1, SC):
            delta[c] = Dl[c - 1]
            for d in range(2, P["DF"] + 1):
                if c - d >= 0:
                    delta[c] += P["W"][c, d - 2] @ Dl[c - d]
    X = np.zeros((SC, KD))
    for c in range(SC):
        X[c] = v[c, :KD] + Xi[c * CH:c * CH + KD] @ delta[c]
    tin = np.zeros((SC, KD))                            # S2
    if seq:
        for c in range(SC - 1, 0, -1):
            tin[c - 1] = X[c] + P["Rm"][c] @ tin[c]
    else:
        for c in range(SC - 1):
            tin[c] = X[c + 1]
            for d in range(2, P["DB"] + 1):
                if c + d < SC:
                    tin[c] += P["V"][c, d - 2] @ X[c + d]
    x = np.zeros(SC * CH)
    for c in range(SC):                                 # B3
        j0 = c * CH
        x[j0:j0 + CH] = v[c, :CH] + Psi[j0:j0 + CH] @ tin[c] + Xi[j0:j0 + CH] @ delta[c]
    return x[:n]


PLAN_CASES = [
    # (p, elements, kind, h, fix)
    (1, 40, 0, 0.0, 0), (2, 12, 0, 0.0, 0), (2, 100, 0, 0.0, 1), (3, 70, 0, 0.0, 0), (3, 64, 3, 0.005, 0),
    (4, 66, 0, 0.0, 0), (5, 64, 0, 0.0, 0), (5, 61, 0, 0.0, 1), (4, 70, 3, 3.0, 1), (2, 33, 3, 3.0, 3),
    (5, 28, 3, 0.5, 0), (2, 31, 0, 0.0, 0), (2, 64, 0, 0.0, 0),
]


@pytest.mark.parametrize("p,ne,kind,h,fix", PLAN_CASES)
def test_chunked_plan_is_dgbtrs(oracle, p, ne, kind, h, fix):
    ab = ads.matrix_1d(kind, p, ne, h=h, fix=fix)
    lu, piv = ads.band_factorize(ab, p, p)
    P = get_plan(lu, piv, p, p)
    assert P["KL"] >= p and P["SC"] * P["CH"] >= ne + p and P["SC"] == P["ST"] * P["R"]
    rng = np.random.default_rng(p * 100 + ne)
    b = rng.standard_normal(ne + p)
    want = oracle.solve_factorized(lu, piv, p, p, b)
    for force_seq in (False, True):
        got = emulate_chunked_sweep(P, b, force_seq)
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-13, (P["piv"], P["KL"], P["KD"], P["seq"])


def test_plan_covers_pivoting_and_plain_variants():
    lu, piv = ads.band_factorize(ads.matrix_1d(0, 2, 64), 2, 2)
    P = get_plan(lu, piv, 2, 2)
    assert (P["piv"], P["KL"], P["KD"]) == (0, 2, 2)      # Gram p=2: no interchange, U keeps ku
    lu, piv = ads.band_factorize(ads.matrix_1d(0, 5, 64), 5, 5)
    P = get_plan(lu, piv, 5, 5)
    assert (P["piv"], P["KL"], P["KD"]) == (1, 5, 10)     # Gram p=5 pivots


def test_plan_general_band_kat(oracle):
    # tests/ads/lin/band_solve_test.cpp:16-46 -- kl=1, ku=2, n=6, 4 right-hand sides
    kl, ku, dense, b, x = kats.band_solve_kat()
    lu, piv = ads.band_factorize(kats.to_band(dense, kl, ku), kl, ku)
    P = get_plan(lu, piv, kl, ku)
    for r in range(b.shape[0]):
        got = emulate_chunked_sweep(P, b[r])
        assert np.abs(got - x[r]).max() < 1e-5
        np.testing.assert_allclose(got, np.linalg.solve(dense, b[r]), rtol=1e-12)


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ads.AdsbError) as e:
        ads.Context((14, 14, 14))
    assert e.value.code == -2
    with pytest.raises(ads.AdsbError):
        sim = ads.heat_3d(2, 4, ads.timesteps_config(1, 1e-7))
        sim.prepare_matrices()
    # the one-process slab host (adsb_slabs_*) refuses as loudly, and reports no devices
    lib = _lib.load()
    assert lib.adsb_device_count() == 0
    h = ctypes.c_void_p()
    dev = np.zeros(2, dtype=np.int32)
    n = np.array([14, 14, 14], dtype=np.int32)
    assert lib.adsb_slabs_create(2, _lib.i_(dev), _lib.i_(n), ctypes.byref(h)) == -2
    assert b"no CUDA device" in lib.adsb_last_error()


def test_output_writers_follow_the_reference_formats():
    """include/ads/output/vtk.hpp:45-82 and gnuplot.hpp:45-56 with DEFAULT_FMT = fixed, precision 10, width 18"""
    import io

    from iga_ads_b200.output import linspace, write_gnuplot_2d, write_vtk

    assert np.array_equal(linspace(0.0, 1.0, 4), [0.0, 0.25, 0.5, 0.75, 1.0])
    vals = np.arange(24, dtype=float).reshape((2, 3, 4), order="F") / 7
    s = io.StringIO()
    write_vtk(s, vals)
    lines = s.getvalue().splitlines()
    assert lines[0] == '<?xml version="1.0"?>'
    assert lines[2] == '  <ImageData WholeExtent="0 1 0 2 0 3" origin="0 0 0" spacing="1 1 1">'
    assert lines[3] == '    <Piece Extent="0 1 0 2 0 3">'
    assert lines[5] == '        <DataArray Name="Result"  type="Float32" format="ascii" NumberOfComponents="1">'
    assert lines[6] == "      0.0000000000" and lines[7] == "      0.1428571429"   # memory order, first index fastest
    assert lines[6 + 24:] == ["        </DataArray>", "      </PointData>", "    </Piece>", "  </ImageData>", "</VTKFile>"]
    s = io.StringIO()
    write_gnuplot_2d(s, [0.0, 0.5], [0.0, 1.0, 2.0], np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]))
    rows = s.getvalue().splitlines()
    assert rows[1] == "      0.0000000000      1.0000000000      2.0000000000"
    assert rows[3] == "      0.5000000000      0.0000000000      4.0000000000" and len(rows) == 6


@pytest.mark.parametrize("p,ne,kind,h,fix", [(4, 9, 0, 0.0, 0), (5, 6, 0, 0.0, 0), (5, 43, 0, 0.0, 1), (4, 60, 0, 0.0, 3),
                                             (4, 9, 3, 3.0, 1), (5, 300, 0, 0.0, 1), (5, 40, 3, 0.02, 0), (4, 768, 0, 0.0, 1)])
def test_unpivoted_refactorisation_solves_like_dgbtrs(oracle, p, ne, kind, h, fix):
    """adsb_band_unpivot: the factor the sweeps really use for matrices whose dgbtrf factor has row interchanges
    (every Gram matrix of degree >= 4) solves the same system as the oracle's dgbtrs with the original factor"""
    from iga_ads_b200.host import band_unpivot

    n = ne + p
    ab = ads.matrix_1d(kind, p, ne, h=h, fix=fix)
    lu, piv = ads.band_factorize(ab, p, p)
    assert (piv != np.arange(1, n + 1)).any(), "this case is meant to pivot"
    lu2 = band_unpivot(lu, piv, p, p)
    if kind == 3 and h >= 1.0 and fix:
        # fix_left leaves column 0 of a stiffness-dominated K with entries >> 1 under a unit pivot: multipliers
        # above the bound, the elimination is refused and the caller's interchanges stay
        assert lu2 is None
        return
    assert lu2 is not None
    assert np.abs(lu2[:, 2 * p + 1:]).max() <= 4.0                 # multipliers
    assert np.all(lu2[:, :p] == 0.0)                               # U keeps ku super-diagonals: no fill rows
    rng = np.random.default_rng(3)
    b = rng.standard_normal((7, n))
    want = oracle.solve_factorized(lu, piv, p, p, b).reshape(7, n)
    got = oracle.solve_factorized(lu2, np.arange(1, n + 1, dtype=np.int32), p, p, b).reshape(7, n)
    cond = 1.0
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 2e-12 * cond
    # and the chunk plan of the unpivoted factor is a parallel chain within the kernel's depth
    dims = np.zeros(16, dtype=np.int32)
    from iga_ads_b200 import _lib
    _lib.check(_lib.load().adsb_sweep_plan(n, p, p, 3 * p + 1, _lib.d_(lu2), _lib.i_(np.arange(1, n + 1, dtype=np.int32)),
                                           _lib.i_(dims), None, None, None, None, None, None, None, None))
    assert dims[2] == 0 and dims[1] == p
    if kind == 0:
        assert dims[13] == 0, "Gram factors must not need the sequential chain"
