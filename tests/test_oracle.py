"""CPU tests that PIN the oracle (oracle/ads_oracle.c): reference KATs, golden vectors made by the
compiled unmodified reference (tests/golden/make_golden.py), and -- when oracle/_ref loads -- the
reference itself, live."""
import numpy as np
import pytest

import kats
from oracle.oracle import NDIM, rel_l2, synthetic_state


def _factor_all(oracle, mats, kl=1, ku=1):
    out = [oracle.factorize(kats.to_band(m, kl, ku), kl, ku) for m in mats]
    return [o[0] for o in out], [o[1] for o in out]


@pytest.mark.parametrize("kat", [kats.ADS_1D, kats.ADS_2D, kats.ADS_3D], ids=["1d", "2d", "3d"])
def test_ads_solve_reference_kats(oracle, kat):
    # tests/ads/solver_test.cpp:63-255; Catch Approx => rel 1.2e-5, we demand far tighter
    f, piv = _factor_all(oracle, kat["mats"])
    nd = len(kat["shape"])
    x = oracle.ads_solve(kat["shape"], f, piv, [1] * nd, [1] * nd, np.array(kat["rhs"], float))
    np.testing.assert_allclose(x, kat["expected"], rtol=1e-13, atol=1e-13)


def test_pivoting_kat_really_pivots(oracle):
    _, piv, info = oracle.factorize(kats.to_band(kats.MX, 1, 1), 1, 1)
    assert info == 0 and piv[0] == 2  # row interchange in the first column


def test_band_solve_kat(oracle):
    kl, ku, dense, b, x = kats.band_solve_kat()
    f, piv, info = oracle.factorize(kats.to_band(dense, kl, ku), kl, ku)
    assert info == 0
    got = oracle.solve_factorized(f, piv, kl, ku, b).reshape(b.shape)
    assert np.abs(got - x).max() < 1e-5           # the reference's own tolerance
    np.testing.assert_allclose(got, np.linalg.solve(dense, b.T).T, rtol=1e-12)


def test_tensor_layout_and_rotation_kat(oracle):
    # tests/ads/lin/tensor_test.cpp:13-18: t(2,1) of a 5x3 tensor lives at index 5+2
    assert 2 + 5 * 1 == 7
    shape, a, e = kats.rotation_kat()
    out = oracle.cyclic_transpose(shape, a)
    assert np.array_equal(out, e)
    s1 = shape[1:] + shape[:1]
    s2 = s1[1:] + s1[:1]
    back = oracle.cyclic_transpose(s2, oracle.cyclic_transpose(s1, out))
    assert np.array_equal(back, a)


def test_bspline_kats(oracle):
    # tests/ads/bspline/bspline_test.cpp:14-68
    k = oracle.knots(2, 4)
    assert np.array_equal(k, [0, 0, 0, 0.25, 0.5, 0.75, 1, 1, 1])
    for x, s in ((-1, 2), (2.0, 5), (0.0, 2), (1.0, 5), (0.25, 3), (0.75, 5), (0.1, 2), (0.3, 3),
                 (0.7, 4), (0.9, 5)):
        assert oracle.find_span(x, k, 2) == s
    rep = [0, 0, 0, 0, 1, 1, 2, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5]
    for x, s in ((-4, 3), (8, 13), (0, 3), (5, 13), (1, 5), (2, 6), (3, 9), (4, 13), (0.3, 3),
                 (1.2, 5), (2.8, 6), (3.1, 9), (4.9, 13)):
        assert oracle.find_span(x, rep, 3) == s
    assert np.array_equal(oracle.basis_tables(2, 4)["first_dof"], [0, 1, 2, 3])


def test_partition_of_unity(oracle):
    # tests/ads/bspline/eval_test.cpp:24-78
    p, ne = 2, 5
    k = oracle.knots(p, ne)
    for i in range(101):
        x = (1 - i / 100) * 0.0 + (i / 100) * 1.0
        span = oracle.find_span(x, k, p)
        d = oracle.basis_ders(span, x, k, p, 2)
        assert abs(d[0].sum() - 1.0) < 1e-12
        assert abs(d[1].sum()) < 1e-7 and abs(d[2].sum()) < 1e-7


def test_gauss_against_reference_table(oracle, golden):
    g = golden["setup"]
    for q in range(2, 9):
        x, w = oracle.gauss(q)
        if q <= 7:  # every rule the configs use (q = p+1 <= 6) must be bit-identical
            assert np.array_equal(x, g[f"gauss_x_{q}"]) and np.array_equal(w, g[f"gauss_w_{q}"])
        else:
            np.testing.assert_allclose(x, g[f"gauss_x_{q}"], rtol=0, atol=2.3e-16)
            np.testing.assert_allclose(w, g[f"gauss_w_{q}"], rtol=0, atol=2.3e-16)
        assert abs(w.sum() - 2.0) < 1e-14


@pytest.mark.parametrize("p,ne", [(1, 5), (2, 12), (3, 7), (4, 9), (5, 6)])
def test_setup_tables_bitwise_vs_golden(oracle, golden, p, ne):
    g = golden["setup"]
    t = oracle.basis_tables(p, ne)
    for k in ("b", "x", "w", "J", "first_dof"):
        assert np.array_equal(t[k], g[f"tab_{p}_{ne}_{k}"]), k
    assert np.array_equal(oracle.knots(p, ne), g[f"tab_{p}_{ne}_knots"])
    for kind, h, fix in ((0, 0.0, 0), (0, 0.0, 1), (1, 0.0, 0), (2, 0.0, 0), (3, 0.005, 0), (3, 3.0, 1)):
        tag = f"mat_{p}_{ne}_{kind}_{fix}_{h}"
        m = oracle.matrix_1d(kind, p, ne, h=h, fix=fix)
        assert np.array_equal(m, g[tag]), tag
        if kind in (0, 3):
            f, piv, info = oracle.factorize(m, p, p)
            assert info == 0
            assert np.array_equal(piv, g[tag + "_ipiv"]), tag
            np.testing.assert_allclose(f, g[tag + "_lu"], rtol=1e-14, atol=1e-300)


def test_golden_has_pivoting_cases(golden):
    g = golden["setup"]
    piv = g["mat_5_6_0_0_0.0_ipiv"]
    assert (piv != np.arange(1, len(piv) + 1)).any()


@pytest.mark.parametrize("tag", ["p2", "p3fix", "p5", "p4K"])
def test_ads_solve_vs_golden(oracle, golden, tag):
    g = golden["solve"]
    p, ne, kind, fix = (int(v) for v in g[f"{tag}_meta"])
    h = float(g[f"{tag}_h"][0])
    f, piv, _ = oracle.factorize(oracle.matrix_1d(kind, p, ne, h=h, fix=fix), p, p)
    n = ne + p
    for nd in (1, 2, 3):
        x = oracle.ads_solve((n,) * nd, [f] * nd, [piv] * nd, [p] * nd, [p] * nd, g[f"{tag}_{nd}d_rhs"])
        assert rel_l2(x, g[f"{tag}_{nd}d_x"]) < 1e-14


def test_ads_solve_mixed_shapes_vs_golden(oracle, golden):
    g = golden["solve"]
    mats, pivs, ps = [], [], []
    for p, ne in ((2, 12), (3, 7), (5, 6)):
        f, piv, _ = oracle.factorize(oracle.matrix_1d(0, p, ne), p, p)
        mats.append(f), pivs.append(piv), ps.append(p)
    x = oracle.ads_solve((14, 10, 11), mats, pivs, ps, ps, g["mixed_rhs"])
    assert rel_l2(x, g["mixed_x"]) < 1e-14
    assert np.array_equal(oracle.cyclic_transpose((2, 3, 4), g["rot_in"]), g["rot_out"])


def _problem_tags(golden_file):
    return sorted(k[:-5] for k in golden_file.files if k.endswith("_meta"))


def test_problems_vs_golden(oracle, golden):
    """Every example on the path: shipped initial state, full trajectories from the shipped and
    the synthetic state, single step, and each compute_rhs alone -- against reference outputs."""
    g = golden["problems"]
    tags = _problem_tags(g)
    assert len(tags) >= 13
    for tag in tags:
        pid, p, ne, ns = (int(v) for v in g[tag + "_meta"])
        dt = float(g[tag + "_dt"][0])
        u, _ = oracle.run(pid, p, ne, dt, 0)
        assert rel_l2(u, g[tag + "_shipped_init"]) < 1e-14, tag
        u, _ = oracle.run(pid, p, ne, dt, ns)
        assert rel_l2(u, g[tag + "_shipped"]) < 1e-13, tag
        u0 = g[tag + "_u0"]
        assert np.array_equal(u0, synthetic_state((ne + p,) * NDIM[pid]))
        u, _ = oracle.run(pid, p, ne, dt, int(g[tag + "_syn_steps"][0]), u0=u0)
        assert rel_l2(u, g[tag + "_syn"]) < 1e-13, tag
        u, _ = oracle.run(pid, p, ne, dt, 1, u0=u0)
        assert rel_l2(u, g[tag + "_syn_step1"]) < 1e-14, tag
        s = 1
        while f"{tag}_rhs{s}" in g.files:
            rhs, _ = oracle.run(pid, p, ne, dt, 1, u0=u0, stage=s)
            assert rel_l2(rhs, g[f"{tag}_rhs{s}"]) < 1e-14, (tag, s)
            s += 1


def test_heat3d_checksum_from_survey(golden):
    # heat_3d p=2, 12^3, dt=1e-7, 100 steps
    u = golden["problems"]["heat_3d_p2_n12_shipped"]
    assert abs(u.sum() - 132.96044839648852) < 1e-9
    assert abs(np.linalg.norm(u) - 10.367682819844902) < 1e-11


def test_live_reference_agrees(oracle, ref):
    """Only where oracle/_ref/libads_ref.so loads (this container, or the GPU box via the
    travelling prebuilt .so): different sizes than the golden files."""
    for name, p, ne, dt in (("heat_3d", 2, 5, 1e-7), ("scalability_3d", 5, 6, 1e-6),
                            ("heat_2d", 3, 20, 1e-5), ("implicit_2d", 4, 12, 1e-2),
                            ("implicit_3d", 3, 4, 1e-2)):
        nd = 3 if name.endswith("3d") else 2
        u0 = synthetic_state((ne + p,) * nd, seed=5)
        a, _ = oracle.run(name, p, ne, dt, 2, u0=u0)
        b, _ = ref.run(name, p, ne, dt, 2, u0=u0)
        assert rel_l2(a, b) < 1e-14, name


def test_kronecker_heat_rhs_equals_the_element_loop(oracle):
    """pins oracle.kronecker_heat_rhs (the independent full-size check of K1, tests/test_gpu.py) to the oracle's
    restatement of heat_3d.hpp:49-67"""
    from oracle.oracle import kronecker_heat_rhs, rel_l2, synthetic_state

    for p, ne in ((2, 12), (3, 7)):
        n = ne + p
        u0 = synthetic_state((n,) * 3)
        want, _ = oracle.run("heat_3d", p, ne, 1e-7, 1, u0=u0, stage=1)
        assert rel_l2(kronecker_heat_rhs(oracle, p, ne, 1e-7, u0), want) < 5e-15
