"""General pointwise forms (csrc/quadbrick.cuh: brick quadrature kernel, adsb_compute_rhs_pointwise and method
ADSB_RHS_QUADRATURE): the nonlinear flow form against the compiled reference's examples/flow (tests/golden/flow.npz),
linear forms with advection and sources against the numpy restatement of the reference's element loop
(oracle.pointwise_rhs, pinned to the reference in the CPU test below), determinism, and a size with many bricks,
colours and z segments."""
import numpy as np
import pytest

import iga_ads_b200 as ads
from iga_ads_b200 import PointForm, U, U_PREV
from oracle.oracle import Oracle, flow_form, pointwise_rhs, rel_l2, synthetic_state

FLOW_CASES = [(2, 9), (3, 6), (1, 11), (2, 20)]


def _kq(g, tag, p, ne):
    nq = ne * (p + 1)
    return g[tag + "_kq"].reshape((nq,) * 3, order="F")   # [x, y, z] points


def test_numpy_restatement_of_the_element_loop_equals_the_reference_flow(golden):
    """pins oracle.pointwise_rhs: compute_rhs of the compiled reference's examples/flow/flow.hpp:74-101"""
    g, o = golden["flow"], Oracle()
    for p, ne in FLOW_CASES:
        tag = f"flow_p{p}_n{ne}"
        t = o.basis_tables(p, ne)
        rhs = pointwise_rhs([t] * 3, g[tag + "_u0"], flow_form(float(g[tag + "_dt"][0]), _kq(g, tag, p, ne)))
        assert rel_l2(rhs, g[tag + "_rhs"]) < 5e-15, tag


def _bump(x, y, z):
    """ads::bump(0.1, 0.5, .) (examples/flow/geometry.hpp:48-64), the initial state of flow.hpp:36-41"""
    t = np.sqrt((x - 0.5) ** 2 + (y - 0.5) ** 2 + (z - 0.5) ** 2)
    r, R = 0.05, 0.25
    h = (t - r) / (R - r)
    return np.where(t < r, 1.0, np.where(t > R, 0.0, ((h - 1) * (h + 1)) ** 2))


@pytest.mark.gpu
@pytest.mark.parametrize("p,ne", FLOW_CASES)
def test_flow_nonlinear_form_vs_reference_golden(golden, p, ne):
    g = golden["flow"]
    tag = f"flow_p{p}_n{ne}"
    dt, ns = float(g[tag + "_dt"][0]), int(g[tag + "_meta"][2])
    kq = _kq(g, tag, p, ne)
    sim = ads.flow(p, ne, ads.timesteps_config(ns, dt), permeability=kq.transpose(2, 1, 0), init_state=None)
    sim.before()
    ctx = sim.ctx
    ctx.upload(U_PREV, g[tag + "_u0"])
    ctx.compute_rhs_pointwise(PointForm.flow(dt), U_PREV, U)
    assert rel_l2(ctx.download(U), g[tag + "_rhs"]) < 1e-13
    sim.set_state(g[tag + "_u0"])
    sim.advance(ns)
    assert rel_l2(sim.state(), g[tag + "_syn"]) < ns * 1e-12
    if tag + "_shipped" in g:
        ship = ads.flow(p, ne, ads.timesteps_config(ns, dt), permeability=kq.transpose(2, 1, 0), init_state=_bump)
        ship.before()
        assert rel_l2(ship.state(), g[tag + "_shipped_init"]) < 1e-12
        ship.advance(ns)
        assert rel_l2(ship.state(), g[tag + "_shipped"]) < ns * 1e-12


def _linear_form(alpha, beta, adv, gamma, source, d3):
    def f(x, y, z=None):
        dx, dy = x - 0.5, y - 0.5
        if source == 1:
            if d3:
                return np.exp(-np.sqrt(dx * dx + dy * dy + (z - 0.5) ** 2)) + 1 + np.cos(np.pi * x) * np.cos(np.pi * y) * np.cos(np.pi * z)
            return np.exp(-np.sqrt(dx * dx + dy * dy)) + 1 + np.cos(np.pi * x) * np.cos(np.pi * y)
        return 1 + np.sin(2 * np.pi * x) * np.sin(2 * np.pi * y) * np.sin(2 * np.pi * z)

    if d3:
        def form(u, ux, uy, uz, x, y, z):
            k0 = alpha * u - (adv[0] * ux + adv[1] * uy + adv[2] * uz)
            if source:
                k0 = k0 + gamma * f(x, y, z)
            return k0, -beta[0] * ux, -beta[1] * uy, -beta[2] * uz
    else:
        def form(u, ux, uy, x, y):
            k0 = alpha * u - (adv[0] * ux + adv[1] * uy)
            if source:
                k0 = k0 + gamma * f(x, y)
            return k0, -beta[0] * ux, -beta[1] * uy
    return form, f


@pytest.mark.gpu
@pytest.mark.parametrize("nd,p,ne,source,plain", [(3, 2, 21, 1, False), (3, 3, 13, 2, False), (3, 1, 19, 0, False),
                                                   (3, 4, 9, 1, True), (3, 5, 7, 0, False), (3, 2, 35, 1, True),
                                                   (2, 3, 37, 1, False), (2, 2, 50, 0, False), (2, 5, 11, 1, True)])
def test_linear_pointwise_forms_vs_element_loop_restatement(nd, p, ne, source, plain):
    """advection + diffusion + reaction with a built-in source (with and without the test function)"""
    if nd == 2 and source == 2:
        pytest.skip("the flow forcing is 3-D")
    steps = ads.timesteps_config(1, 1e-3)
    c = ads.dim_config(p, ne)
    sim = ads.simulation_3d(c, c, c, steps) if nd == 3 else ads.simulation_2d(c, c, steps)
    ctx = sim._context()
    alpha, beta, adv, gamma = 0.9, (2e-3, 1e-3, 3e-3)[:nd], (0.3, -0.2, 0.1)[:nd], 0.05
    u0 = synthetic_state(sim.shape())
    ctx.upload(U_PREV, u0)
    ctx.compute_rhs_pointwise(PointForm.linear(alpha, beta, adv, gamma, source, plain), U_PREV, U)
    got = ctx.download(U)
    form, f = _linear_form(alpha, beta, adv, 0.0 if plain else gamma, source, nd == 3)
    tabs = [d.basis for d in sim.dims]
    want = pointwise_rhs(tabs, u0, form, plain=(lambda *a: gamma * f(*a[nd + 1:])) if (plain and source) else None)
    assert rel_l2(got, want) < 1e-13
    ctx.compute_rhs_pointwise(PointForm.linear(alpha, beta, adv, gamma, source, plain), U_PREV, U)
    assert np.array_equal(ctx.download(U), got), "the brick kernel must be deterministic"


@pytest.mark.gpu
def test_pointwise_many_bricks_colours_and_segments():
    """heat form at 150^3 p=2 (10 x 19 bricks, z segments, 8 colour launches) against the collapsed kernel, and
    p=3 at 70^3"""
    for p, ne in ((2, 150), (3, 70)):
        dt = 1e-7
        sim = ads.heat_3d(p, ne, ads.timesteps_config(1, dt))
        sim.prepare_matrices()
        u0 = synthetic_state(sim.shape())
        sim.ctx.upload(U_PREV, u0)
        sim.ctx.compute_rhs(sim.substeps()[0].form, U_PREV, U)
        want = sim.ctx.download(U)
        sim.ctx.compute_rhs_pointwise(PointForm.linear(1.0, (dt, dt, dt)), U_PREV, U)
        assert rel_l2(sim.ctx.download(U), want) < 1e-14, (p, ne)


# ---------------------------------------------------------------------- generalised ADS (per-line factors)
@pytest.mark.gpu
@pytest.mark.parametrize("nd,p,ne,axis", [(3, 2, 10, 0), (3, 2, 10, 1), (3, 3, 7, 2), (2, 3, 20, 0), (2, 2, 33, 1),
                                          (3, 5, 6, 1), (3, 4, 30, 0), (3, 1, 40, 2)])
def test_generalised_ads_with_a_factor_per_line(oracle, nd, p, ne, axis):
    """ads_solve with a special dimension (include/ads/solver.hpp:56-96,:170-195): every line along `axis` has its
    own matrix M + h_l S (the shape of examples/maxwell/maxwell_ads.hpp:139-163), the other axes the Gram factor;
    against the oracle's dgbtrs, line by line, special axis first."""
    n = ne + p
    shape = (n,) * nd
    rng = np.random.default_rng(5)
    lines = n ** (nd - 1)
    h = rng.uniform(0.0, 0.05, lines)
    lus, pivs = [], []
    for l in range(lines):
        lu, piv = ads.band_factorize(ads.matrix_1d(3, p, ne, h=float(h[l]), fix=3 if l % 3 == 0 else 0), p, p)
        lus.append(lu), pivs.append(piv)
    gram_lu, gram_piv = ads.band_factorize(ads.matrix_1d(0, p, ne), p, p)
    ctx = ads.Context(shape)
    for ax in range(nd):
        ctx.set_factor(ax, 0, gram_lu, gram_piv, p, p)
    ctx.set_line_factors(axis, np.stack(lus), np.stack(pivs), p, p)
    rhs = rng.standard_normal(n ** nd)
    ctx.upload(U, rhs)
    ctx.solve_special(U, axis)
    got = ctx.download(U).reshape(shape, order="F")
    # oracle: the special axis line by line (lines numbered over the other axes, x fastest), then the others
    X = rhs.reshape(shape, order="F").copy()
    Xs = np.moveaxis(X, axis, 0)                       # [j, other axes in order]
    flat = Xs.reshape(n, lines, order="F")             # column l = line l
    for l in range(lines):
        flat[:, l] = oracle.solve_factorized(lus[l], pivs[l], p, p, flat[:, l])
    X = np.moveaxis(flat.reshape(Xs.shape, order="F"), 0, axis)
    for ax in range(nd):
        if ax == axis:
            continue
        Y = np.moveaxis(X, ax, 0).reshape(n, -1, order="F")
        sol = oracle.solve_factorized(gram_lu, gram_piv, p, p, np.ascontiguousarray(Y.T)).reshape(-1, n).T
        X = np.moveaxis(sol.reshape(np.moveaxis(X, ax, 0).shape, order="F"), 0, ax)
    assert rel_l2(got, X) < 1e-13
