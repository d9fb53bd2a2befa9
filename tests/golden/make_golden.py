"""Generate tests/golden/*.npz from the compiled, UNMODIFIED reference (oracle/_ref/libads_ref.so).

Run once in the build container (needs /root/reference to build oracle/_ref):
    make -C oracle ref && python tests/golden/make_golden.py
The .npz files are committed; the GPU box and the CPU test suite only read them.
Every array is an OUTPUT OF THE REFERENCE ITSELF for the stated inputs; inputs that are not
derivable from (problem, p, elements, dt, nsteps) are stored beside the outputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import NDIM, PROBLEMS, Ref, synthetic_state  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    r = Ref()
    # ---- 1-D setup data: Gauss rules, basis tables, matrices, LAPACK factors ------------------
    setup = {}
    for q in range(2, 9):
        x, w = r.gauss(q)
        setup[f"gauss_x_{q}"], setup[f"gauss_w_{q}"] = x, w
    for p, ne in ((1, 5), (2, 12), (3, 7), (4, 9), (5, 6)):
        t = r.basis_tables(p, ne)
        for k, v in t.items():
            setup[f"tab_{p}_{ne}_{k}"] = v
        for kind, h, fix in ((0, 0.0, 0), (0, 0.0, 1), (1, 0.0, 0), (2, 0.0, 0), (3, 0.005, 0),
                             (3, 3.0, 1)):
            m = r.matrix_1d(kind, p, ne, h=h, fix=fix)
            tag = f"mat_{p}_{ne}_{kind}_{fix}_{h}"
            setup[tag] = m
            if kind in (0, 3):
                f, piv, info = r.factorize(m, p, p)
                setup[tag + "_lu"], setup[tag + "_ipiv"] = f, piv
    np.savez_compressed(os.path.join(OUT, "setup.npz"), **setup)

    # ---- ads_solve on random right-hand sides (pivoting and non-pivoting factors) ------------
    solve = {}
    rng = np.random.default_rng(7)
    for tag, p, ne, kind, h, fix in (("p2", 2, 12, 0, 0.0, 0), ("p3fix", 3, 7, 0, 0.0, 1),
                                     ("p5", 5, 6, 0, 0.0, 0), ("p4K", 4, 9, 3, 3.0, 1)):
        n = ne + p
        f, piv, _ = r.factorize(r.matrix_1d(kind, p, ne, h=h, fix=fix), p, p)
        for nd in (1, 2, 3):
            rhs = rng.standard_normal(n ** nd)
            x = r.ads_solve((n,) * nd, [f] * nd, [piv] * nd, [p] * nd, [p] * nd, rhs)
            solve[f"{tag}_{nd}d_rhs"], solve[f"{tag}_{nd}d_x"] = rhs, x
        solve[f"{tag}_meta"] = np.array([p, ne, kind, fix], dtype=np.int64)
        solve[f"{tag}_h"] = np.array([h])
    # rectangular tensor with three different matrices
    shapes = (14, 10, 11)
    mats, pivs, ps = [], [], []
    for (p, ne) in ((2, 12), (3, 7), (5, 6)):
        f, piv, _ = r.factorize(r.matrix_1d(0, p, ne), p, p)
        mats.append(f), pivs.append(piv), ps.append(p)
    rhs = rng.standard_normal(int(np.prod(shapes)))
    solve["mixed_rhs"] = rhs
    solve["mixed_x"] = r.ads_solve(shapes, mats, pivs, ps, ps, rhs)
    t_in = rng.standard_normal(2 * 3 * 4)
    solve["rot_in"], solve["rot_out"] = t_in, r.cyclic_transpose((2, 3, 4), t_in)
    np.savez_compressed(os.path.join(OUT, "solve.npz"), **solve)

    # ---- whole problems ------------------------------------------------------------------------
    cases = [
        # (problem, p, elements, dt, nsteps)
        ("heat_3d", 2, 12, 1e-7, 100),       # BASELINE.json configs[0]
        ("heat_3d", 3, 5, 1e-7, 3),
        ("heat_2d", 3, 24, 1e-5, 5),
        ("heat_2d", 2, 16, 1e-5, 5),
        ("implicit_2d", 3, 24, 1e-2, 5),
        ("implicit_2d", 2, 16, 1e-2, 3),
        ("scalability_3d", 2, 8, 1e-6, 2),
        ("scalability_3d", 3, 6, 1e-6, 2),
        ("scalability_3d", 4, 6, 1e-6, 2),
        ("scalability_3d", 5, 6, 1e-6, 2),
        ("scalability_2d", 3, 16, 1e-6, 3),
        ("implicit_3d", 3, 8, 1e-2, 2),
        ("implicit_3d", 2, 6, 1e-2, 2),
    ]
    prob = {}
    for name, p, ne, dt, ns in cases:
        pid = PROBLEMS[name]
        n = ne + p
        tag = f"{name}_p{p}_n{ne}"
        print("generating", tag, flush=True)
        prob[tag + "_meta"] = np.array([pid, p, ne, ns], dtype=np.int64)
        prob[tag + "_dt"] = np.array([dt])
        u_ship, _ = r.run(name, p, ne, dt, ns)             # shipped before() + ns steps
        prob[tag + "_shipped"] = u_ship
        u_ship0, _ = r.run(name, p, ne, dt, 0)             # shipped initial state only
        prob[tag + "_shipped_init"] = u_ship0
        u0 = synthetic_state((n,) * NDIM[pid])
        prob[tag + "_u0"] = u0
        steps = min(ns, 3)
        u_syn, _ = r.run(name, p, ne, dt, steps, u0=u0)     # synthetic state + steps
        prob[tag + "_syn"] = u_syn
        prob[tag + "_syn_steps"] = np.array([steps], dtype=np.int64)
        u1, _ = r.run(name, p, ne, dt, 1, u0=u0)
        prob[tag + "_syn_step1"] = u1
        # noise floor of the comparison: how far the REFERENCE's own one-step result moves when
        # every input coefficient is changed by one unit in the last place (random direction)
        sgn = np.random.default_rng(11).choice([-1.0, 1.0], size=u0.size)
        u1p, _ = r.run(name, p, ne, dt, 1, u0=np.nextafter(u0, u0 + sgn))
        prob[tag + "_ulp_floor"] = np.array([np.linalg.norm(u1p - u1) / np.linalg.norm(u1)])
        nsub = 2 if name == "implicit_2d" else 3 if name == "implicit_3d" else 1
        for s in range(1, nsub + 1):
            rhs, _ = r.run(name, p, ne, dt, 1, u0=u0, stage=s)
            prob[tag + f"_rhs{s}"] = rhs
    np.savez_compressed(os.path.join(OUT, "problems.npz"), **prob)
    for f in ("setup.npz", "solve.npz", "problems.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
