"""Generate tests/golden/flow.npz from the compiled, UNMODIFIED reference (oracle/_ref/libads_ref.so):
examples/flow/flow.hpp -- the nonlinear pointwise form (general quadrature path).

    make -C oracle ref && python tests/golden/make_golden_flow.py
Every array is an output of the reference for the stated inputs: the permeability table its
fill_permeability_map() produced (environment seed 1, examples/flow/flow.hpp:23,:53-60), its shipped initial state
(before(): projection of ads::bump(0.1, 0.5, .) + solve), compute_rhs alone and whole steps from a synthetic state.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Ref, synthetic_state  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    r = Ref()
    out = {}
    for p, ne, dt, ns in ((2, 9, 1e-7, 3), (3, 6, 1e-7, 2), (1, 11, 1e-7, 2), (2, 20, 1e-7, 1)):
        tag = f"flow_p{p}_n{ne}"
        print("generating", tag, flush=True)
        n = ne + p
        u0 = 0.05 * synthetic_state((n,) * 3)          # exp(10 u) stays moderate
        out[tag + "_meta"] = np.array([p, ne, ns], dtype=np.int64)
        out[tag + "_dt"] = np.array([dt])
        out[tag + "_u0"] = u0
        rhs, kq = r.flow(p, ne, dt, 1, u0=u0, stage=1)
        out[tag + "_kq"] = kq.astype(np.float32) if False else kq
        out[tag + "_rhs"] = rhs
        out[tag + "_syn"], _ = r.flow(p, ne, dt, ns, u0=u0)
        if ne <= 11:
            out[tag + "_shipped_init"], _ = r.flow(p, ne, dt, 0)
            out[tag + "_shipped"], _ = r.flow(p, ne, dt, ns)
    np.savez_compressed(os.path.join(OUT, "flow.npz"), **out)
    print("flow.npz", os.path.getsize(os.path.join(OUT, "flow.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
