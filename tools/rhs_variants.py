#!/usr/bin/env python
"""Time the collapsed right-hand side alone for a list of kernel variants (one subprocess per variant,
the variant is an environment variable read once by libadsb200) and check every variant against the
cp.async kernel (ADSB_RHS_VARIANT=10) on the same input.

    python tools/rhs_variants.py [--p 2] [--elements 512] [--problem heat_3d] 0:0 0:1 0:2 10:0 ...

Each spec is ADSB_RHS_VARIANT:ADSB_RHS_TMA_VARIANT.  Needs a GPU."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(args):
    import numpy as np
    import torch  # noqa: F401  (CUDA context / allocator warm-up is not needed; events come from the library)

    import iga_ads_b200 as ads
    from iga_ads_b200 import U, U_PREV, _lib

    p, ne = args.p, args.elements
    sim = ads.PROBLEMS[args.problem](p, ne, ads.timesteps_config(1, args.dt))
    sim.prepare_matrices()
    shape = sim.shape()
    N = int(np.prod(shape))
    rng = np.random.default_rng(1)
    u0 = rng.standard_normal(N)
    ctx = sim.ctx
    ctx.upload(U_PREV, u0)
    dt = args.dt
    form = _lib.Form.make(1.0, (dt,) * len(shape))
    for _ in range(3):
        ctx.compute_rhs(form, U_PREV, U)
    ctx.synchronize()
    ctx.enable_timing(True)
    ctx.stage_times()
    reps = 10
    for _ in range(reps):
        ctx.compute_rhs(form, U_PREV, U)
    ctx.synchronize()
    ms = ctx.stage_times()["rhs"] / reps
    out = ctx.download(U)
    np.save(args.out, out[:: args.stride].copy())
    print(json.dumps({"ms": ms, "gbs": 16 * N / ms / 1e6, "finite": bool(np.isfinite(out).all()),
                      "norm": float(np.linalg.norm(out))}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--p", type=int, default=2)
    ap.add_argument("--elements", type=int, default=512)
    ap.add_argument("--problem", default="heat_3d")
    ap.add_argument("--dt", type=float, default=1e-7)
    ap.add_argument("--stride", type=int, default=7)
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--out", default="")
    ap.add_argument("specs", nargs="*")
    args = ap.parse_args()
    if args.child:
        child(args)
        return
    import numpy as np

    specs = args.specs or ["10:0", "0:0"]
    if "10:0" not in specs:
        specs = ["10:0"] + specs
    ref = None
    for spec in specs:
        v, t = spec.split(":")
        env = dict(os.environ, ADSB_RHS_VARIANT=v, ADSB_RHS_TMA_VARIANT=t)
        out = f"/tmp/rhs_variant_{v}_{t}.npy"
        cmd = [sys.executable, os.path.abspath(__file__), "--child", "--out", out, "--p", str(args.p), "--elements",
               str(args.elements), "--problem", args.problem, "--dt", str(args.dt), "--stride", str(args.stride)]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
        if r.returncode != 0:
            print(spec, "FAILED", r.stderr[-600:], flush=True)
            continue
        res = json.loads(r.stdout.strip().splitlines()[-1])
        a = np.load(out)
        if spec == "10:0":
            ref = a
        err = float(np.linalg.norm(a - ref) / np.linalg.norm(ref)) if ref is not None else None
        print(f"{args.problem} p={args.p} n={args.elements} variant {spec}: {res['ms']:.4f} ms  {res['gbs']:.0f} GB/s  "
              f"finite={res['finite']}  rel diff vs cp.async kernel = {err}", flush=True)


if __name__ == "__main__":
    main()
