"""Times the brick quadrature kernel (csrc/quadbrick.cuh) against its FP64 flop model.
    python tools/time_quadrature.py [p:elements ...]      default: 2:256 3:192 2:512 + the flow form at 2:256
Flop model: 2 (2 m^2 + 3 m^3 + 4 m^4) FMA + 10 m^3 pointwise flop per DOF, m = p + 1; the
measured FMA-chain peak of this pool's B200 is 37 TFLOP/s (profiles/r1_ubench_fp64.txt)."""
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
import iga_ads_b200 as ads  # noqa: E402
from iga_ads_b200 import PointForm, U, U_PREV  # noqa: E402

PEAK = 37.0


def flops_per_dof(p):
    m = p + 1
    return 2 * 2 * (2 * m ** 2 + 3 * m ** 3 + 4 * m ** 4) + 10 * m ** 3


def run(p, ne, flow=False, reps=5):
    dt = 1e-7
    c = ads.dim_config(p, ne)
    sim = ads.simulation_3d(c, c, c, ads.timesteps_config(1, dt))
    ctx = sim._context()
    n = ne + p
    rng = np.random.default_rng(0)
    ctx.upload(U_PREV, 0.05 * rng.standard_normal(n ** 3))
    if flow:
        ctx.set_point_coefficient(1.0 + rng.random((ne * (p + 1)) ** 3))
        form = PointForm.flow(dt)
    else:
        form = PointForm.linear(1.0, (dt, dt, dt))
    l0 = ctx.launch_count()
    ctx.compute_rhs_pointwise(form, U_PREV, U)
    launches = ctx.launch_count() - l0
    ctx.compute_rhs_pointwise(form, U_PREV, U)
    ctx.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ctx.compute_rhs_pointwise(form, U_PREV, U)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    tf = flops_per_dof(p) * n ** 3 / (ms * 1e-3) / 1e12
    print(f"{'flow' if flow else 'heat'} form p={p} {ne}^3: {ms:.3f} ms, {launches} launches, {n ** 3 / ms / 1e6:.2f} GDOF/s, "
          f"{tf:.2f} TFLOP/s (model) = {tf / PEAK:.3f} of the measured FP64 peak", flush=True)
    ctx.close()


if __name__ == "__main__":
    cases = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(2, 256), (3, 192), (2, 512)]
    for p, ne in cases:
        run(p, ne)
    if len(sys.argv) == 1:
        run(2, 256, flow=True)
