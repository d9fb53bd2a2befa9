#!/usr/bin/env python
"""bench.py -- ADS step throughput (DOF-updates/s) for heat_3d 512^3 p=2 on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--elements E] [--p P]

One JSON line on stdout (rank 0).  A "step" is one explicit ADS time step (right-hand side by
pre-integrated sum factorisation + three batched banded sweeps) on a synthetic coefficient tensor
(SURVEY.md 8d).  `value` times K steps with the state resident in HBM (CUDA events on the
library's stream, barrier + synchronize on both sides, max over ranks); `e2e` times the same step
through the public API with HOST buffers (pinned upload of u, step, download of u, every step).
`--impl reference` times the reference's own CPU implementation (oracle/_ref, compiled from the
unmodified reference sources) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ADS step DOF-updates/s (heat_3d p=2)"
UNIT = "DOF/s"
BYTES_PER_DOF_STEP = 64  # RHS 16 B + 3 sweeps x 16 B
BYTES_PER_DOF_SWEEP = 16


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def measured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/r1c_traffic.json), valid
    for the default workload only; None when the file is missing."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1c_traffic.json")) as f:
            return float(json.load(f)[kernel])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def synthetic_local(n, lo, cnt, seed=20260101):
    """Synthetic coefficient tensor of SURVEY.md 8d restricted to a box (x fastest)."""
    import numpy as np

    rng = np.random.default_rng(seed + 7919 * lo[2])
    x = np.linspace(0.0, 1.0, n[0])[lo[0]:lo[0] + cnt[0]]
    y = np.linspace(0.0, 1.0, n[1])[lo[1]:lo[1] + cnt[1]]
    z = np.linspace(0.0, 1.0, n[2])[lo[2]:lo[2] + cnt[2]]
    u = (1 + z[:, None, None] ** 2) * np.cos(2 * y)[None, :, None] * np.sin(3 * x)[None, None, :]
    u += 0.1 * rng.uniform(-1, 1, size=u.shape)
    return np.ascontiguousarray(u).ravel()


def run_reference(args):
    """The reference's own CPU step (unmodified sources in oracle/_ref; else the C restatement)."""
    import numpy as np

    from oracle.oracle import Oracle, Ref, synthetic_state

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    p, ne, dt = args.p, args.ref_elements, 1e-7
    if Ref.available():
        impl, kind = Ref(), "reference"
    else:
        impl, kind = Oracle(), "port"
    n = ne + p
    u0 = synthetic_state((n, n, n))
    for _ in range(args.warmup and 1):
        impl.run("heat_3d", p, ne, dt, 1, u0=u0)
    t0 = time.perf_counter()
    u, _ = impl.run("heat_3d", p, ne, dt, args.steps, u0=u0)
    dt_s = time.perf_counter() - t0
    assert np.isfinite(u).all()
    value = n ** 3 * args.steps / dt_s
    sample = f"heat_3d p={p} {ne}^3 elements ({n**3} DOF), {args.steps} steps, shipped compute_rhs + ads_solve, 1 thread"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt_s / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"heat_3d p={p} {args.elements}^3 (timed on a {ne}^3 sample; cost is linear in DOF)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def cpu_baseline(p, ne=24, steps=2):
    import numpy as np

    from oracle.oracle import Oracle, Ref, synthetic_state

    impl, kind = (Ref(), "reference") if Ref.available() else (Oracle(), "port")
    n = ne + p
    u0 = synthetic_state((n, n, n))
    t0 = time.perf_counter()
    u, _ = impl.run("heat_3d", p, ne, 1e-7, steps, u0=u0)
    el = time.perf_counter() - t0
    assert np.isfinite(u).all()
    return {"value": n ** 3 * steps / el, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"heat_3d p={p} {ne}^3 elements, {steps} steps, examples/heat/heat_3d.hpp as shipped "
                      f"(sequential loop), {el:.1f} s on {os.cpu_count()} host cores available"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--elements", type=int, default=512)
    ap.add_argument("--p", type=int, default=2)
    ap.add_argument("--ref-elements", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        if args.steps > 4:
            args.steps = 4  # bounded sample: ~1.5 s per 32^3 step on one core
        run_reference(args)
        return

    import numpy as np
    import torch

    import iga_ads_b200 as ads
    from iga_ads_b200 import U, U_PREV

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libadsb200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from iga_ads_b200.sharded import run_sharded_bench

        run_sharded_bench(args, rank, world, local_rank)
        return

    p, ne, dt = args.p, args.elements, 1e-7
    n = ne + p
    N = n ** 3
    hbm, peak_kind = peaks()
    sim = ads.heat_3d(p, ne, ads.timesteps_config(args.steps, dt), device=local_rank)
    stream = torch.cuda.current_stream()
    sim._context().set_stream(stream.cuda_stream)
    sim.prepare_matrices()
    ctx = sim.ctx
    u0 = synthetic_local((n, n, n), (0, 0, 0), (n, n, n))
    ctx.upload(U, u0)
    ctx.upload(U_PREV, u0)

    # ---- device-resident throughput
    sim.advance(args.warmup)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    sim.advance(args.steps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop()
    value = N * args.steps / (ms * 1e-3)

    # ---- per-stage device times (separate pass, CUDA events around every launch on the same stream)
    ctx.enable_timing(True)
    ctx.stage_times()
    sim.advance(args.steps)
    st = ctx.stage_times()
    ctx.enable_timing(False)
    stages = {k: v / args.steps for k, v in st.items() if k != "other"}
    sweep_ms = (stages["sweep_x"] + stages["sweep_y"] + stages["sweep_z"]) / 3
    achieved = BYTES_PER_DOF_SWEEP * N / (sweep_ms * 1e-3) / 1e9
    default_workload = (p, ne) == (2, 512)
    roofline = {"bound": "hbm", "kernel": "sweep_tile_kernel (K2, 3 launches per step, 16 B/DOF each)", "achieved": achieved,
                "peak": hbm, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / hbm,
                "traffic": measured_traffic("sweep_tile_kernel") if default_workload else None,
                "algorithmic_bytes_per_launch": BYTES_PER_DOF_SWEEP * N,
                "step_frac": BYTES_PER_DOF_STEP * N / (ms * 1e-3 / args.steps) / 1e9 / hbm,
                "stage_ms": stages,
                "stage_gbs": {k: BYTES_PER_DOF_SWEEP * N / (v * 1e-3) / 1e9 for k, v in stages.items()}}
    state_ok = bool(np.isfinite(ctx.download(U)[:: 4099]).all())

    # ---- end to end through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        host_in = torch.from_numpy(u0).pin_memory()
        host_out = torch.empty_like(host_in).pin_memory()
        k2 = max(2, min(args.steps, 3))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(k2):
            ctx.upload(U, host_in.numpy())
            sim.advance(1)
            check_out = host_out.numpy()
            ads._lib.check(ctx.lib.adsb_download(ctx.h, U, ads._lib.d_(check_out)))
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        e2e = {"value": N * k2 / el, "unit": UNIT, "h2d_bytes_per_step": 8 * N, "d2h_bytes_per_step": 8 * N,
               "steps": k2, "ms_per_step": 1e3 * el / k2,
               "note": "adsb_upload(u) + adsb_step + adsb_download(u) per step, pinned host buffers"}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"heat_3d p={p} {ne}^3 elements ({N} DOF), explicit ADS step, dt={dt}",
                   "rhs": "collapsed (pre-integrated sum factorisation)",
                   "l2": f"state {8 * N / 1e9:.2f} GB per tensor >> 126 MB L2, no flush needed between steps",
                   "timing": "CUDA events on the library's stream around K steps, synchronize on both sides",
                   "parallelism": "1 GPU"},
        "roofline": roofline, "clocks": clocks, "gpu_launches": launches, "finite": state_ok,
    }
    if e2e:
        out["e2e"] = e2e
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(p, ne=32)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
