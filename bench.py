#!/usr/bin/env python
"""bench.py -- ADS step throughput (DOF-updates/s) on N B200s; default: heat_3d 512^3 p=2.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--config heat_3d|heat_2d|implicit_3d|scalability_3d] [--p P] [--elements E]

One JSON line on stdout (rank 0).  A "step" is one ADS time step of the named problem (every sub-step:
right-hand side by pre-integrated sum factorisation + one batched banded sweep per axis) on a synthetic
coefficient tensor.  `value` times K steps with the state resident in HBM (CUDA events on
the library's stream, barrier + synchronize on both sides, max over ranks); `e2e` times the same step
through the public API with HOST buffers: upload of the step's input from pinned memory, step, download of
its result, every step (pipelined over three streams; the serial figure is reported beside it).
`roofline` is the whole step against the measured HBM copy bandwidth, with every kernel's own fraction
beside it.  `--impl reference` times the reference's own CPU implementation (oracle/_ref, compiled from the
unmodified reference sources) on a bounded sample of the same workload.  N > 1 (torchrun): z-slabs with the
distributed z substitution (iga_ads_b200/slab.py); the line then carries `parity` (a 62^3 run of the same
code path against the CPU oracle) and `checksum` (sum and norm of the final state, equal across N).
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

UNIT = "DOF/s"
# BASELINE.json configs: problem, p, elements, dt, algorithmic bytes per DOF and step
# (right-hand side 16 B + 16 B per axis sweep, per sub-step), sub-steps per step
CONFIGS = {
    "heat_3d": dict(problem="heat_3d", p=2, elements=512, dt=1e-7, bytes=64, ndim=3, nsub=1),
    "heat_2d": dict(problem="heat_2d", p=3, elements=4096, dt=1e-5, bytes=48, ndim=2, nsub=1),
    "implicit_3d": dict(problem="implicit_3d", p=3, elements=256, dt=1e-2, bytes=192, ndim=3, nsub=3),
    "scalability_3d": dict(problem="scalability_3d", p=2, elements=768, dt=1e-6, bytes=64, ndim=3, nsub=1),
}
BYTES_PER_DOF_PASS = 16


def metric_name(cfg):
    return f"ADS step DOF-updates/s ({cfg['problem']} p={cfg['p']})"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def static_traffic(cfg):
    """DRAM bytes per step from the committed ncu capture of this command (profiles/traffic.json); only valid
    for the exact workload it was captured on, None otherwise.  Static: not measured in this run."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        key = f"{cfg['problem']}_p{cfg['p']}_n{cfg['elements']}"
        return t.get(key)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index, self.first = [], None, index, 0

    def start(self, wait_s=5.0):
        """starts nvidia-smi sampling every 50 ms and returns once the first sample has arrived (nvidia-smi takes a
        moment to come up; a timed region of a few milliseconds would otherwise be over before it)"""
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < wait_s:
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def mark(self):
        """samples taken from here on count as `under load`"""
        self.first = len(self.rows)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        time.sleep(0.06)  # let the last sample of the loaded region arrive
        for r in self.rows[self.first:] or self.rows[-1:]:
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
    """Synthetic coefficient tensor of SURVEY.md 8d restricted to a box (x fastest).  The noise of every
    slowest-axis plane has its own seed, so any slab decomposition sees the same global state."""
    import numpy as np

    nd = len(n)
    axes = [np.linspace(0.0, 1.0, n[d])[lo[d]:lo[d] + cnt[d]] for d in range(nd)]
    if nd == 3:
        u = (1 + axes[2][:, None, None] ** 2) * np.cos(2 * axes[1])[None, :, None] * np.sin(3 * axes[0])[None, None, :]
    else:
        u = np.cos(2 * axes[1])[:, None] * np.sin(3 * axes[0])[None, :]
    u = np.ascontiguousarray(u)
    plane = (n[1], n[0]) if nd == 3 else (n[0],)
    sl = (slice(lo[1], lo[1] + cnt[1]), slice(lo[0], lo[0] + cnt[0])) if nd == 3 else (slice(lo[0], lo[0] + cnt[0]),)
    for k in range(cnt[-1]):
        rng = np.random.default_rng(seed * 1009 + lo[-1] + k)
        u[k] += 0.1 * rng.uniform(-1, 1, size=plane)[sl]
    return u.ravel()


# ------------------------------------------------------------------------------ reference arm / CPU baseline
def _ref_impl():
    from oracle.oracle import Oracle, Ref

    return (Ref(), "reference") if Ref.available() else (Oracle(), "port")


def reference_sample(cfg, ne, steps, warm=0):
    """The reference's own step on a bounded sample (`ne` elements per axis): examples/<problem> exactly as
    shipped.  heat_3d's compute_rhs is a plain sequential loop (examples/heat/heat_3d.hpp:49-67), so it can use
    exactly one thread."""
    import numpy as np

    from oracle.oracle import synthetic_state

    impl, kind = _ref_impl()
    n = ne + cfg["p"]
    u0 = synthetic_state((n,) * cfg["ndim"])
    for _ in range(warm):
        impl.run(cfg["problem"], cfg["p"], ne, cfg["dt"], 1, u0=u0)
    t0 = time.perf_counter()
    u, _ = impl.run(cfg["problem"], cfg["p"], ne, cfg["dt"], steps, u0=u0)
    el = time.perf_counter() - t0
    assert np.isfinite(u).all()
    return n ** cfg["ndim"] * steps / el, el, kind


def hoisted_split(p, ne, steps, threads):
    """RHS / solve split on the reference's own benchmark class (examples/scalability/test3d.hpp:66-95: eval_fun
    hoisted out of the test-function loop, executor.for_each over elements); threads > 1 uses the std::thread
    stand-in for the Galois executor (oracle/shim/ads/executor/galois.hpp; Galois itself is not installed)."""
    import numpy as np

    from oracle.oracle import Ref, synthetic_state, _d

    if not Ref.available():
        return None
    r = Ref()
    n = ne + p
    u = synthetic_state((n, n, n)).copy()
    tm = np.zeros(8)
    r.set_threads(threads)
    try:
        r.lib.ref_time_heat3d_hoisted(p, ne, 1e-7, steps, _d(u), _d(tm))
    finally:
        r.set_threads(0)
    N = n ** 3
    return {"threads": threads, "elements": ne, "steps": steps, "dof_per_s": N * steps / tm[0],
            "rhs_dof_per_s": N * steps / tm[1], "solve_dof_per_s": N * steps / tm[3],
            "rhs_s_per_step": tm[1] / steps, "solve_s_per_step": tm[3] / steps}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ne = args.ref_elements or (32 if cfg["ndim"] == 3 else 256)
    steps = max(1, min(args.steps, 4))
    value, el, kind = reference_sample(cfg, ne, steps, warm=1 if args.warmup else 0)
    n = ne + cfg["p"]
    sample = (f"{cfg['problem']} p={cfg['p']} {ne}^{cfg['ndim']} elements ({n ** cfg['ndim']} DOF), {steps} steps, "
              f"compute_rhs + ads_solve as shipped, 1 thread (the shipped element loop is sequential), {el:.1f} s")
    print(json.dumps({
        "impl": "reference", "metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{cfg['problem']} p={cfg['p']} {cfg['elements']}^{cfg['ndim']} (timed on a "
                               f"{ne}^{cfg['ndim']} sample; cost is linear in DOF)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def cpu_baseline(cfg):
    """the compiled reference on this box's host cores, bounded samples: the shipped
    class (the `value`), and for the 3-D configs the hoisted benchmark class split into RHS and solve, on one
    thread and on all cores."""
    ne = 32 if cfg["ndim"] == 3 else 512
    if cfg["p"] >= 3 and cfg["ndim"] == 3:
        ne = 16
    value, el, kind = reference_sample(cfg, ne, 1)
    out = {"value": value, "unit": UNIT, "cores": 1, "kind": kind,
           "sample": f"{cfg['problem']} p={cfg['p']} {ne}^{cfg['ndim']} elements, 1 step, examples/ class as shipped, "
                     f"1 thread, {el:.1f} s; {os.cpu_count()} host cores available"}
    if cfg["ndim"] == 3:
        nh = 48 if cfg["p"] == 2 else 24
        try:
            out["hoisted_form"] = [hoisted_split(cfg["p"], nh, 1, 1), hoisted_split(cfg["p"], nh, 1, os.cpu_count() or 1)]
        except Exception as e:  # noqa: BLE001 -- the split is an extra; never lose the bench line over it
            out["hoisted_form"] = f"unavailable: {e}"
    return out


# ------------------------------------------------------------------------------ one GPU
def run_single(args, cfg, local_rank):
    import numpy as np
    import torch

    import iga_ads_b200 as ads
    from iga_ads_b200 import U, U_PREV
    from iga_ads_b200.simulation import SCRATCH

    p, ne, dt, nd = cfg["p"], cfg["elements"], cfg["dt"], cfg["ndim"]
    n = ne + p
    N = n ** nd
    hbm, peak_kind = peaks()
    sim = ads.PROBLEMS[cfg["problem"]](p, ne, ads.timesteps_config(args.steps, dt), device=local_rank)
    stream = torch.cuda.current_stream()
    sim._context().set_stream(stream.cuda_stream)
    sim.prepare_matrices()
    ctx = sim.ctx
    shape = (n,) * nd
    u0 = synthetic_local(shape, (0,) * nd, shape)
    ctx.upload(U, u0)
    ctx.upload(U_PREV, u0)

    # ---- device-resident throughput
    sampler = ClockSampler(local_rank)
    sampler.start()
    sim.advance(args.warmup)
    torch.cuda.synchronize()
    # clocks under load: the timed region lasts only milliseconds, so the same steps keep running for ~0.3 s
    # right before it (untimed, state is restored) while nvidia-smi samples; the timed steps follow back to back
    sampler.mark()
    t_load = time.time()
    extra = 0
    while time.time() - t_load < 0.3:
        sim.advance(5)
        extra += 5
        torch.cuda.synchronize()
    ctx.upload(U, u0)
    ctx.upload(U_PREV, u0)
    warm = args.warmup + (args.warmup & 1)   # even, as the multi-GPU leg needs it: the checksums stay comparable
    sim.advance(warm)
    torch.cuda.synchronize()
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
    final = ctx.download(U)
    checksum = {"steps": warm + args.steps, "sum": float(final.sum()), "l2": float(np.linalg.norm(final))}
    state_ok = bool(np.isfinite(final).all())
    del final

    # ---- per-kernel device times (separate pass, CUDA events around every launch on the same stream)
    ctx.enable_timing(True)
    ctx.stage_times()
    sim.advance(args.steps)
    st = ctx.stage_times()
    ctx.enable_timing(False)
    names = ["rhs", "sweep_x", "sweep_y"] + (["sweep_z"] if nd == 3 else [])
    step_s = ms * 1e-3 / args.steps
    per_kernel = {}
    for k in names:
        t = st[k] / args.steps / cfg["nsub"]  # ms per launch group (one per sub-step)
        gbs = BYTES_PER_DOF_PASS * N / (t * 1e-3) / 1e9
        per_kernel[k] = {"ms": t, "gbs": gbs, "frac": gbs / hbm}
    slowest = min(per_kernel, key=lambda k: per_kernel[k]["frac"])
    achieved = cfg["bytes"] * N / step_s / 1e9
    traffic = static_traffic(cfg)
    roofline = {"bound": "hbm", "kernel": f"whole step ({launches // args.steps} launches)", "achieved": achieved,
                "peak": hbm, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / hbm,
                "traffic": traffic, "traffic_source": "static: ncu capture of this workload, profiles/traffic.json"
                if traffic else None,
                "algorithmic_bytes_per_step": cfg["bytes"] * N, "per_kernel": per_kernel,
                "slowest_kernel": slowest, "sum_of_kernels_ms": sum(v["ms"] for v in per_kernel.values()) * cfg["nsub"]}

    # ---- end to end through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        host_in = torch.from_numpy(u0).pin_memory()
        host_out = [torch.empty_like(host_in).pin_memory() for _ in range(2)]
        k2 = max(2, min(args.steps, 3))
        # serial: upload, step, download on one stream
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(k2):
            ctx.upload(U, host_in.numpy())
            sim.advance(1)
            ads._lib.check(ctx.lib.adsb_download(ctx.h, U, ads._lib.d_(host_out[0].numpy())))
        torch.cuda.synchronize()
        serial = time.perf_counter() - t0
        # pipelined: the upload of input k+1 and the download of result k-1 run on their own streams (both PCIe
        # directions) while step k computes; three managed buffers rotate through U by pointer swaps
        NEXT, DONE = SCRATCH, SCRATCH + 1
        up, down = torch.cuda.Stream(), torch.cuda.Stream()
        k3 = 6
        ev_up = [torch.cuda.Event() for _ in range(k3 + 1)]
        ev_step = [torch.cuda.Event() for _ in range(k3)]
        ev_down = [torch.cuda.Event() for _ in range(k3)]
        ctx.upload(NEXT, host_in.numpy())      # allocates the two extra buffers (outside the timed region)
        ctx.upload(DONE, host_in.numpy())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.upload_async(NEXT, host_in.data_ptr(), up.cuda_stream)
        ev_up[0].record(up)
        for k in range(k3):
            stream.wait_event(ev_up[k])              # input k has landed in NEXT
            ctx.swap(U, NEXT)                        # U <- input k
            sim.advance(1)                           # result k in U (the step swaps U / U_PREV itself)
            ev_step[k].record(stream)
            ctx.swap(U, DONE)                        # DONE <- result k
            down.wait_event(ev_step[k])
            ctx.download_async(DONE, host_out[k & 1].data_ptr(), down.cuda_stream)
            ev_down[k].record(down)
            if k + 1 < k3:
                if k >= 2:
                    up.wait_event(ev_down[k - 2])    # NEXT's memory was the source of download k-2
                ctx.upload_async(NEXT, host_in.data_ptr(), up.cuda_stream)
                ev_up[k + 1].record(up)
        torch.cuda.synchronize()
        piped = time.perf_counter() - t0
        e2e = {"value": N * k3 / piped, "unit": UNIT, "h2d_bytes_per_step": 8 * N, "d2h_bytes_per_step": 8 * N,
               "steps": k3, "ms_per_step": 1e3 * piped / k3, "serial_ms_per_step": 1e3 * serial / k2,
               "note": "per step: adsb_upload_async(input) | adsb_step | adsb_download_async(result) on three streams, "
                       "pinned host buffers; serial = the same three calls back to back on one stream"}

    # ---- the general quadrature kernel (north star (a): pointwise forms by sum-factorised Gauss quadrature): its bound is
    # the FP64 pipe, not HBM.  Timed alone on a 256^3 problem of the same form; flop model of SURVEY.md 8d:
    # 2 (2 m^2 + 3 m^3 + 4 m^4) FMA-flops + 10 m^3 pointwise flops per DOF, m = p + 1.
    quad = None
    if nd == 3 and not args.no_quadrature:
        try:
            neq = min(ne, 256)
            qs = ads.PROBLEMS[cfg["problem"]](p, neq, ads.timesteps_config(1, dt), method=ads.RHS_QUADRATURE, device=local_rank)
            qs._context().set_stream(stream.cuda_stream)
            qs.prepare_matrices()
            nq = neq + p
            qs.ctx.upload(U_PREV, synthetic_local((nq,) * 3, (0, 0, 0), (nq,) * 3))
            form = qs.substeps()[0].form
            for _ in range(2):
                qs.ctx.compute_rhs(form, U_PREV, U)
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            q0.record(stream)
            for _ in range(5):
                qs.ctx.compute_rhs(form, U_PREV, U)
            q1.record(stream)
            torch.cuda.synchronize()
            qms = q0.elapsed_time(q1) / 5
            m = p + 1
            flop_per_dof = 2 * 2 * (2 * m ** 2 + 3 * m ** 3 + 4 * m ** 4) + 10 * m ** 3
            fp64_peak = 37.0  # TFLOP/s, measured FMA-chain peak of this pool's B200 (profiles/r1_ubench_fp64.txt)
            tf = flop_per_dof * nq ** 3 / (qms * 1e-3) / 1e12
            quad = {"kernel": "quad_brick_kernel (ADSB_RHS_QUADRATURE: init + brick quadrature, 8 colour launches, deterministic)",
                    "elements": neq, "dof": nq ** 3, "ms": qms, "dof_per_s": nq ** 3 / (qms * 1e-3),
                    "flop_per_dof_model": flop_per_dof, "tflops": tf, "bound": "fp64", "peak_tflops": fp64_peak,
                    "frac": tf / fp64_peak}
            qs.ctx.close()
        except Exception as e:  # noqa: BLE001 -- an extra record; never lose the bench line over it
            quad = {"error": str(e)}

    out = {
        "metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{cfg['problem']} p={p} {ne}^{nd} elements ({N} DOF), ADS step "
                               f"({cfg['nsub']} sub-step{'s' if cfg['nsub'] > 1 else ''}), dt={dt}",
                   "rhs": "collapsed (pre-integrated sum factorisation)",
                   "l2": f"state {8 * N / 1e9:.2f} GB per tensor vs 126 MB L2: inputs exceed L2, no flush between steps",
                   "timing": "CUDA events on the library's stream around K steps, synchronize on both sides",
                   "parallelism": "1 GPU"},
        "roofline": roofline, "clocks": clocks, "gpu_launches": launches, "finite": state_ok, "checksum": checksum,
    }
    if quad:
        out["rhs_quadrature"] = quad
    if e2e:
        out["e2e"] = e2e
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(cfg)
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="heat_3d", choices=sorted(CONFIGS))
    ap.add_argument("--elements", type=int, default=0)
    ap.add_argument("--p", type=int, default=0)
    ap.add_argument("--ref-elements", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-quadrature", action="store_true")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.p:
        cfg["p"] = args.p
    if args.elements:
        cfg["elements"] = args.elements
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3  # timing rule: at least three untimed warm-up steps

    if args.impl == "reference":
        run_reference(args, cfg)
        return

    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libadsb200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist

        if cfg["ndim"] != 3:
            raise SystemExit("the 2-D configuration is a 1-GPU configuration (BASELINE.json configs[1])")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from iga_ads_b200.slab_bench import run_multi_gpu_bench

        run_multi_gpu_bench(args, cfg, rank, world, local_rank)
        return
    run_single(args, cfg, local_rank)


if __name__ == "__main__":
    main()
