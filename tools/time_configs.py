import time, numpy as np, sys
sys.path.insert(0, '/root/repo')
import iga_ads_b200 as ads
from iga_ads_b200 import U, U_PREV
for name, p, ne, dt in (("heat_2d", 3, 4096, 1e-5), ("heat_2d", 2, 4094, 1e-5), ("implicit_3d", 3, 256, 1e-2), ("scalability_3d", 3, 256, 1e-6)):
    try:
        sim = ads.PROBLEMS[name](p, ne, ads.timesteps_config(1, dt)); sim.prepare_matrices()
        n = ne + p; nd = len(sim.shape())
        u0 = np.random.default_rng(0).standard_normal(n ** nd); sim.set_state(u0)
        sim.advance(2); sim.ctx.synchronize()
        sim.ctx.enable_timing(True); sim.ctx.stage_times()
        t = time.perf_counter(); sim.advance(5); sim.ctx.synchronize(); el = (time.perf_counter() - t) / 5
        st = sim.ctx.stage_times()
        print(name, p, ne, "ms/step %.3f" % (el * 1e3), {k: round(v / 5, 3) for k, v in st.items()}, "finite", bool(np.isfinite(sim.state()).all()), flush=True)
    except Exception as e:
        print(name, p, ne, "FAILED", e, flush=True)
