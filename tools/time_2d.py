import time, numpy as np, sys
sys.path.insert(0, '/root/repo')
import iga_ads_b200 as ads
for name, p, ne, dt in (("heat_2d", 2, 4094, 1e-5), ("heat_2d", 3, 4096, 1e-5)):
    sim = ads.PROBLEMS[name](p, ne, ads.timesteps_config(1, dt)); sim.prepare_matrices()
    n = ne + p
    sim.set_state(np.random.default_rng(0).standard_normal(n * n))
    sim.advance(2); sim.ctx.synchronize(); sim.ctx.enable_timing(True); sim.ctx.stage_times()
    sim.advance(5); sim.ctx.synchronize(); st = sim.ctx.stage_times()
    print(name, p, ne, {k: round(v / 5, 3) for k, v in st.items() if v}, bool(np.isfinite(sim.state()).all()), flush=True)
