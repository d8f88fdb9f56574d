"""BASELINE config 4 on the engine's side: bootstrap replicates at 50 taxa (230,300 quartets).  Per replicate the
reference resamples every obsCF from its credibility interval (sampleCFfromCI!, src/bootstrap.jl:52-70: three uniforms,
then renormalise), overwrites the table (readtableCF!, src/readquartetdata.jl:211-217) and runs the search; here:
host resampling (numpy, outside the engine), snaq_set_obscf (H2D of the table + weights), then what a search does many
times: snaq_optbl on a candidate network; and the device route (snaq_set_ci once, snaq_resample_obscf per replicate).
12 replicates, as one GPU's share at 8 GPUs."""
import json
import sys
import time

sys.path.insert(0, '/root/repo')
import numpy as np

import snaq_jl_b200 as S

n, h, reps = 50, 2, 12
net = S.simulate.random_level1_network(n, h, 20261022)
taxa = [f"t{i + 1}" for i in range(n)]
qt = S.simulate.all_quartets(n)
eng = S.Engine(0)
eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
flat = net.flatten(taxa)
eng.set_network(flat)
x0, _, upper = eng.get_params()
obs = S.simulate.obs_from_expcf(eng.eval_quartets(x0)[0], 300, 4)
lo, hi = np.clip(obs - 0.05, 0, 1), np.clip(obs + 0.05, 0, 1)             # CI = obs -/+ 0.05 clipped
rng = np.random.default_rng(0)
eng.set_obscf(obs)
eng.optbl(x0)                                                              # warm-up
t_host = t_set = t_opt = 0.0
fmins = []
for r in range(reps):
    t0 = time.perf_counter()
    c = (hi - lo) * rng.uniform(size=obs.shape) + lo
    c /= c.sum(1, keepdims=True)
    t1 = time.perf_counter()
    eng.set_obscf(c)
    t2 = time.perf_counter()
    xm, info = eng.optbl(x0)
    t3 = time.perf_counter()
    t_host += t1 - t0; t_set += t2 - t1; t_opt += t3 - t2
    fmins.append(info["fmin"])
# the same on the device: the intervals stay in HBM, a replicate is one kernel (snaq_resample_obscf) + the weights
eng.set_ci(lo, hi)
eng.resample_obscf(0)
t0 = time.perf_counter()
for r in range(reps):
    eng.resample_obscf(1000 + r)
t_dev = (time.perf_counter() - t0) / reps
print(json.dumps({"ntaxa": n, "nq": len(qt), "replicates": reps, "device_resampling_ms": t_dev * 1e3,
                  "host_resampling_ms": t_host / reps * 1e3,
                  "set_obscf_ms": t_set / reps * 1e3, "optbl_ms": t_opt / reps * 1e3,
                  "table_bytes": int(obs.nbytes), "fmin_range": [min(fmins), max(fmins)]}))
