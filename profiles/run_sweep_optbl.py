"""Randomized check of snaq_optbl on the GPU against a tight bound-constrained optimisation (scipy L-BFGS-B with numerical
gradients) of the ORACLE objective from the same start: how often the device optimum is at least as good, and by how much
it is worse when it is not (different local optimum).  Usage: python profiles/run_sweep_optbl.py [n_seeds]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.optimize import minimize

import snaq_jl_b200 as S
from snaq_jl_b200 import simulate
from oracle import oracle

nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
better = equal = worse = 0
worst = 0.0
for seed in range(nseeds):
    rng = np.random.default_rng(seed)
    n, h = int(rng.integers(7, 15)), int(rng.integers(0, 3))
    try:
        net = simulate.random_level1_network(n, h, 4000 + seed, min_cycle=3 if seed % 2 else 4)
    except S.SnaqError:
        continue
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = simulate.all_quartets(n)
    flat = net.flatten(taxa)
    p = oracle.parameters(flat)
    truth_cf = oracle.extract(flat, qt)["expcf"]
    # half of the cases: data simulated on ANOTHER network (the situation of a candidate topology in the search)
    if seed % 2:
        other = None
        for attempt in range(40):
            try:
                other = simulate.random_level1_network(n, h, 9000 + seed + 1000 * attempt)
                break
            except S.SnaqError:
                pass
        if other is None:
            continue
        truth_cf = oracle.extract(other.flatten(taxa), qt)["expcf"]
    obs = simulate.obs_from_expcf(truth_cf, 100, seed)
    eng = S.Engine(0)
    eng.set_data(S.DataCF(taxa, qt, obs))
    eng.set_network(flat)
    x0, upper = p["x"], p["upper"]
    start = np.clip(x0 * rng.uniform(0.7, 1.3, len(x0)), 1e-3, upper - 1e-3)
    fo = lambda x: float(oracle.evaluate(flat, qt, obs, x.reshape(1, -1))[0])
    ref = minimize(fo, start, method="L-BFGS-B", bounds=list(zip(np.zeros_like(upper), upper)),
                   options=dict(ftol=1e-15, gtol=1e-9, maxiter=2000, maxfun=200000))
    xmin, res = eng.optbl(start, ftol_rel=1e-12, ftol_abs=1e-10, xtol_rel=1e-10, xtol_abs=1e-10)
    assert abs(fo(xmin) - res["fmin"]) <= 1e-10 * max(1.0, abs(res["fmin"])), (seed, 'value at xmin')
    assert np.all(xmin >= 0) and np.all(xmin <= upper), (seed, 'bounds')
    d = (res["fmin"] - ref.fun) / max(1.0, abs(ref.fun))
    if d < -1e-9:
        better += 1
    elif d <= 1e-7:
        equal += 1
    else:
        worse += 1
        worst = max(worst, d)
        print('seed', seed, 'n', n, 'h', h, 'device', res["fmin"], 'scipy', ref.fun, 'rel', d, 'status', res["status"], 'evals', res["nevals"])
print('device optimum better', better, 'equal', equal, 'worse', worse, 'worst relative excess', worst)
