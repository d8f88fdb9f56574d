"""Randomized parity sweep of the evaluation kernels on the GPU against the oracle: random level-1 networks
(0-4 hybrids, cycles of 3 or more nodes), random batch sizes (every kernel: thread-per-quartet, lane-per-candidate,
persistent), sampling masks, plain and deduplicated tables, and the objective+gradient pass.
Usage: python profiles/run_sweep_eval.py [n_seeds]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import snaq_jl_b200 as S
from snaq_jl_b200 import simulate
from oracle import oracle

nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 60
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
worst = 0.0
done = 0
for seed in range(first, first + nseeds):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(6, 36))
    h = int(rng.integers(0, 5))
    try:
        net = simulate.random_level1_network(n, h, 7000 + seed, min_cycle=3 if seed % 2 else 4)
    except S.SnaqError:
        continue
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = simulate.all_quartets(n)
    if seed % 3 == 0:                                      # ragged, permuted table
        qt = qt[rng.uniform(size=len(qt)) < 0.6]
        qt = np.array([rng.permutation(q) for q in qt], dtype=np.uint16)
    flat = net.flatten(taxa)
    p = oracle.parameters(flat)
    o = oracle.extract(flat, qt)
    obs = simulate.obs_from_expcf(o["expcf"], int(rng.integers(5, 300)), seed)
    mask = (rng.uniform(size=len(qt)) < 0.7).astype(np.uint8) if seed % 4 == 1 else None
    P = int(rng.choice([1, 2, 3, 4, 5, 9, 16, 17, 40, 64, 130, 300]))
    X = simulate.parameter_batch(p["x"], p["upper"], P, seed)
    if seed % 5 == 0:                                      # corners of the box: caps, zero lengths, gamma 0 / 1
        X[-1] = np.where(rng.uniform(size=X.shape[1]) < 0.5, 0.0, p["upper"])
    want = oracle.evaluate(flat, qt, obs, X, sampled=mask)
    for dedup in (False, True):
        eng = S.Engine(0, dedup=dedup)
        eng.set_data(S.DataCF(taxa, qt, obs))
        if mask is not None:
            eng.set_sampled(mask)
        eng.set_network(flat)
        got = eng.eval(X)
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), fin), (seed, dedup, 'finite')
        scale = np.maximum(np.abs(want[fin]), 1.0 if dedup else 1e-300)
        rel = np.max(np.abs(got[fin] - want[fin]) / scale) if fin.any() else 0.0
        assert rel <= (1e-9 if dedup else 1e-10), (seed, dedup, n, h, P, rel)
        worst = max(worst, rel)
        k = int(rng.integers(0, P))
        if np.isfinite(want[k]) and want[k] < 1e14:
            f, g = eng.eval_grad(X[k])
            assert abs(f - want[k]) <= 1e-9 * max(abs(want[k]), 1.0), (seed, dedup, 'grad f', f, want[k])
            j = int(rng.integers(0, X.shape[1]))
            hstep = 1e-6 * max(1.0, abs(X[k, j]))
            xp, xm = X[k].copy(), X[k].copy()
            xp[j] = min(X[k, j] + hstep, p["upper"][j]); xm[j] = max(X[k, j] - hstep, 0.0)
            if xp[j] > xm[j] and 0.02 < X[k, j] < 0.98 * p["upper"][j]:
                fd = (oracle.evaluate(flat, qt, obs, xp[None], sampled=mask)[0] -
                      oracle.evaluate(flat, qt, obs, xm[None], sampled=mask)[0]) / (xp[j] - xm[j])
                # central difference of the ORACLE: rounding noise ~ |f| * 1e-15 / step on top of the comparison tolerance
                noise = 4e-15 * abs(want[k]) / (xp[j] - xm[j])
                assert abs(g[j] - fd) <= 1e-4 * max(abs(fd), 1.0) + 1e-3 + noise, (seed, dedup, 'grad', j, g[j], fd, want[k], noise)
    done += 1
print('networks', done, 'worst relative difference', worst)
