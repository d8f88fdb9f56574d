"""Randomized parity sweep on LARGE tables (45-70 taxa, 0-5 hybrids: the streaming gradient kernel, the ring of the
single-call kernel, the batched kernels) against the oracle, and of the gradient against the deduplicated engine (an
independent route).  Usage: python profiles/run_sweep_large.py [n_cases [first_seed]]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import snaq_jl_b200 as S
from snaq_jl_b200 import simulate
from oracle import oracle

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
done = 0
worst = worst_g = 0.0
for seed in range(first, first + ncases):
    rng = np.random.default_rng(50000 + seed)
    n, h = int(rng.integers(45, 71)), int(rng.integers(0, 6))
    net = None
    for attempt in range(20):
        try:
            net = simulate.random_level1_network(n, h, 60000 + seed + 1000 * attempt, min_cycle=3 if seed % 2 else 4)
            break
        except S.SnaqError:
            pass
    if net is None:
        continue
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = simulate.all_quartets(n)
    flat = net.flatten(taxa)
    p = oracle.parameters(flat)
    obs = simulate.obs_from_expcf(oracle.extract(flat, qt)["expcf"], int(rng.integers(10, 200)), seed)
    mask = (rng.uniform(size=len(qt)) < 0.75).astype(np.uint8) if seed % 3 == 0 else None
    P = int(rng.choice([1, 3, 8, 20]))
    X = simulate.parameter_batch(p["x"], p["upper"], P, seed)
    if seed % 4 == 0:
        X[-1] = np.where(rng.uniform(size=X.shape[1]) < 0.5, 0.0, p["upper"])     # a corner of the box
    want = oracle.evaluate(flat, qt, obs, X, sampled=mask)
    eng = S.Engine(0, dedup=False)
    ded = S.Engine(0, dedup=True)
    for e in (eng, ded):
        e.set_data(S.DataCF(taxa, qt, obs))
        if mask is not None:
            e.set_sampled(mask)
        e.set_network(flat)
    got = eng.eval(X)
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin), (seed, 'finite')
    rel = float(np.max(np.abs(got[fin] - want[fin]) / np.maximum(np.abs(want[fin]), 1e-300))) if fin.any() else 0.0
    assert rel <= 1e-10, (seed, n, h, P, rel)
    worst = max(worst, rel)
    k = 0                                                   # the network's own parameters: finite, no penalty
    x = np.clip(X[k], 1e-4, p["upper"] - 1e-4)
    f, g, hd = eng.eval_grad(x, want_hdiag=True)
    fd, gd, hdd = ded.eval_grad(x, want_hdiag=True)
    fo = oracle.evaluate(flat, qt, obs, x[None], sampled=mask)[0]
    assert abs(f - fo) <= 1e-10 * abs(fo) and abs(fd - fo) <= 1e-9 * max(abs(fo), 1.0), (seed, 'grad f', f, fd, fo)
    rg = float(np.max(np.abs(g - gd) / (1 + np.abs(gd))))
    rh = float(np.max(np.abs(hd - hdd) / (1 + np.abs(hdd))))
    # the plain table quantises every quartet's contribution to 2^-32 before the integer sums, the deduplicated one every
    # run's: with 10^5..10^6 quartets on a slot the two differ by a few 1e-7 (random walk of 1.2e-10 roundings)
    assert rg < 5e-6 and rh < 5e-6, (seed, n, h, 'gradient', rg, rh)
    worst_g = max(worst_g, rg, rh)
    done += 1
print('large tables', done, 'worst relative difference of the objective', worst, 'of gradient / curvature between the two routes', worst_g)
