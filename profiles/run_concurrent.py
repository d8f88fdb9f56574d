"""Candidate topologies optimised concurrently on ONE GPU: T host threads, one context each
(own stream, own table), every thread runs set_network + snaq_optbl on its own candidate network over the same
data.  ctypes releases the GIL during the calls.  Prints optBL!/s for T = 1, 2, 4, 8, 16."""
import json
import sys
import threading
import time

sys.path.insert(0, '/root/repo')
import numpy as np

import snaq_jl_b200 as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
h = int(sys.argv[2]) if len(sys.argv) > 2 else 2
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
taxa = [f"t{i + 1}" for i in range(n)]
qt = S.simulate.all_quartets(n)
truth = S.simulate.random_level1_network(n, h, 20261021)
e0 = S.Engine(0)
e0.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
e0.set_network(truth.flatten(taxa))
x0, _, _ = e0.get_params()
obs = S.simulate.obs_from_expcf(e0.eval_quartets(x0)[0], 300, 5)
cands = []
seed = 0
while len(cands) < 16:
    seed += 1
    try:
        cands.append(S.simulate.random_level1_network(n, h, 5000 + seed).flatten(taxa))
    except S.SnaqError:
        pass
res = {"ntaxa": n, "nq": len(qt), "h": h, "reps": reps, "threads": {}}
TMAX = 16
engines = []
for t in range(TMAX):
    e = S.Engine(0)
    e.set_data(S.DataCF(taxa, qt, obs))
    engines.append(e)


def work(t, out):
    try:
        _work(t, out)
    except Exception as ex:                                # surfaces in the main thread
        out[t] = ex


def _work(t, out):
    eng = engines[t]
    fs = []
    for r in range(reps):
        eng.set_network(cands[(t + r) % len(cands)])
        x, _, up = eng.get_params()
        try:
            xm, info = eng.optbl(x)
        except Exception as ex:
            raise RuntimeError(f"thread {t} rep {r} nparams {len(x)}: {ex}")
        fs.append(info["fmin"])
    out[t] = fs


ref = None
for T in (1, 1, 2, 4, 8, 16):
    out = {}
    th = [threading.Thread(target=work, args=(t, out)) for t in range(T)]
    t0 = time.perf_counter()
    for x in th:
        x.start()
    for x in th:
        x.join()
    dt = time.perf_counter() - t0
    for v in out.values():
        if isinstance(v, Exception):
            raise v
    if ref is None:
        ref = out[0]
    assert out[0] == ref                                  # same candidates, same result whatever runs beside them
    res["threads"][T] = {"optbl_per_s": T * reps / dt, "ms_per_optbl_per_thread": dt / reps * 1e3}
# the same through the C ABI's own worker pool (snaq_optbl_topologies)
batch = [cands[(r) % len(cands)] for r in range(4 * len(cands))]
main = engines[0]
res["native"] = {}
for W in (1, 2, 4, 8, 16):
    main.optbl_topologies(batch[:W], workers=W)            # creates / syncs the workers
    t0 = time.perf_counter()
    out = main.optbl_topologies(batch, workers=W)
    dt = time.perf_counter() - t0
    res["native"][W] = {"optbl_per_s": len(batch) / dt}
res["per_optbl"] = {k: float(np.mean([o[1][k] for o in out])) for k in ("nevals", "nlaunches", "niter")}
t0 = time.perf_counter()
for c in batch:
    main.set_network(c)
res["set_network_ms"] = (time.perf_counter() - t0) / len(batch) * 1e3
print(json.dumps(res))
