"""optBL! on the device: time of one objective+gradient pass (snaq_eval_grad) and of a whole snaq_optbl call
with the search-time and the final tolerances, on a synthetic n-taxon level-1 network with h hybrids."""
import json
import sys
import time

sys.path.insert(0, '/root/repo')
import numpy as np

import snaq_jl_b200 as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
h = int(sys.argv[2]) if len(sys.argv) > 2 else 3
net = S.simulate.random_level1_network(n, h, 20261021)
taxa = [f"t{i + 1}" for i in range(n)]
qt = S.simulate.all_quartets(n)
eng = S.Engine(0)
eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
eng.set_network(net.flatten(taxa))
x0, _, upper = eng.get_params()
e, _, _ = eng.eval_quartets(x0)
eng.set_obscf(S.simulate.obs_from_expcf(e, 300, 5))
rng = np.random.default_rng(0)
start = np.clip(x0 * rng.uniform(0.7, 1.3, len(x0)), 0, upper)
res = {"ntaxa": n, "nq": len(qt), "nparams": len(x0)}
for _ in range(3):
    eng.eval_grad(start)
t = time.perf_counter()
for _ in range(20):
    eng.eval_grad(start)
res["eval_grad_us"] = (time.perf_counter() - t) / 20 * 1e6
t = time.perf_counter()
for _ in range(20):
    eng.eval(start)
res["eval_P1_host_us"] = (time.perf_counter() - t) / 20 * 1e6
for name, tol in (("search", (1e-6, 1e-6, 1e-2, 1e-3)), ("final", (1e-12, 1e-10, 1e-10, 1e-10))):
    t = time.perf_counter()
    xmin, r = eng.optbl(start, *tol)
    r["seconds"] = time.perf_counter() - t
    r["f_start"] = float(eng.eval(start)[0])
    r["f_true_x"] = float(eng.eval(x0)[0])
    res[name] = r
t = time.perf_counter()
xmin, r = eng.optbl(start, 1e-6, 1e-6, 1e-2, 1e-3, fd=True)
r["seconds"] = time.perf_counter() - t
res["search_fd"] = r
print(json.dumps(res))
