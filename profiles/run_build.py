"""Steady-state cost of snaq_set_network (extractQuartet! for every quartet) and of one candidate topology
(descriptor rebuild + optBL! with the search-time tolerances)."""
import json, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import snaq_jl_b200 as S
res = {}
for n, h in ((30, 2), (100, 3)):
    nets = [S.simulate.random_level1_network(n, h, 20261021 + k) for k in range(4)]
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = S.simulate.all_quartets(n)
    eng = S.Engine(0)
    eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
    flats = [nt.flatten(taxa) for nt in nets]
    eng.set_network(flats[0])
    x0, _, _ = eng.get_params()
    e, _, _ = eng.eval_quartets(x0)
    eng.set_obscf(S.simulate.obs_from_expcf(e, 300, 5))
    for f in flats:
        eng.set_network(f)
    t = time.perf_counter()
    for r in range(5):
        for f in flats:
            eng.set_network(f)
    res[f"set_network_ms_{n}taxa"] = (time.perf_counter() - t) / 20 * 1e3
    t = time.perf_counter()
    for f in flats:
        eng.set_network(f)
        x, _, _ = eng.get_params()
        eng.optbl(x)
    res[f"candidate_topology_ms_{n}taxa"] = (time.perf_counter() - t) / len(flats) * 1e3
print(json.dumps(res))
# detail of the last size: iterations / launches per candidate
for f in flats:
    eng.set_network(f)
    x, _, _ = eng.get_params()
    t = time.perf_counter()
    xm, r = eng.optbl(x)
    r["ms"] = (time.perf_counter() - t) * 1e3
    r["f_start"] = float(eng.eval(x)[0])
    print(json.dumps(r))
