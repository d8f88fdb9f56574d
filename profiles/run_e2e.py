"""End-to-end time of snaq_eval (host X -> host out) on config 2, against the device-only time."""
import json, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench, snaq_jl_b200 as S
S_, net, taxa, qt = bench.workload()
eng = S.Engine(0)
eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
eng.set_network(net.flatten(taxa))
x0, _, upper = eng.get_params()
X = S.simulate.parameter_batch(x0, upper, bench.NVEC, 1)
Xp = torch.from_numpy(X).pin_memory().numpy()
res = {}
for name, arr in (("pageable", X), ("pinned", Xp)):
    for _ in range(3):
        eng.eval(arr)
    t = time.perf_counter()
    for _ in range(30):
        eng.eval(arr)
    res[name + "_ms"] = (time.perf_counter() - t) / 30 * 1e3
print(json.dumps(res))
