import os, sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
S, net, taxa, qt = bench.workload()
for pinned in (True, False):
    for ch in ("1", "2", "4"):
        os.environ["SNAQ_EVAL_CHUNKS"] = ch
        eng = S.Engine(0)
        eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
        eng.set_network(net.flatten(taxa))
        x0, _, upper = eng.get_params()
        e0, _, _ = eng.eval_quartets(x0)
        eng.set_obscf(S.simulate.obs_from_expcf(e0, 300, 1))
        X = S.simulate.parameter_batch(x0, upper, 4096, 7)
        if pinned:
            X = torch.from_numpy(X).pin_memory().numpy()
        for _ in range(5): eng.eval(X)
        t0 = time.perf_counter()
        for _ in range(50): eng.eval(X)
        dt = (time.perf_counter() - t0) / 50
        print("pinned", pinned, "chunks", ch, "ms %.4f" % (dt * 1e3), "evals/s %.3g" % (len(qt) * 4096 / dt), flush=True)
