"""Times one objective call (P=1..8) on the 100-taxon table: cold (L2 flushed) and warm.  Used for quick
experiments and as the ncu target for k_eval_direct (profiles/*.md)."""
import json
import sys

sys.path.insert(0, '/root/repo')
import numpy as np
import torch

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
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
res = {"nq": len(qt), "bytes": eng.counters()["single_call_bytes"]}
for P in (1, 2, 4, 8):
    X = torch.from_numpy(S.simulate.parameter_batch(x0, upper, P, 3)).to(dev)
    out = torch.zeros(P, dtype=torch.float64, device=dev)
    with torch.cuda.stream(stream):
        for _ in range(5):
            eng.eval_device(X, out, stream)
    torch.cuda.synchronize()
    for mode in ("cold", "warm"):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        with torch.cuda.stream(stream):
            for a, b in ev:
                if mode == "cold":
                    flush.fill_(1)
                a.record(stream)
                eng.eval_device(X, out, stream)
                b.record(stream)
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)
        res[f"P{P}_{mode}_us"] = round(1e3 * float(np.mean(ms)), 2)
        res[f"P{P}_{mode}_min_us"] = round(1e3 * ms[0], 2)
eng.check_status()
print(json.dumps(res))
