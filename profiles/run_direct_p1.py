"""One objective call (P = 1) on the 100-taxon per-quartet table: the launch bench.py's `roofline_hbm` times (k_eval_direct).
For ncu:  ncu -k regex:k_eval_direct -s 5 -c 1 ... python profiles/run_direct_p1.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import snaq_jl_b200 as S
import bench

n, net, taxa, qt = bench.table100(S)
eng = S.Engine(0, dedup=False)
eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
eng.set_network(net.flatten(taxa))
x0, _, _ = eng.get_params()
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(dev)
dX = torch.from_numpy(x0.reshape(1, -1)).to(dev)
dout = torch.zeros(1, dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ms = bench.timed_launches(torch, eng, dX, dout, st, flush, 10, warm=5)
print("ms per call", float(ms.mean()), "value", float(dout.item()), "bytes", eng.counters()["single_call_bytes"])
