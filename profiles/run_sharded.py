"""BASELINE.json configs[4]: synthetic 200-taxon h=5 network, all 64,684,950 quartets, quartet-sharded over the
ranks of one node (`python -m torch.distributed.run --nproc-per-node N profiles/run_sharded.py`; N=1 works too).
Every rank holds the contiguous block [r*nq/N, (r+1)*nq/N) of the descriptor table and evaluates the same
parameter vectors; the per-candidate partial sums are combined by ONE NCCL all-reduce of P doubles.
Timed on the device (CUDA events on the launching stream), max over ranks."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import snaq_jl_b200 as S

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
h = int(sys.argv[2]) if len(sys.argv) > 2 else 5
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
net = S.simulate.random_level1_network(n, h, 20261023)
taxa = [f"t{i + 1}" for i in range(n)]
qt = S.simulate.all_quartets(n)
nq = len(qt)
lo, hi = S.sharded.shard_range(nq, rank, world)
eng = S.Engine(local, rank=rank, world_size=world)
t0 = time.perf_counter()
eng.set_data(S.DataCF(taxa, qt, np.full((nq, 3), 1 / 3)))
t_data = time.perf_counter() - t0
t0 = time.perf_counter()
eng.set_network(net.flatten(taxa))
t_build = time.perf_counter() - t0
x0, _, upper = eng.get_params()
stream = torch.cuda.Stream(dev)
res = {"workload": f"{n}-taxon h={h}, {nq} quartets, quartet-sharded over {world} GPU(s)", "n_gpus": world, "nparams": len(x0),
       "set_data_s": t_data, "descriptor_build_s": t_build, "quartets_per_rank": hi - lo}
ref = None
for P in (1, 64):
    X = torch.from_numpy(S.simulate.parameter_batch(x0, upper, P, 3)).to(dev)
    out = torch.zeros(P, dtype=torch.float64, device=dev)

    def step():
        eng.eval_device(X, out, stream)
        if world > 1:
            dist.all_reduce(out, op=dist.ReduceOp.SUM)          # on torch's current stream = `stream` below

    with torch.cuda.stream(stream):
        for _ in range(3):
            step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    with torch.cuda.stream(stream):
        for a, b in ev:
            a.record(stream)
            step()
            b.record(stream)
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    eng.check_status()
    res[f"P{P}_ms"] = float(t.item())
    res[f"P{P}_evals_per_s"] = nq * P / (float(t.item()) * 1e-3)
    res[f"P{P}_value0"] = float(out[0].item())
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
