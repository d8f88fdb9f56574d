"""Quartet-sharded contexts with the collective inside the library (snaq_comm_init): correctness and strong scaling.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      profiles/run_comm.py [ntaxa nhyb]

Every rank holds the block [r*nq/N, (r+1)*nq/N) of the quartets; all ranks evaluate the same parameter vectors.  Checks:
totals identical on every rank (bitwise), equal to the unsharded engine on rank 0 (1e-12), for one objective call (fused
peer-memory epilogue), small batches, a batch of 64 (ncclAllReduce), the gradient pass and a whole optBL!.  Then times one
objective call and a batch of 64 with CUDA events (max over ranks)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import snaq_jl_b200 as S


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    h = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")                       # only to hand the 128-byte id around and to gather the checks
    net = S.simulate.random_level1_network(n, h, 20261022)
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = S.simulate.all_quartets(n)
    nq = len(qt)
    rng = np.random.default_rng(3)
    obs = rng.random((nq, 3)) + 0.05
    obs /= obs.sum(axis=1, keepdims=True)
    d = S.DataCF(taxa, qt, obs)
    flat = net.flatten(taxa)
    eng = S.Engine(local, rank=rank, world_size=world)
    eng.set_data(d)
    eng.set_network(flat)
    if world > 1 and os.environ.get("SNAQ_COMM_SKIP", "0") != "1":      # SNAQ_COMM_SKIP=1: partial sums only (what the collective costs)
        ids = [S.Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.comm_init(ids[0])
    info = eng.comm_info()
    x0, _, upper = eng.get_params()
    X = S.simulate.parameter_batch(x0, upper, 64, 5)
    out = {"ntaxa": n, "nq": nq, "world": world, "comm": info}
    v1 = eng.eval(X[:1]); v3 = eng.eval(X[:3]); v7 = eng.eval(X[:7]); v64 = eng.eval(X)
    nocomm = world > 1 and not info["has_comm"]
    if nocomm:
        f, g, xmin, res = 0.0, np.zeros(1), np.zeros(1), {"fmin": 0.0}
    else:
        f, g = eng.eval_grad(X[1])
        start = np.clip(x0 * np.random.default_rng(0).uniform(0.7, 1.3, len(x0)), 0, upper)
        xmin, res = eng.optbl(start)
    # identical on every rank
    mine = np.concatenate([v1, v3, v7, v64, [f], g, xmin, [res["fmin"]]])
    allv = [None] * world
    dist.all_gather_object(allv, mine.tobytes())
    out["identical_on_all_ranks"] = all(b == allv[0] for b in allv)
    np.testing.assert_allclose(v3, v64[:3], rtol=1e-12); np.testing.assert_allclose(v7, v64[:7], rtol=1e-12)
    assert nocomm or abs(f - v64[1]) <= 1e-11 * abs(f)
    # timing: device-resident X, CUDA events, max over ranks
    dev = torch.device("cuda", local)
    st = torch.cuda.Stream(dev)
    for P in (1, 64):
        dX = torch.from_numpy(X[:P].copy()).to(dev); dout = torch.zeros(P, dtype=torch.float64, device=dev)
        for _ in range(5):
            eng.eval_device(dX, dout, st)
        torch.cuda.synchronize(); dist.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        with torch.cuda.stream(st):
            for a, b in ev:
                a.record(st); eng.eval_device(dX, dout, st); b.record(st)
        torch.cuda.synchronize()
        ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[f"P{P}_ms"] = float(t.item())
        out[f"P{P}_evals_per_s"] = nq * P / (float(t.item()) * 1e-3)
        out[f"P{P}_value0"] = float(dout[0].item())
        # from the host (snaq_eval: one call = launch + wait)
        t0 = time.perf_counter()
        for _ in range(20):
            eng.eval(X[:P])
        out[f"P{P}_host_call_ms"] = (time.perf_counter() - t0) / 20 * 1e3
    out["optbl"] = res
    if rank == 0 and world > 1 and not nocomm and os.environ.get("SNAQ_COMM_CHECK_UNSHARDED", "1") == "1":
        ref = S.Engine(local)
        ref.set_data(d); ref.set_network(flat)
        r64 = ref.eval(X)
        out["max_rel_diff_vs_unsharded"] = float(np.max(np.abs(r64 - v64) / np.abs(r64)))
        fr, gr = ref.eval_grad(X[1])
        out["grad_max_rel_diff_vs_unsharded"] = float(np.max(np.abs(gr - g) / (1 + np.abs(gr))))
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
