"""CPU tier, world_size 2 over gloo: the quartet-sharded objective.  Each rank evaluates its contiguous block of the
quartet list (with the CPU build of the engine's device functions standing in for the GPU) and one all-reduce of P
doubles must reproduce the unsharded objective of the oracle."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

import snaq_jl_b200 as S
from snaq_jl_b200 import sharded, simulate


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    from tests import hostsim_py as hs

    n = 16
    net = simulate.random_level1_network(n, 2, 4242)
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = simulate.all_quartets(n)
    flat = net.flatten(taxa)
    obs = simulate.obs_from_expcf(oracle.extract(flat, qt)["expcf"], 100, 3)
    p = oracle.parameters(flat)
    X = simulate.parameter_batch(p["x"], p["upper"], 12, 5)
    lo, hi = sharded.shard_range(len(qt), rank, world)

    def local_eval(Xb):
        return hs.run(flat, n, qt[lo:hi], obs=obs[lo:hi], X=Xb)["lik"]

    obj = sharded.ShardedObjective(local_eval)
    total = obj.eval(X)
    want = oracle.evaluate(flat, qt, obs, X)
    err = float(np.max(np.abs(total - want) / np.abs(want)))
    q.put((rank, lo, hi, err, total.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_the_list():
    for nq in (0, 1, 15, 27405, 3921225):
        for world in (1, 2, 3, 8):
            r = [sharded.shard_range(nq, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == nq
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_quartet_sharded_objective_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == 0 and res[0][2] == res[1][1]            # contiguous, disjoint blocks
    for rank, lo, hi, err, total in res:
        assert err <= 1e-12, (rank, err)
    assert res[0][4] == res[1][4]                                # every rank ends with the same totals
