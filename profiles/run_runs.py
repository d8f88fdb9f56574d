"""Times the batched path (device-resident X) on BASELINE configs[1] (30 taxa, 4096 vectors), on the north-star table
(100 taxa, 3.92 M quartets, 256 vectors) and two larger ones: the default run table (runs = -1), and on the per-quartet
table (SNAQ_OPT_PER_QUARTET) the run-length kernel k_eval_runs (runs = 2: forced, 1: when runs are long) and k_eval_lanes
(runs = 0).  One leg with few repetitions for ncu:  python profiles/run_runs.py n h P seed runs"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import snaq_jl_b200 as S

def leg(n, h, P, seed, runs, reps=20):
    os.environ["SNAQ_RUNS"] = str(max(int(runs), 0))
    net = S.simulate.random_level1_network(n, h, seed)
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = S.simulate.all_quartets(n)
    eng = S.Engine(0, dedup=None if int(runs) < 0 else False)
    eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
    eng.set_network(net.flatten(taxa))
    t0 = time.perf_counter()
    for _ in range(5):
        eng.set_network(net.flatten(taxa))
    build_ms = (time.perf_counter() - t0) / 5 * 1e3
    x0, _, upper = eng.get_params()
    e0, _, _ = eng.eval_quartets(x0)
    eng.set_obscf(S.simulate.obs_from_expcf(e0, 300, 1))
    X = S.simulate.parameter_batch(x0, upper, P, 7)
    dev = torch.device("cuda", 0)
    dX = torch.from_numpy(X).to(dev); dout = torch.zeros(P, dtype=torch.float64, device=dev)
    st = torch.cuda.Stream(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        eng.eval_device(dX, dout, st)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    with torch.cuda.stream(st):
        for a, b in ev:
            flush.fill_(1); a.record(st); eng.eval_device(dX, dout, st); b.record(st)
    torch.cuda.synchronize()
    eng.check_status()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    return dict(n=n, P=P, runs=runs, ms=ms, evals_per_s=len(qt) * P / ms * 1e3, build_ms=build_ms, out0=float(dout[0].item()),
                out_last=float(dout[-1].item()), programs=eng.counters()["distinct_programs"])

if __name__ == "__main__":
    res = []
    if len(sys.argv) > 1:          # one leg, few repetitions (for ncu): n h P seed runs
        a = sys.argv[1:]
        print(json.dumps(leg(int(a[0]), int(a[1]), int(a[2]), int(a[3]), int(a[4]), reps=3)))
        sys.exit(0)
    for n, h, P, seed in ((30, 2, 4096, 20261020), (100, 3, 256, 20261021), (100, 3, 4096, 20261021), (200, 5, 64, 20261022)):
        for runs in (-1, 1, 0):
            r = leg(n, h, P, seed, runs)
            print(json.dumps(r), flush=True)
            res.append(r)
