#!/usr/bin/env python
"""bench.py — quartet pseudolikelihood evaluations per second (FP64) and the search-shaped workloads around them.

One "step" = one pass of the hot path over one batch: every quartet of the table evaluated under every parameter
vector of an optBL!-style batch (BASELINE.json configs[1]: synthetic 30-taxon level-1 network, h=2, all 27,405
quartets x 4096 parameter vectors = 112.3 M evals per step).

  python bench.py --gpus N --steps K --warmup W            # the engine
  python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port)

The engine's default evaluates every distinct quartet program once per parameter vector with the summed weights of its
quartets (the run table, include/snaq_b200.h); every (quartet, vector) pair is in the result, so `value` counts them all.
The per-quartet kernels (SNAQ_OPT_PER_QUARTET: one weighted contribution computed per pair) are timed beside it.

Legs of the JSON line (rank 0 prints it):
  value / e2e / roofline      configs[1], one independent batch per GPU (weak scaling, no data-path collective)
  north_star                  100 taxa, 3,921,225 quartets, 256 vectors per launch (the north star's config): device-timed,
                              end to end from host buffers, roofline on EXECUTED double-precision flops
  per_quartet                 both of the above on the per-quartet kernels (k_eval_lanes / k_eval_runs), with their rooflines
  roofline_hbm                the 100-taxon table, one objective call (P = 1): HBM roofline of the per-quartet streaming
                              kernel, the default beside it; optBL! timings
  sharded                     configs[4]: 200 taxa, 64,684,950 quartets split over the N GPUs, collective inside the
                              library (fused peer-memory epilogue for P = 1, ncclAllReduce for P = 64): strong scaling
  search_chains               configs[2]'s shape: 8 hill-climbing chains of 50 candidate topologies (NNI proposals,
                              descriptor patch + optBL! at the search tolerances) on the 100-taxon table, chains dealt to GPUs
  bootstrap                   configs[3]'s shape: 96 replicates (CI resampling + updateBL! table + optBL!) at 50 taxa, dealt to GPUs
  candidates                  concurrent candidate optimisation on one GPU (worker contexts)
  cpu_baseline                the oracle port on the host cores (rank 0, N = 1 only)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

NTAXA, NHYB, NVEC = 30, 2, 4096
NET_SEED, OBS_SEED, X_SEED = 20261020, 1, 7
# SURVEY.md 8d's convention, kept as a secondary figure only (it counts work the kernels do not execute, see roofline)
FLOP_TREE, FLOP_T5_EXTRA, FLOP_PER_TERM = 110.0, 75.0, 90.0
EXECUTED = os.path.join(ROOT, "profiles", "r02_executed_flops.json")


def workload(ntaxa=NTAXA, nhyb=NHYB, nvec=NVEC):
    import snaq_jl_b200 as S
    net = S.simulate.random_level1_network(ntaxa, nhyb, NET_SEED)
    taxa = [f"t{i + 1}" for i in range(ntaxa)]
    qt = S.simulate.all_quartets(ntaxa)
    return S, net, taxa, qt


def executed(key):
    """Per-launch counts of the timed kernel from the committed ncu capture of this very command (profiles/): thread-level
    DFMA / DADD / DMUL instructions and DRAM bytes.  They are deterministic for a given table and batch."""
    try:
        return json.load(open(EXECUTED)).get(key)
    except Exception:  # noqa: BLE001
        return None


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback of B200_PROFILING.md"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu = gpu
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port(S, net, taxa, qt, obs, X, threads):
    """Times the oracle port (the reference's algorithm restated in C++, OpenMP over quartets) on a
    bounded sample of the same workload: all quartets x the first len(X) parameter vectors."""
    from oracle import oracle
    flat = net.flatten(taxa)
    t0 = time.perf_counter()
    oracle.evaluate(flat, qt, obs, X, threads=threads)
    dt = time.perf_counter() - t0
    return len(qt) * len(X) / dt, dt


def run_reference(args):
    """Reference arm: the reference is pure Julia (+PhyloNetworks.jl, NLopt) and neither Julia nor
    those packages exist in the image, so oracle/_ref cannot be built; this arm times the C++ port
    of the reference's algorithm (oracle/) with all host threads, on the same config and metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    S, net, taxa, qt = workload()
    from oracle import oracle
    flat = net.flatten(taxa)
    cores = os.cpu_count() or 1
    exp0 = oracle.extract(flat, qt)["expcf"]
    obs = S.simulate.obs_from_expcf(exp0, 300, OBS_SEED)
    p = oracle.parameters(flat)
    Xall = S.simulate.parameter_batch(p["x"], p["upper"], NVEC, X_SEED)
    per_step = args.ref_vectors          # bounded sample: this many of the 4096 vectors per step
    for w in range(args.warmup):
        oracle.evaluate(flat, qt, obs, Xall[:max(8, per_step // 8)], threads=cores)
    t0 = time.perf_counter()
    for k in range(args.steps):
        lo = (k * per_step) % NVEC
        oracle.evaluate(flat, qt, obs, Xall[lo:lo + per_step], threads=cores)
    dt = time.perf_counter() - t0
    evals = len(qt) * per_step * args.steps
    val = evals / dt
    sample = f"{len(qt)} quartets x {per_step} of the {NVEC} parameter vectors per step"
    line = {"impl": "reference", "metric": "quartet pseudolik evals/s (FP64)", "value": val, "unit": "evals/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"synthetic {NTAXA}-taxon level-1 network h={NHYB}, {len(qt)} quartets x {NVEC} parameter vectors",
                       "sample": sample},
            "cpu_baseline": {"value": val, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def timed_launches(torch, eng, dX, dout, stream, flush, reps, warm=3):
    """mean ms per eng.eval_device launch: CUDA events on the launching stream, L2 flushed before every timed launch"""
    with torch.cuda.stream(stream):
        for _ in range(warm):
            eng.eval_device(dX, dout, stream)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    with torch.cuda.stream(stream):
        for a, b in ev:
            if flush is not None:
                flush.fill_(1)
            a.record(stream)
            eng.eval_device(dX, dout, stream)
            b.record(stream)
    torch.cuda.synchronize()
    return np.array([a.elapsed_time(b) for a, b in ev])


def fp64_roofline(key, kernel_s, peak_tf, evals, convention_flop_per_eval):
    """roofline.frac = EXECUTED double-precision flops / time / measured DFMA peak.  Executed flops per launch come from
    the committed ncu capture (2 x DFMA + DADD + DMUL thread instructions of the dominant kernel); the SURVEY 8d
    convention (one exp/log set per quartet and vector) is reported beside it but is NOT the fraction: the kernels
    evaluate a program once per run of identical programs, so the convention counts work that is not executed."""
    ex = executed(key)
    out = {"bound": "fp64", "peak": peak_tf, "unit": "TFLOP/s",
           "peak_source": "DFMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
           "convention_flop_per_eval": convention_flop_per_eval,
           "convention_tflops": evals * convention_flop_per_eval / kernel_s / 1e12}
    if ex:
        flops = 2.0 * ex["dfma"] + ex["dadd"] + ex["dmul"]
        out.update({"achieved": flops / kernel_s / 1e12, "frac": flops / kernel_s / 1e12 / peak_tf, "traffic": ex.get("dram_bytes"),
                    "executed_flop_per_launch": flops, "executed_flop_per_eval": flops / evals, "kernel": ex.get("kernel"),
                    "fp64_pipe_active_pct_ncu": ex.get("fp64_pipe_pct"), "source": ex.get("source")})
    else:
        out.update({"achieved": None, "frac": None, "traffic": None, "note": "profiles/r02_executed_flops.json missing"})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-vectors", type=int, default=128, help="parameter vectors per step of the CPU arms")
    ap.add_argument("--cpu-vectors", type=int, default=4096, help="parameter vectors of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hbm", action="store_true", help="skip the secondary single-GPU legs (north star, P=1, dedup, candidates)")
    ap.add_argument("--no-candidates", action="store_true", help="skip the concurrent-candidates leg (thousands of small launches)")
    ap.add_argument("--no-sharded", action="store_true", help="skip the 200-taxon quartet-sharded leg")
    ap.add_argument("--no-search", action="store_true", help="skip the search-chain and bootstrap legs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    S, net, taxa, qt = workload()
    eng = S.Engine(local)
    flat = net.flatten(taxa)
    d0 = S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3))
    eng.set_data(d0)
    t_build0 = time.perf_counter()
    eng.set_network(flat)
    t_build = time.perf_counter() - t_build0
    x0, numht, upper = eng.get_params()
    exp0, _, _ = eng.eval_quartets(x0)
    obs = S.simulate.obs_from_expcf(exp0, 300, OBS_SEED + rank)      # each rank: its own replicate of the data
    eng.set_obscf(obs)
    Xh = S.simulate.parameter_batch(x0, upper, NVEC, X_SEED + rank)
    nq, P, npar = len(qt), NVEC, len(x0)
    evals_per_step = nq * P

    dev = torch.device("cuda", local)
    dX = torch.from_numpy(Xh).to(dev)
    dout = torch.zeros(P, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.Stream(dev)      # a real (non-default) stream: the kernels and the timing events share it
    torch.cuda.synchronize()

    def step():
        eng.eval_device(dX, dout, stream)

    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            flush.fill_(1)
            step()
    torch.cuda.synchronize()
    eng.check_status()
    c0 = eng.counters()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with torch.cuda.stream(stream):
        for a, b in ev:
            flush.fill_(1)           # L2 flush between timed iterations (inputs are smaller than L2)
            a.record(stream)
            step()
            b.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = np.array([a.elapsed_time(b) for a, b in ev])
    elapsed = float(ms.sum()) * 1e-3
    c1 = eng.counters()
    # nvidia-smi samples every 100 ms: when the timed region is shorter than ~1.5 s keep the same kernel running
    # (untimed) so that the clocks are sampled under the same load
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < max(0.0, 1.5 - elapsed):
        with torch.cuda.stream(stream):
            for _ in range(50):
                step()
        torch.cuda.synchronize()
    clocks = sampler.stop()
    clocks["window"] = "timed steps + the same kernel repeated (untimed) up to 1.5 s"
    eng.check_status()
    if world > 1:
        t = torch.tensor([elapsed], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    value = world * evals_per_step * args.steps / elapsed
    launches = (c1["eval_launches"] - c0["eval_launches"]) * 3      # x expansion + batched kernel + fixed-order reduction per step

    # ---- end to end through the public API: host X (pinned staging inside the C ABI) -> out on host
    pinX = torch.from_numpy(Xh).pin_memory()
    Xp = pinX.numpy()
    for _ in range(2):
        eng.eval(Xp)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = eng.eval(Xp)
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e_val = world * evals_per_step * args.steps / e2e_dt
    assert np.allclose(res, dout.cpu().numpy(), rtol=1e-12)

    line = None
    peak_tf = None
    if rank == 0:
        desc = eng.descriptors()
        f5 = float(np.mean(desc["which"] == 2))
        cbar = float(np.mean(np.sum((desc["types"] >= 2) & (desc["types"] <= 4), axis=1)))
        flop_per_eval = FLOP_TREE + FLOP_T5_EXTRA * f5 + FLOP_PER_TERM * cbar
        kernel_s = float(ms.mean()) * 1e-3           # per launch (batched kernel + x expansion + its reduction pass)
        peak_tf = eng.measure_fp64_peak()
        roofline = fp64_roofline("config2", kernel_s, peak_tf, evals_per_step, flop_per_eval)
        line = {"metric": "quartet pseudolik evals/s (FP64)", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": float(ms.mean()), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"synthetic {NTAXA}-taxon level-1 network h={NHYB}, {nq} quartets x {P} parameter vectors "
                                       f"per step (BASELINE.json configs[1]); one independent batch per GPU",
                           "nparams": npar, "l2": "flushed between timed iterations (inputs < L2)", "descriptor_build_s": t_build},
                "e2e": {"value": e2e_val, "unit": "evals/s", "h2d_bytes_per_step": int(P * npar * 8), "d2h_bytes_per_step": int(P * 8 + 4)},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
    single = world == 1 and not args.no_hbm
    if rank == 0 and single:
        line["config"]["distinct_programs"] = int(eng.counters()["distinct_programs"])
        for name, fn in (("north_star", lambda: north_star_leg(S, eng.__class__, local, torch, peak_tf)),
                         ("per_quartet", lambda: per_quartet_leg(S, eng.__class__, local, torch, taxa, qt, obs, flat, dX, dout, stream, flush,
                                                                 evals_per_step, peak_tf, flop_per_eval)),
                         ("roofline_hbm", lambda: hbm_leg(S, eng.__class__, local, torch)),
                         ("candidates", lambda: candidates_leg(S, eng, taxa, qt, obs))):
            if name == "candidates" and args.no_candidates:
                continue
            try:
                line[name] = fn()
            except Exception as exc:  # noqa: BLE001
                line[name] = {"error": str(exc)[:300]}
    del eng, dX, dout
    torch.cuda.empty_cache()
    # ---- legs that use all N GPUs (every rank takes part) ---------------------------------------------------------
    multi = {}
    if not args.no_search:
        for name, fn in (("search_chains", search_chains_leg), ("bootstrap", bootstrap_leg)):
            try:
                multi[name] = fn(S, local, rank, world, dist, torch)
            except Exception as exc:  # noqa: BLE001
                multi[name] = {"error": str(exc)[:300]}
    if not args.no_sharded:
        try:
            multi["sharded"] = sharded_leg(S, local, rank, world, dist, torch, flush)
        except Exception as exc:  # noqa: BLE001
            multi["sharded"] = {"error": str(exc)[:300]}
    if rank == 0:
        line.update(multi)
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, dt = cpu_port(S, net, taxa, qt, obs, Xh[: args.cpu_vectors], cores)
            line["cpu_baseline"] = {"value": v, "unit": "evals/s", "cores": cores, "kind": "port", "seconds": dt,
                                    "sample": f"{nq} quartets x first {args.cpu_vectors} of the {P} parameter vectors, oracle port "
                                              f"(C++ restatement of the reference, OpenMP over quartets)"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def table100(S):
    n = 100
    net = S.simulate.random_level1_network(n, 3, 20261021)
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = S.simulate.all_quartets(n)
    return n, net, taxa, qt


def north_star_leg(S, Engine, local, torch, peak_tf, dedup=None, key="north_star"):
    """BASELINE.json north_star: batched logPseudoLik at 100 taxa (3,921,225 quartets), 256 parameter vectors per launch."""
    n, net, taxa, qt = table100(S)
    eng = Engine(local, dedup=dedup)
    eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
    flat = net.flatten(taxa)
    eng.set_network(flat)
    x0, _, upper = eng.get_params()
    e0, _, _ = eng.eval_quartets(x0)
    eng.set_obscf(S.simulate.obs_from_expcf(e0, 300, 5))
    P = 256
    X = S.simulate.parameter_batch(x0, upper, P, 11)
    dev = torch.device("cuda", local)
    dX = torch.from_numpy(X).to(dev)
    dout = torch.zeros(P, dtype=torch.float64, device=dev)
    stream = torch.cuda.Stream(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    ms = timed_launches(torch, eng, dX, dout, stream, flush, 20)
    eng.check_status()
    evals = len(qt) * P
    desc = eng.descriptors()
    f5 = float(np.mean(desc["which"] == 2))
    cbar = float(np.mean(np.sum((desc["types"] >= 2) & (desc["types"] <= 4), axis=1)))
    out = {"workload": f"100-taxon h=3, {len(qt)} quartets x {P} parameter vectors per launch, L2 flushed before every timed launch",
           "ms_per_launch": float(ms.mean()), "value": evals / (float(ms.mean()) * 1e-3), "unit": "evals/s",
           "distinct_programs": int(eng.counters()["distinct_programs"])}
    out["roofline"] = fp64_roofline(key, float(ms.mean()) * 1e-3, peak_tf, evals, FLOP_TREE + FLOP_T5_EXTRA * f5 + FLOP_PER_TERM * cbar)
    # end to end: host X (pageable, as a Julia caller's) -> out on the host, through snaq_eval
    for _ in range(2):
        r = eng.eval(X)
    t0 = time.perf_counter()
    for _ in range(10):
        r = eng.eval(X)
    dt = (time.perf_counter() - t0) / 10
    assert np.allclose(r, dout.cpu().numpy(), rtol=1e-12)
    out["e2e"] = {"value": evals / dt, "unit": "evals/s", "ms_per_call": dt * 1e3, "h2d_bytes_per_step": int(X.nbytes), "d2h_bytes_per_step": int(P * 8 + 4)}
    return out


def per_quartet_leg(S, Engine, local, torch, taxa, qt, obs, flat, dX, dout, stream, flush, evals_per_step, peak_tf, flop_per_eval):
    """The main step and the north-star launch on engines created with dedup=False (SNAQ_OPT_PER_QUARTET): one weighted
    contribution computed per (quartet, vector) pair -- k_eval_lanes on configs[1] (short runs), k_eval_runs at 100 taxa."""
    eng = Engine(local, dedup=False)
    eng.set_data(S.DataCF(taxa, qt, obs))
    eng.set_network(flat)
    ref = dout.clone()
    ms = timed_launches(torch, eng, dX, dout, stream, flush, 20)
    eng.check_status()
    out = {"note": "SNAQ_OPT_PER_QUARTET: every (quartet, vector) contribution computed, programs of neighbouring quartets memoised",
           "max_rel_diff_vs_default": float(((dout - ref).abs() / ref.abs()).max().item()),
           "config2": {"ms_per_step": float(ms.mean()), "value": evals_per_step / (float(ms.mean()) * 1e-3), "unit": "evals/s",
                       "roofline": fp64_roofline("config2_per_quartet", float(ms.mean()) * 1e-3, peak_tf, evals_per_step, flop_per_eval)}}
    del eng
    out["north_star"] = north_star_leg(S, Engine, local, torch, peak_tf, dedup=False, key="north_star_per_quartet")
    return out


def candidates_leg(S, eng, taxa, qt, obs, ncand=32):
    """optBL! of `ncand` random candidate topologies (0-2 hybrids) on the bench's own 30-taxon data: one context after the
    other (workers = 1) against the worker pool of snaq_optbl_topologies.  The results are identical, only the
    throughput changes."""
    n = len(taxa)
    flats, seed = [], 0
    while len(flats) < ncand:
        seed += 1
        try:
            flats.append(S.simulate.random_level1_network(n, seed % 3, 90000 + seed).flatten(taxa))
        except S.SnaqError:
            pass
    out = {"workload": f"{ncand} random candidate topologies (0-2 hybrids) on the {n}-taxon table, search-time tolerances",
           "host_cores": os.cpu_count()}
    ref = None
    for w in (1, 8, 16):
        eng.optbl_topologies(flats[:w], workers=w)                 # untimed: creates the worker contexts
        t0 = time.perf_counter()
        res = eng.optbl_topologies(flats, workers=w)
        dt = time.perf_counter() - t0
        f = [r[1]["fmin"] for r in res]
        ref = ref or f
        out[f"workers_{w}"] = {"optbl_per_s": ncand / dt, "identical_to_workers_1": f == ref}
    out["launches_per_optbl"] = float(np.mean([r[1]["nlaunches"] for r in res]))
    return out


def hbm_leg(S, Engine, local, torch):
    """One optBL! objective call (P = 1) on 100 taxa / h=3 / 3,921,225 quartets, two ways.
    (a) the per-quartet streaming kernel (k_eval_direct over the full table, SNAQ_OPT_PER_QUARTET; also the kernel behind the
        per-quartet read-backs): HBM roofline, bytes per launch = what the layout makes it read once (engine counter: 16 B
        weights + 8 B per program chunk for additive quartets, 37 B + programs for the rest), peak = MEASURED_PEAKS.json
        hbm_gbs, timed cold (L2 flushed before every call) and warm;
    (b) the default: the same call on the run table with run-summed weights (one entry per distinct program; the sums are
        formed once per set_obscf), which reads ~100x less and is launch-latency bound; and optBL! on top of it."""
    n, net, taxa, qt = table100(S)
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    flat = net.flatten(taxa)
    res = {}
    for name, mode in (("per_quartet_kernel", False), ("default", None)):
        eng = Engine(local, dedup=mode)
        eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
        t0 = time.perf_counter()
        eng.set_network(flat)
        t_build = time.perf_counter() - t0
        x0, _, upper = eng.get_params()
        dX = torch.from_numpy(x0.reshape(1, -1)).to(dev)
        dout = torch.zeros(1, dtype=torch.float64, device=dev)
        ms = float(timed_launches(torch, eng, dX, dout, stream, flush, 20, warm=5).mean())
        ms_warm = float(timed_launches(torch, eng, dX, dout, stream, None, 50, warm=5).mean())
        for _ in range(5):
            eng.eval(x0)
        t0 = time.perf_counter()
        for _ in range(50):
            eng.eval(x0)
        host_us = (time.perf_counter() - t0) / 50 * 1e6
        res[name] = {"ms_per_call": ms, "ms_per_call_warm_l2": ms_warm, "host_call_us": host_us, "evals_per_s": len(qt) / (ms * 1e-3),
                     "value": float(dout.item()), "descriptor_build_s": t_build}
        if name == "per_quartet_kernel":
            bytes_alg = eng.counters()["single_call_bytes"]
            res[name]["bytes_per_quartet"] = bytes_alg / len(qt)
        else:
            res[name]["distinct_programs"] = eng.counters()["distinct_programs"]
    peak, src = hbm_peak()
    ach = bytes_alg / (res["per_quartet_kernel"]["ms_per_call"] * 1e-3) / 1e9
    ex = executed("hbm_p1") or {}
    # optBL! on the same table (src/snaq_optimization.jl:341-390): whole optimisations through snaq_optbl, from a
    # start 30 % off the simulating parameters, with the search-time and the final tolerances of the reference
    exp0, _, _ = eng.eval_quartets(x0)
    eng.set_obscf(S.simulate.obs_from_expcf(exp0, 300, 5))
    rng = np.random.default_rng(0)
    start = np.clip(x0 * rng.uniform(0.7, 1.3, len(x0)), 0, upper)
    optbl = {"nparams": int(len(x0)), "f_start": float(eng.eval(start)[0]), "f_at_simulating_x": float(eng.eval(x0)[0])}
    for _ in range(3):
        eng.eval_grad(start)
    t0 = time.perf_counter()
    for _ in range(20):
        eng.eval_grad(start)
    optbl["objective_plus_gradient_pass_us"] = (time.perf_counter() - t0) / 20 * 1e6
    eng.optbl(start, 1e-6, 1e-6, 1e-2, 1e-3)                      # untimed: first use of some kernel instantiations
    for name, tol in (("search_tolerances", (1e-6, 1e-6, 1e-2, 1e-3)), ("final_tolerances", (1e-12, 1e-10, 1e-10, 1e-10))):
        t0 = time.perf_counter()
        _, r = eng.optbl(start, *tol)
        r["seconds"] = time.perf_counter() - t0
        optbl[name] = r
    rel = abs(res["default"]["value"] - res["per_quartet_kernel"]["value"]) / abs(res["per_quartet_kernel"]["value"])
    return {"optbl": optbl, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": ex.get("dram_bytes"), "peak_source": src, "kernel": "k_eval_direct on the per-quartet table (SNAQ_OPT_PER_QUARTET)",
            "workload": f"100-taxon h=3, {len(qt)} quartets, P=1 (one objective call), L2 flushed before every timed call",
            "per_quartet_kernel": res["per_quartet_kernel"], "default_run_table": res["default"], "rel_diff_default_vs_per_quartet": rel}


def gather_max(dist, torch, world, v, dev):
    if world == 1:
        return v
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sharded_leg(S, local, rank, world, dist, torch, flush):
    """BASELINE.json configs[4]: 200 taxa, h=5, 64,684,950 quartets split over the N GPUs (contiguous blocks); every rank
    evaluates the same parameter vectors and the library combines the per-candidate partial sums itself: a fused
    peer-memory epilogue for the single objective call, ncclAllReduce for the batch of 64.  Strong scaling: the table is
    fixed, N grows.  Timed on the device (CUDA events on the launching stream), max over ranks.  Both table modes: the
    default run table (each rank sums the weights of ITS quartets per distinct program: the evaluation is then a few tens of
    microseconds on any N, what shards is the table build and the memory) and the per-quartet kernels, where the
    evaluation itself scales with the shard."""
    n, h = 200, 5
    net = S.simulate.random_level1_network(n, h, 20261022)
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = S.simulate.all_quartets(n)
    nq = len(qt)
    rng = np.random.default_rng(3)
    obs = rng.random((nq, 3)) + 0.05
    obs /= obs.sum(axis=1, keepdims=True)
    dev = torch.device("cuda", local)
    data = S.DataCF(taxa, qt, obs)
    flat = net.flatten(taxa)
    out = {"workload": f"200-taxon h={h}, {nq} quartets in {world} contiguous shard(s), same parameter vectors on every rank",
           "scaling": "strong"}
    stream = torch.cuda.Stream(dev)
    allvals = []
    for mode, dd in (("default", None), ("per_quartet", False)):
        eng = S.Engine(local, rank=rank, world_size=world, dedup=dd)
        eng.set_data(data)
        t0 = time.perf_counter()
        eng.set_network(flat)
        t_build = time.perf_counter() - t0
        if world > 1:
            ids = [S.Engine.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            eng.comm_init(ids[0])
        x0, _, upper = eng.get_params()
        X = S.simulate.parameter_batch(x0, upper, 64, 5)
        res = {"comm": eng.comm_info(), "descriptor_build_s": gather_max(dist, torch, world, t_build, dev),
               "distinct_programs_rank0": int(eng.counters()["distinct_programs"])}
        vals = {}
        for P in (1, 64):
            dX = torch.from_numpy(X[:P].copy()).to(dev)
            dout = torch.zeros(P, dtype=torch.float64, device=dev)
            if world > 1:
                dist.barrier()
            ms = float(timed_launches(torch, eng, dX, dout, stream, None, 20, warm=5).mean())
            ms = gather_max(dist, torch, world, ms, dev)
            vals[P] = dout.cpu().numpy().copy()
            res[f"P{P}"] = {"ms_per_launch": ms, "evals_per_s": nq * P / (ms * 1e-3), "value0": float(vals[P][0])}
            t0 = time.perf_counter()
            for _ in range(10):
                eng.eval(X[:P])
            res[f"P{P}"]["host_call_ms"] = gather_max(dist, torch, world, (time.perf_counter() - t0) / 10 * 1e3, dev)
        eng.check_status()
        if world > 1:
            mine = np.concatenate([vals[1], vals[64]]).tobytes()
            allv = [None] * world
            dist.all_gather_object(allv, mine)
            res["identical_on_all_ranks"] = all(b == allv[0] for b in allv)
        res["P1_equals_P64_row0_rel"] = float(abs(vals[1][0] - vals[64][0]) / abs(vals[64][0]))
        allvals.append(vals[64])
        out[mode] = res
        del eng
        torch.cuda.empty_cache()
    out["max_rel_diff_default_vs_per_quartet"] = float(np.max(np.abs(allvals[0] - allvals[1]) / np.abs(allvals[1])))
    return out


def search_chains_leg(S, local, rank, world, dist, torch, nchains=8, K=50):
    """The engine's side of `snaq! hmax=3 runs=8` on 100 taxa (BASELINE.json configs[2]) without the Julia control plane:
    8 independent hill-climbing chains (one per run, dealt to the GPUs), each proposing K = 50 candidate topologies (a
    nearest-neighbour interchange of the current network), and for every candidate what optTopLevel! does
    (src/snaq_optimization.jl:1380-1412): descriptor table for the candidate + optBL! at the search tolerances, accept iff
    the pseudo-deviance improves by more than liktolAbs = 1e-6, else undo.  Host-side Python (move + flattening of the
    network) is the stand-in for the Julia search and is timed apart."""
    n, net0, taxa, qt = table100(S)
    eng = S.Engine(local)
    eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
    eng.set_network(net0.flatten(taxa))
    x_true, _, _ = eng.get_params()
    e0, _, _ = eng.eval_quartets(x_true)
    eng.set_obscf(S.simulate.obs_from_expcf(e0, 300, 5))
    t_net = t_opt = t_host = 0.0
    ncand = nacc = npart = 0
    launches = 0
    results = []
    if world > 1:
        dist.barrier()
    t_wall0 = time.perf_counter()
    for chain in range(rank, nchains, world):
        rng = np.random.default_rng(1000 + chain)
        net = S.simulate.random_level1_network(n, 3, 20261021)          # the simulating topology ...
        for _ in range(6):
            S.simulate.nni_move(net, rng)                                # ... six interchanges away: the start of the run
        eng.set_network(net.flatten(taxa))
        x, _, _ = eng.get_params()
        xcur, r = eng.optbl(x)
        fcur = r["fmin"]
        net.update_parameters(list(xcur))
        for _ in range(K):
            th = time.perf_counter()
            tok = S.simulate.nni_move(net, rng)
            flat = net.flatten(taxa)
            t1 = time.perf_counter()
            touched = eng.patch_network(flat)                           # extractQuartet! for the edited network, incrementally
            npart += touched >= 0
            x, _, _ = eng.get_params()
            t2 = time.perf_counter()
            xn, r = eng.optbl(x)
            t3 = time.perf_counter()
            ncand += 1
            launches += r["nlaunches"]
            if r["fmin"] < fcur - 1e-6:
                fcur = r["fmin"]
                net.update_parameters(list(xn))
                nacc += 1
            else:
                S.simulate.nni_undo(net, tok)
            t4 = time.perf_counter()
            t_host += (t1 - th) + (t4 - t3)
            t_net += t2 - t1
            t_opt += t3 - t2
        results.append(fcur)
    wall = time.perf_counter() - t_wall0
    dev = torch.device("cuda", local)
    wall = gather_max(dist, torch, world, wall, dev)
    mine = [ncand, nacc, t_net, t_opt, t_host, launches, npart]
    allv = [mine]
    if world > 1:
        allv = [None] * world
        dist.all_gather_object(allv, mine)
    tot = np.sum(np.array(allv, dtype=float), axis=0)
    return {"workload": f"{nchains} chains x {K} NNI candidates on the 100-taxon h=3 table ({len(qt)} quartets), search tolerances, "
                        f"chains dealt to {world} GPU(s)", "wall_s": wall, "candidates": int(tot[0]), "accepted": int(tot[1]),
            "candidates_per_s": tot[0] / wall, "engine_ms_per_candidate": {"descriptor_table": tot[2] / tot[0] * 1e3, "optBL": tot[3] / tot[0] * 1e3},
            "host_standin_ms_per_candidate": tot[4] / tot[0] * 1e3, "engine_launches_per_optBL": tot[5] / tot[0],
            "incremental_patches": int(tot[6]), "note": "descriptor_table = snaq_patch_network: classification only for quartets with a taxon "
            "whose root path changed, the rest of the build (sort, emission, run table, weights) in full; a rebuild from scratch is ~1.45 ms",
            "final_deviance_rank0": results}


def bootstrap_leg(S, local, rank, world, dist, torch, nrep=96):
    """The engine's side of `bootsnaq` with 96 replicates on a 50-taxon table with credibility intervals (BASELINE.json
    configs[3]), replicates dealt to the GPUs: per replicate sampleCFfromCI! + readtableCF! on the device
    (snaq_resample_obscf), the updateBL! table of the start tree (snaq_update_bl), and optBL! of the start network at the
    search and then the final tolerances."""
    n, h = 50, 2
    net = S.simulate.random_level1_network(n, h, 20261022)
    tree = S.simulate.random_level1_network(n, 0, 20261022)
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = S.simulate.all_quartets(n)
    eng = S.Engine(local)
    eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
    fnet, ftree = net.flatten(taxa), tree.flatten(taxa)
    eng.set_network(fnet)
    x0, _, upper = eng.get_params()
    obs = S.simulate.obs_from_expcf(eng.eval_quartets(x0)[0], 300, 4)
    eng.set_ci(np.clip(obs - 0.05, 0, 1), np.clip(obs + 0.05, 0, 1))           # CI = obs -/+ 0.05 clipped
    eng.resample_obscf(0)
    eng.optbl(x0)                                                              # untimed warm-up
    if world > 1:
        dist.barrier()
    fm = []
    t0 = time.perf_counter()
    for rep in range(rank, nrep, world):
        eng.resample_obscf(5000 + rep)
        eng.set_network(ftree)
        eng.update_bl()
        eng.set_network(fnet)
        xm, r = eng.optbl(x0)
        xm, r = eng.optbl(xm, 1e-12, 1e-10, 1e-10, 1e-10)
        fm.append(r["fmin"])
    wall = time.perf_counter() - t0
    wall = gather_max(dist, torch, world, wall, torch.device("cuda", local))
    return {"workload": f"{nrep} bootstrap replicates at 50 taxa ({len(qt)} quartets, h={h}): CI resampling + updateBL! table + optBL! "
                        f"(search, then final tolerances), replicates dealt to {world} GPU(s)", "wall_s": wall,
            "replicates_per_s": nrep / wall, "ms_per_replicate_per_gpu": wall / max(1, len(range(rank, nrep, world))) * 1e3,
            "deviance_range_rank0": [float(min(fm)), float(max(fm))]}


if __name__ == "__main__":
    main()
