#!/usr/bin/env python
"""bench.py — quartet pseudolikelihood evaluations per second (FP64).

One "step" = one pass of the hot path over one batch: every quartet of the table
evaluated under every parameter vector of an optBL!-style batch (BASELINE.json
configs[1]: synthetic 30-taxon level-1 network, h=2, all 27,405 quartets x 4096
parameter vectors = 112.3 M evals per step).

  python bench.py --gpus N --steps K --warmup W            # the engine
  python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port)

N>1 (torchrun): the path shards by independent units — every rank evaluates its own
batch on its own GPU with no data-path collective (independent snaq! runs / bootstrap
replicates) => weak scaling; value = all ranks' evals / max-over-ranks time.
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
# algorithmic work per evaluation: F = 110 + 75*f5 + 90*cbar flop
FLOP_TREE, FLOP_T5_EXTRA, FLOP_PER_TERM = 110.0, 75.0, 90.0


NCU_TRAFFIC_LANES_BYTES = 4477952          # profiles/r01h_ncu_summary.md: 4.477952 MB read, 0 written per launch
NCU_TRAFFIC_DIRECT_BYTES = 113707776       # profiles/r01h_ncu_summary.md: 109.73 MB read + 3.97 MB written per launch


def workload(ntaxa=NTAXA, nhyb=NHYB, nvec=NVEC):
    import snaq_jl_b200 as S
    net = S.simulate.random_level1_network(ntaxa, nhyb, NET_SEED)
    taxa = [f"t{i + 1}" for i in range(ntaxa)]
    qt = S.simulate.all_quartets(ntaxa)
    return S, net, taxa, qt


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-vectors", type=int, default=128, help="parameter vectors per step of the CPU arms")
    ap.add_argument("--cpu-vectors", type=int, default=4096, help="parameter vectors of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hbm", action="store_true", help="skip the secondary 100-taxon P=1 HBM-roofline measurement")
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
    launches = (c1["eval_launches"] - c0["eval_launches"]) * 3      # k_expand_x + k_eval_lanes + k_reduce_partials_t per step

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

    if rank == 0:
        # ---- roofline of the dominant kernel (k_eval_lanes): FP64 pipe
        desc = eng.descriptors()
        f5 = float(np.mean(desc["which"] == 2))
        cbar = float(np.mean(np.sum((desc["types"] >= 2) & (desc["types"] <= 4), axis=1)))
        flop_per_eval = FLOP_TREE + FLOP_T5_EXTRA * f5 + FLOP_PER_TERM * cbar
        kernel_s = float(ms.mean()) * 1e-3           # per launch (eval kernel + its tiny reduction pass)
        achieved_tf = evals_per_step * flop_per_eval / kernel_s / 1e12
        peak_tf = eng.measure_fp64_peak()
        roofline = {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                    "traffic": NCU_TRAFFIC_LANES_BYTES, "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one k_eval_lanes "
                    "launch in the committed ncu --set full capture (profiles/*_ncu_summary.md); inputs are L2/shared-memory resident", "flop_per_eval": flop_per_eval, "f5": f5, "cbar": cbar,
                    "peak_source": "DFMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                    "note": "achieved = evaluations x SURVEY 8d's flop count per evaluation (one exp/log set per quartet and per "
                            "hybridisation term). The kernel sorts quartets by program and reuses t1/exp/log and the terms while they "
                            "repeat between neighbours, so it executes fewer flops than the convention counts and frac can exceed 1; "
                            "every quartet still gets its own weighted contribution"}
        line = {"metric": "quartet pseudolik evals/s (FP64)", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": float(ms.mean()), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"synthetic {NTAXA}-taxon level-1 network h={NHYB}, {nq} quartets x {P} parameter vectors "
                                       f"per step (BASELINE.json configs[1]); one independent batch per GPU",
                           "nparams": npar, "l2": "flushed between timed iterations (inputs < L2)", "descriptor_build_s": t_build},
                "e2e": {"value": e2e_val, "unit": "evals/s", "h2d_bytes_per_step": int(P * npar * 8), "d2h_bytes_per_step": int(P * 8 + 4)},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
        # ---- the same launch with the reuse between neighbouring quartets switched off (SNAQ_REUSE=0): every quartet gets its
        # own exp/log and its own terms, which is what SURVEY 8d's flop count describes
        try:
            os.environ["SNAQ_REUSE"] = "0"
            eng0 = eng.__class__(local)
            del os.environ["SNAQ_REUSE"]
            eng0.set_data(S.DataCF(taxa, qt, obs))
            eng0.set_network(flat)
            out0 = torch.zeros(P, dtype=torch.float64, device=dev)
            with torch.cuda.stream(stream):
                for _ in range(3):
                    eng0.eval_device(dX, out0, stream)
            torch.cuda.synchronize()
            ev0 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            with torch.cuda.stream(stream):
                for a, b in ev0:
                    flush.fill_(1)
                    a.record(stream)
                    eng0.eval_device(dX, out0, stream)
                    b.record(stream)
            torch.cuda.synchronize()
            ms0 = float(np.mean([a.elapsed_time(b) for a, b in ev0]))
            tf0 = evals_per_step * flop_per_eval / (ms0 * 1e-3) / 1e12
            line["roofline_no_reuse"] = {"bound": "fp64", "value": evals_per_step / (ms0 * 1e-3), "unit_value": "evals/s", "ms_per_step": ms0,
                                         "achieved": tf0, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf0 / peak_tf,
                                         "max_rel_diff_vs_default": float(((out0 - dout).abs() / dout.abs()).max().item()),
                                         "note": "same kernel, same sorted table, reuse disabled: one full evaluation per (quartet, vector)"}
            del eng0
        except Exception as exc:  # noqa: BLE001
            os.environ.pop("SNAQ_REUSE", None)
            line["roofline_no_reuse"] = {"error": str(exc)[:200]}
        # ---- opt-in program deduplication (SNAQ_OPT_DEDUP): the same batch on the table with one entry per DISTINCT program.
        # Reported apart from `value`: it evaluates fewer programs, so it says nothing about the kernel's roofline.
        try:
            line["dedup"] = dedup_leg(S, eng.__class__, local, torch, taxa, qt, obs, flat, dX, dout, stream, flush, evals_per_step,
                                      with_100_taxa=not args.no_hbm)
        except Exception as exc:  # noqa: BLE001
            line["dedup"] = {"error": str(exc)[:200]}
        # ---- secondary: the 100-taxon table (3.92 M quartets): one objective call (P=1, HBM-bound) and a batch
        if not args.no_hbm:
            try:
                line["roofline_hbm"] = hbm_leg(S, eng.__class__, local, torch, peak_tf)
            except Exception as exc:  # noqa: BLE001
                line["roofline_hbm"] = {"error": str(exc)[:200]}
        # ---- secondary: candidate topologies optimised concurrently on this GPU (snaq_optbl_topologies)
        if not args.no_hbm:
            try:
                line["candidates"] = candidates_leg(S, eng, taxa, qt, obs)
            except Exception as exc:  # noqa: BLE001
                line["candidates"] = {"error": str(exc)[:200]}
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, dt = cpu_port(S, net, taxa, qt, obs, Xh[: args.cpu_vectors], cores)
            line["cpu_baseline"] = {"value": v, "unit": "evals/s", "cores": cores, "kind": "port", "seconds": dt,
                                    "sample": f"{nq} quartets x first {args.cpu_vectors} of the {P} parameter vectors, oracle port "
                                              f"(C++ restatement of the reference, OpenMP over quartets)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def dedup_leg(S, Engine, local, torch, taxa, qt, obs, flat, dX, dout, stream, flush, evals_per_step, with_100_taxa=True):
    """Same work as the main step, on an engine created with dedup=True (include/snaq_b200.h: SNAQ_OPT_DEDUP)."""
    eng = Engine(local, dedup=True)
    eng.set_data(S.DataCF(taxa, qt, obs))
    eng.set_network(flat)
    out = {"note": "opt-in: quartets with identical evaluation programs are evaluated once with the sum of their weights; "
                   "'effective' rates count every quartet of the table", "quartets": int(len(qt)),
           "distinct_programs": int(eng.counters()["distinct_programs"])}
    ref = dout.clone()
    with torch.cuda.stream(stream):
        for _ in range(3):
            eng.eval_device(dX, dout, stream)
    torch.cuda.synchronize()
    out["max_rel_diff_vs_plain"] = float(((dout - ref).abs() / ref.abs()).max().item())
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    with torch.cuda.stream(stream):
        for a, b in ev:
            flush.fill_(1)
            a.record(stream)
            eng.eval_device(dX, dout, stream)
            b.record(stream)
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    out["config2_ms_per_step"] = ms
    out["config2_effective_evals_per_s"] = evals_per_step / (ms * 1e-3)
    if with_100_taxa:
        n = 100
        net = S.simulate.random_level1_network(n, 3, 20261021)
        tx = [f"t{i + 1}" for i in range(n)]
        q100 = S.simulate.all_quartets(n)
        e2 = Engine(local, dedup=True)
        e2.set_data(S.DataCF(tx, q100, np.full((len(q100), 3), 1 / 3)))
        fl = net.flatten(tx)
        e2.set_network(fl)
        t0 = time.perf_counter()
        for _ in range(5):
            e2.set_network(fl)
        out["t100_set_network_ms"] = (time.perf_counter() - t0) / 5 * 1e3
        out["t100_quartets"] = int(len(q100))
        out["t100_distinct_programs"] = int(e2.counters()["distinct_programs"])
        x0, _, upper = e2.get_params()
        exp0, _, _ = e2.eval_quartets(x0)
        e2.set_obscf(S.simulate.obs_from_expcf(exp0, 300, 5))
        rng = np.random.default_rng(0)
        start = np.clip(x0 * rng.uniform(0.7, 1.3, len(x0)), 0, upper)
        for _ in range(3):
            e2.eval(start)
            e2.eval_grad(start)
        t0 = time.perf_counter()
        for _ in range(50):
            e2.eval(start)
        out["t100_objective_call_from_host_us"] = (time.perf_counter() - t0) / 50 * 1e6
        t0 = time.perf_counter()
        for _ in range(50):
            e2.eval_grad(start)
        out["t100_objective_plus_gradient_pass_us"] = (time.perf_counter() - t0) / 50 * 1e6
        e2.optbl(start, 1e-6, 1e-6, 1e-2, 1e-3)                   # untimed warm-up
        for name, tol in (("search_tolerances", (1e-6, 1e-6, 1e-2, 1e-3)), ("final_tolerances", (1e-12, 1e-10, 1e-10, 1e-10))):
            t0 = time.perf_counter()
            _, r = e2.optbl(start, *tol)
            out["t100_optbl_" + name] = {"seconds": time.perf_counter() - t0, "fmin": r["fmin"], "nlaunches": r["nlaunches"]}
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


def hbm_leg(S, Engine, local, torch, peak_tf=None):
    """P=1 on 100 taxa / h=3 / 3,921,225 quartets: one optBL! objective call.  Bytes per launch = what the
    layout makes k_eval_direct read once (engine counter: 16 B weights + 8 B per program chunk for additive
    quartets, 37 B + programs for the rest); peak = MEASURED_PEAKS.json hbm_gbs.  Timed cold (L2 flushed before
    every call: the table is smaller than L2) and warm (back-to-back calls, as inside one optBL!)."""
    n = 100
    net = S.simulate.random_level1_network(n, 3, 20261021)
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = S.simulate.all_quartets(n)
    eng = Engine(local)
    eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
    t0 = time.perf_counter()
    eng.set_network(net.flatten(taxa))
    t_build = time.perf_counter() - t0
    x0, _, upper = eng.get_params()
    dev = torch.device("cuda", local)
    dX = torch.from_numpy(x0.reshape(1, -1)).to(dev)
    dout = torch.zeros(1, dtype=torch.float64, device=dev)
    stream = torch.cuda.Stream(dev)
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    with torch.cuda.stream(stream):
        for _ in range(5):
            eng.eval_device(dX, dout, stream)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    with torch.cuda.stream(stream):
        for a, b in ev:
            flush.fill_(1)                            # L2 flush: every timed call streams from HBM
            a.record(stream)
            eng.eval_device(dX, dout, stream)
            b.record(stream)
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    evw = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
    with torch.cuda.stream(stream):
        for a, b in evw:
            a.record(stream)
            eng.eval_device(dX, dout, stream)
            b.record(stream)
    torch.cuda.synchronize()
    ms_warm = float(np.mean([a.elapsed_time(b) for a, b in evw]))
    bytes_alg = eng.counters()["single_call_bytes"]
    peak = None
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "MEASURED_PEAKS.json hbm_gbs"
    except Exception:  # noqa: BLE001
        peak, src = 6650.0, "fallback of B200_PROFILING.md"
    ach = bytes_alg / (ms * 1e-3) / 1e9
    # batched logPseudoLik on the same table: 256 parameter vectors per launch (lane-per-candidate kernel)
    Pb = 256
    Xb = torch.from_numpy(S.simulate.parameter_batch(x0, upper, Pb, 11)).to(dev)
    outb = torch.zeros(Pb, dtype=torch.float64, device=dev)
    for _ in range(2):
        eng.eval_device(Xb, outb, stream)
    torch.cuda.synchronize()
    evb = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a, b in evb:
        a.record(stream)
        eng.eval_device(Xb, outb, stream)
        b.record(stream)
    torch.cuda.synchronize()
    eng.check_status()
    msb = float(np.mean([a.elapsed_time(b) for a, b in evb]))
    desc = eng.descriptors()
    f5 = float(np.mean(desc["which"] == 2))
    cbar = float(np.mean(np.sum((desc["types"] >= 2) & (desc["types"] <= 4), axis=1)))
    fpe = FLOP_TREE + FLOP_T5_EXTRA * f5 + FLOP_PER_TERM * cbar
    batched = {"P": Pb, "ms_per_launch": msb, "evals_per_s": len(qt) * Pb / (msb * 1e-3), "flop_per_eval": fpe,
               "achieved_tflops": len(qt) * Pb * fpe / (msb * 1e-3) / 1e12}
    if peak_tf:
        batched["frac_fp64"] = batched["achieved_tflops"] / peak_tf
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
    return {"batched": batched, "optbl": optbl, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": NCU_TRAFFIC_DIRECT_BYTES, "peak_source": src,
            "workload": f"100-taxon h=3, {len(qt)} quartets, P=1 (one objective call), L2 flushed before every timed call",
            "ms_per_call": ms, "ms_per_call_warm_l2": ms_warm, "evals_per_s_warm_l2": len(qt) / (ms_warm * 1e-3),
            "achieved_at_survey_8d_56B_per_quartet": len(qt) * 56 / (ms * 1e-3) / 1e9,
            "evals_per_s": len(qt) / (ms * 1e-3), "bytes_per_quartet": bytes_alg / len(qt), "descriptor_build_s": t_build}


if __name__ == "__main__":
    main()
