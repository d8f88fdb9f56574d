"""Turns the committed ncu captures of bench.py's timed kernels into profiles/r02_executed_flops.json: per launch, the
thread-level DFMA / DADD / DMUL instruction counts (executed double-precision flops = 2 DFMA + DADD + DMUL), DRAM bytes and
the FP64-pipe activity.  bench.py divides these by the launch time it measures live.
Usage: python profiles/make_executed.py [tag]  (reads gpurun_out/<tag>_*.ncu-rep written by profiles/capture.sh)"""
import csv, io, json, subprocess, sys

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2]

def f(hdr, row, units, k):
    v = float(row[hdr.index(k)].replace(",", ""))
    u = units[hdr.index(k)]
    if u == "Mbyte": v *= 1e6
    if u == "Kbyte": v *= 1e3
    if u == "Gbyte": v *= 1e9
    if u == "us": v *= 1e-6
    if u == "ms": v *= 1e-3
    if u == "ns": v *= 1e-9
    return v

TAG = sys.argv[1] if len(sys.argv) > 1 else "r02t"
out = {}
for key, rep in (("config2", f"gpurun_out/{TAG}_default_config2.ncu-rep"), ("north_star", f"gpurun_out/{TAG}_default_northstar.ncu-rep"),
                 ("config2_per_quartet", f"gpurun_out/{TAG}_perq_config2.ncu-rep"),
                 ("north_star_per_quartet", f"gpurun_out/{TAG}_perq_northstar.ncu-rep"),
                 ("hbm_p1", f"gpurun_out/{TAG}_direct_p1.ncu-rep")):
    hdr, units, row = raw(rep)
    g = lambda k: f(hdr, row, units, k)
    out[key] = {"kernel": row[hdr.index("Kernel Name")][:80], "grid": row[hdr.index("launch__grid_size")],
                "dfma": g("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"), "dadd": g("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"),
                "dmul": g("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"),
                "dram_bytes": g("dram__bytes_read.sum") + g("dram__bytes_write.sum"),
                "fp64_pipe_pct": g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "warp_inst": g("smsp__inst_executed.sum"), "ncu_duration_s": g("gpu__time_duration.sum"),
                "source": rep.split("/")[-1].replace(".ncu-rep", "") + ": ncu --set full capture by profiles/capture.sh, summarised in profiles/" + TAG + "_ncu_summary.md"}
import os
json.dump(out, open(os.path.join(os.environ.get("SNAQ_PROFILE_OUT", "profiles"), "r02_executed_flops.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
