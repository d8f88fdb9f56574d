"""Turns gpurun_out/*.ncu-rep and the launch list into the text summaries committed under profiles/.
Usage: python profiles/summarize.py r01   (reads gpurun_out/r01_*.{ncu-rep,csv})"""
import collections
import csv
import io
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
GO = "gpurun_out"

KEYS = ["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_read.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed_op_shared_atom.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def kernel_summary(rep, title, fh):
    hdr, units, data = raw(rep)
    name_i = hdr.index("Kernel Name")
    for row in data:
        fh.write(f"## {title}: {row[name_i][:110]}\n\n| metric | value | unit |\n|---|---|---|\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                fh.write(f"| {k} | {row[i]} | {units[i]} |\n")
        fh.write("\nwarp stall reasons (warps per issue-active cycle):\n\n")
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h:
                try:
                    v = float(row[i])
                except ValueError:
                    continue
                if v >= 0.05:
                    fh.write(f"- {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.2f}\n")
        fh.write("\n")
    # instruction mix from the source page
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not idx:
        return
    h = rows[idx[0]]
    si, ei = h.index("Source"), h.index("Instructions Executed")
    end = idx[1] - 1 if len(idx) > 1 else len(rows)
    ops = collections.Counter()
    tot = 0
    nstat = 0
    for r in rows[idx[0] + 1:end]:
        if len(r) <= ei:
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si])
        if not m:
            continue
        e = int(r[ei])
        ops[m.group(2).split(".")[0]] += e
        tot += e
        nstat += 1
    fh.write(f"executed warp instructions {tot}, static SASS instructions {nstat}; mix (first launch):\n\n")
    for op, c in ops.most_common(14):
        fh.write(f"- {op}: {c / tot:.3f}\n")
    fh.write("\n")


def launches(path, fh):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        a = agg.setdefault(r[ki][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    fh.write("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write(f"| `{k}` | {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.1f} | {a[1] / tot:.3f} |\n")
    fh.write("\n")


import os
OUTDIR = os.environ.get("SNAQ_PROFILE_OUT", "profiles")
with open(f"{OUTDIR}/{tag}_ncu_summary.md", "w") as fh:
    fh.write(f"# ncu summaries ({tag})\n\nCommand profiled: `python bench.py --steps 2 --warmup 3 --no-sharded --no-search --no-cpu-baseline --no-candidates` on one B200 "
             "(`--clock-control none`; per-launch times are cold-cache and serialised: compare shares, not absolutes).\n\n")
    fh.write("## launch list (`ncu --metrics gpu__time_duration.sum`)\n\n")
    launches(f"{GO}/{tag}_launches.csv", fh)
    for suffix, title in (("default_config2", "DEFAULT (run table) k_eval_lanes, config 2: 27,405 quartets = 1,290 programs x 4096 vectors; the main line of bench.py"),
                          ("default_northstar", "DEFAULT (run table) k_eval_lanes, north star: 100 taxa, 3,921,225 quartets = 9,430 programs x 256 vectors; bench.py `north_star`"),
                          ("perq_config2", "PER-QUARTET k_eval_lanes, config 2: 27,405 quartets x 4096 vectors; bench.py `per_quartet.config2`"),
                          ("perq_northstar", "PER-QUARTET k_eval_runs, north star: 3,921,225 quartets x 256 vectors; bench.py `per_quartet.north_star`"),
                          ("build_count", "k_build_count (descriptor build, 100 taxa, 3,921,225 quartets)"),
                          ("lanes", "k_eval_lanes (config 2: 27,405 quartets x 4096 vectors)"),
                          ("lanes_config2", "k_eval_lanes (config 2: 27,405 quartets x 4096 vectors; the main line of bench.py)"),
                          ("runs_northstar", "k_eval_runs (north star: 100 taxa, 3,921,225 quartets x 256 vectors; bench.py `north_star`)"),
                          ("direct_p1", "k_eval_direct (100 taxa, 3.92 M quartets, P = 1, per-quartet table; bench.py `roofline_hbm`)"),
                          ("quartets", "first-version single-call kernel (100 taxa, 3.92 M quartets, P=1)"),
                          ("grad", "k_eval_grad (100 taxa, 3.92 M quartets, objective + gradient)"),
                          ("direct", "k_eval_direct (100 taxa, 3.92 M quartets, P=1)")):
        if os.path.exists(f"{GO}/{tag}_{suffix}.ncu-rep"):
            kernel_summary(f"{GO}/{tag}_{suffix}.ncu-rep", title, fh)
print("wrote", f"{OUTDIR}/{tag}_ncu_summary.md")
