#!/bin/bash
# ncu captures behind profiles/<tag>_ncu_summary.md and profiles/r02_executed_flops.json (run on the GPU box, after the same
# commands have exited 0 without ncu):   bash profiles/capture.sh r02t
TAG=${1:-r02t}
O=gpurun_out
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum
FULL="ncu --set full --metrics $M --clock-control none --import-source on -f"
python bench.py --steps 2 --warmup 3 --no-sharded --no-search --no-cpu-baseline --no-candidates > $O/${TAG}_bench_plain.json 2> $O/${TAG}_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-sharded --no-search --no-cpu-baseline --no-candidates > $O/${TAG}_launches.log 2>&1
# default mode (run table): the main line and the north-star leg
$FULL -k regex:k_eval_lanes -s 4 -c 1 -o $O/${TAG}_default_config2 python profiles/run_runs.py 30 2 4096 20261020 -1 > $O/${TAG}_c1.log 2>&1
$FULL -k regex:k_eval_lanes -s 4 -c 1 -o $O/${TAG}_default_northstar python profiles/run_runs.py 100 3 256 20261021 -1 > $O/${TAG}_c2.log 2>&1
# per-quartet kernels
$FULL -k regex:k_eval_lanes -s 4 -c 1 -o $O/${TAG}_perq_config2 python profiles/run_runs.py 30 2 4096 20261020 1 > $O/${TAG}_c3.log 2>&1
$FULL -k regex:k_eval_runs -s 4 -c 1 -o $O/${TAG}_perq_northstar python profiles/run_runs.py 100 3 256 20261021 1 > $O/${TAG}_c4.log 2>&1
$FULL -k regex:k_eval_direct -s 5 -c 1 -o $O/${TAG}_direct_p1 python profiles/run_direct_p1.py > $O/${TAG}_c5.log 2>&1
# table build (k_build_count at 100 taxa)
$FULL -k regex:k_build_count -s 3 -c 1 -o $O/${TAG}_build_count python profiles/run_build_prof.py > $O/${TAG}_c6.log 2>&1
# summaries are made here (the reports are too large to travel back: gpurun_out/ is capped at 64 MiB); two reports are kept
SNAQ_PROFILE_OUT=$O python profiles/make_executed.py $TAG > $O/${TAG}_executed.log 2>&1
SNAQ_PROFILE_OUT=$O python profiles/summarize.py $TAG >> $O/${TAG}_executed.log 2>&1
rm -f $O/${TAG}_default_northstar.ncu-rep $O/${TAG}_perq_config2.ncu-rep $O/${TAG}_perq_northstar.ncu-rep $O/${TAG}_direct_p1.ncu-rep
ls -la $O/${TAG}_*
