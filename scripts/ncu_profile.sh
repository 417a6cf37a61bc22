#!/bin/bash
# Profiles the search kernel of bench.py on one B200 (run under gpurun).  Follows
# /opt/skills/guides/B200_PROFILING.md: plain run first, then the launch list, then one
# --set full capture of the dominant kernel.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --ef 128 --no-uniform --cpu-seconds 1"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
tail -1 gpurun_out/plain.log
# 265 search launches belong to the graph build (setup); the rest: 1 ef probe, 1 stats pass,
# 3 warm-up, 3 timed, e2e and the cpu_baseline check.
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'leann_search|merge_topk|pq_tables' -s 265 -c 40 \
    --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:leann_search -s 268 -c 1 \
    -o gpurun_out/prof_search -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
