#!/bin/bash
# Profiles the search kernel of bench.py on one B200 (run under gpurun).  Follows
# /opt/skills/guides/B200_PROFILING.md: plain run first, then the launch list of the same command.
# Outputs land in gpurun_out/.  (The --set full captures of the individual kernels are taken with
# scripts/bench_quick.py / probe_adc.py / probe_encoder_small.py, see profiles/README.md.)
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-uniform --no-adc --no-encoder --no-hnsw --cpu-seconds 1"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
tail -c 1200 gpurun_out/plain.log
# 265 search launches belong to the graph build (setup); then: ef calibration (one launch per rung),
# stats pass, 3 warm-up, 3 timed, e2e (3 + 3) and the cpu-baseline id check.
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'leann_search|merge_topk|pq_tables' -s 265 -c 60 \
    --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
