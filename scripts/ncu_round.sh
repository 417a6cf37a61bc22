#!/bin/bash
# One gpurun call that regenerates the round's profile inputs (one GPU):
#  1. scripts/ncu_profile.sh   -> gpurun_out/plain.log (bench line), gpurun_out/launches.csv (launch list of the same command)
#  2. --set full of one exact-traversal launch at the headline workload (scripts/bench_quick.py, EFS=104)
#  3. scripts/ncu_adc.sh       -> --set full of adc_traverse_kernel at 1M x 768, m=32, ksub=128, ef=192
# Summaries are made afterwards on the CPU box (scripts/ncu_summary.py, launch_shares.py, ncu_lines.py) and copied to profiles/.
set -u
mkdir -p gpurun_out
bash scripts/ncu_profile.sh
EFS=${EF_SEARCH:-104} python scripts/bench_quick.py > gpurun_out/quick_plain.log 2>&1 || { tail -5 gpurun_out/quick_plain.log; exit 1; }
tail -1 gpurun_out/quick_plain.log
# 265 build launches, then the stats pass, then the timed launches: capture the second timed one
EFS=${EF_SEARCH:-104} ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'leann_search_kernel' -s 267 -c 1 -f -o gpurun_out/prof_search \
    python scripts/bench_quick.py > gpurun_out/quick_ncu.log 2>&1
echo "search ncu rc=$?"
bash scripts/ncu_adc.sh
