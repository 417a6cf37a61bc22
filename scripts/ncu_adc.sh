#!/bin/bash
# --set full capture of the lean ADC traversal kernel (adc_traverse.cuh, register bag) at 1M x 768 (run under gpurun,
# one GPU; the same command must have exited 0 without ncu first).  Report lands in gpurun_out/.
set -u
mkdir -p gpurun_out
export MS=${MS:-32} KSUB=${KSUB:-128} EFS=${EFS:-192}
python scripts/probe_adc.py > gpurun_out/adc_plain.log 2>&1 || { tail -5 gpurun_out/adc_plain.log; exit 1; }
tail -2 gpurun_out/adc_plain.log
# the graph build launches the MODE 0 kernel; only the traversal kernel carries "3, 6>" / "3, 4>" ...
# probe_adc.py calls the search twice per ef: with statistics (visited bitset) and without (bitset-free): SKIP=1 captures
# the second, SKIP=0 the first
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'adc_traverse_kernel' -s ${SKIP:-1} -c 1 -f -o gpurun_out/${OUT:-prof_adc} \
    env python scripts/probe_adc.py > gpurun_out/adc_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/adc_ncu.log
