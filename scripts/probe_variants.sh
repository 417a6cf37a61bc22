#!/bin/bash
# Dev probe: the same ADC traversal + rerank measurement against several builds of the library (.variants/*.so).
for lib in .variants/*.so islands_b200/lib/libislands_b200.so; do
  echo "== $lib"
  ISL_DEV_LIB_PATH=$PWD/$lib ISL_DEV_ALLOW_MISSING=1 MS=32 KSUB=128 EFS=192,320 python scripts/probe_adc.py 2>&1 | grep kernel_ms
done
