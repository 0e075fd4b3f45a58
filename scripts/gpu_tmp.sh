#!/bin/bash
echo "== EAGER module loading"
time (CUDA_MODULE_LOADING=EAGER MMT_BENCH_TRACE=1 timeout 600 python bench.py --steps 40 --no-cpu-baseline --no-roofline 2>&1 >/dev/null | grep "^step" | awk '{printf "%s ", $4} END {print ""}')
echo "== no clock sampler (lazy loading)"
MMT_NO_CLOCKS=1 MMT_BENCH_TRACE=1 timeout 600 python bench.py --steps 40 --no-cpu-baseline --no-roofline 2>&1 >/dev/null | grep "^step" | awk '{printf "%s ", $4} END {print ""}'
