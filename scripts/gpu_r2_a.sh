#!/bin/bash
# round-2 GPU call A: full parity suite + stall diagnostics (per-step trace with allocator counters) + bench + ncu launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/a_pytest.log
MMT_BENCH_TRACE=1 python bench.py --steps 40 --no-cpu-baseline --no-roofline > gpurun_out/a_trace_default.json 2> gpurun_out/a_trace_default.log
MMT_BENCH_TRACE=1 PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True python bench.py --steps 40 --no-cpu-baseline --no-roofline > gpurun_out/a_trace_expandable.json 2> gpurun_out/a_trace_expandable.log
MMT_BENCH_TRACE=1 MMT_PREALLOC_GB=24 python bench.py --steps 40 --no-cpu-baseline --no-roofline > gpurun_out/a_trace_prealloc.json 2> gpurun_out/a_trace_prealloc.log
python bench.py --no-cpu-baseline > gpurun_out/a_bench.json 2> gpurun_out/a_bench.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/a_ncu_bench.log 2>&1
tail -5 gpurun_out/a_pytest.log
