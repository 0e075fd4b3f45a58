#!/bin/bash
# Diagnostic GPU visit: init stability sweep, per-phase host/device split, torch.profiler summary, per-layer conv table.
mkdir -p gpurun_out
for sc in 1.0 0.1 0.03; do
  MMT_SYNTH_ADAPTOR_SCALE=$sc timeout 300 python scripts/loss_trace.py 40 2000 2>/dev/null | cut -c1-260 > gpurun_out/loss_trace_scale_$sc.log
  echo "scale $sc:"; tail -2 gpurun_out/loss_trace_scale_$sc.log
done
timeout 600 python scripts/profile_step.py 3 gpurun_out/profile_step > gpurun_out/profile_step.log 2>&1
cat gpurun_out/profile_step.log | grep -v Warning | tail -20
timeout 600 python scripts/conv_bench.py --all > gpurun_out/conv_bench.log 2>&1
tail -40 gpurun_out/conv_bench.log
