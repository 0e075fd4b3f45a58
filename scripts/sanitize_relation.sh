#!/bin/bash
# Next step for the open defect of DESIGN.md §9 (config 4 diverges 4-15 steps into bench.py's end-to-end loop on the GPU while the
# CPU oracle is stable): the end-to-end loop of `--workload mt_relation` under compute-sanitizer.  initcheck first (an uninitialised
# read whose garbage is occasionally a NaN bit pattern — e.g. recycled int64 -1 / -2 matcher outputs — fits "finite forward, poisoned
# parameters"), then racecheck / memcheck.  Needs a B200 and ~10 minutes per tool; not run in round 2 (GPU budget spent).
mkdir -p gpurun_out
for tool in initcheck memcheck; do
  MMT_BENCH_LOSS_TRACE=1 MMT_NO_CLOCKS=1 timeout 1500 compute-sanitizer --tool $tool --print-limit 40 --log-file gpurun_out/sanitize_relation_$tool.log \
    python bench.py --workload mt_relation --steps 16 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/sanitize_relation_$tool.json 2> gpurun_out/sanitize_relation_$tool.err
  tail -5 gpurun_out/sanitize_relation_$tool.log
done
