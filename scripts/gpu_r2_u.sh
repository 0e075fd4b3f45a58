#!/bin/bash
# where the end-to-end loop's extra 1.7 ms per step goes: host phases of the e2e step (train_step | fresh targets | stage | loss wait)
mkdir -p gpurun_out
for st in 1; do
  MMT_BENCH_STAGE=$st MMT_BENCH_TRACE=2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/u_bench_$st.json 2> gpurun_out/u_bench_$st.log
  grep "host phases" gpurun_out/u_bench_$st.log
  python -c "
import json; j=json.load(open('gpurun_out/u_bench_$st.json')); print('stage=$st dev %.2f e2e %.2f' % (j['ms_per_step'], j['e2e']['ms_per_step']), j['step_ms'])"
done
