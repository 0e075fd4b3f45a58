#!/bin/bash
# Regenerates the round's measured evidence from the SHIPPED build (run under gpurun, 1 GPU):
#   bash scripts/refresh_profiles.sh <tag> [sections]      sections: bench ncu full sanitize (default: all)
# Outputs land in gpurun_out/refresh_<tag>/; `python scripts/summarize_ncu.py <tag>` (run in the build container, where
# the .ncu-rep files can be read) then writes the tracked summaries under profiles/.
TAG=${1:-r02}
SECTIONS=${2:-bench ncu full sanitize}
OUT=gpurun_out/refresh_$TAG
mkdir -p $OUT
python scripts/source_hash.py > $OUT/source_hash.txt
for S in $SECTIONS; do case $S in
bench)
  timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.log
  timeout 300 python bench.py --workload sup --no-cpu-baseline > $OUT/bench_sup.json 2> $OUT/bench_sup.log
  timeout 300 python bench.py --workload mt_relation --no-cpu-baseline > $OUT/bench_relation.json 2> $OUT/bench_relation.log
  MMT_BENCH_TRACE=1 timeout 300 python bench.py --steps 40 --no-cpu-baseline --no-roofline > $OUT/trace.json 2> $OUT/trace.log
  ;;
ncu)
  # every launch of one steady-state iteration (graphs on, as shipped): duration, DRAM bytes, tensor-pipe activity
  timeout 2400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
      --clock-control none -s ${NCU_SKIP:-4400} -c ${NCU_COUNT:-1600} --csv --log-file $OUT/launches.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > $OUT/ncu_bench.log 2>&1
  ;;
full)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 600 -c 6 -f -o $OUT/prof_conv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --no-graphs > $OUT/ncu_full_conv.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"roi_align_fpn|sgd_ema|mgd_level|psm_kernel|rpn_topk|balanced_sample|greedy_scan" -s 8 -c 14 -f -o $OUT/prof_bw \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --no-graphs > $OUT/ncu_full_bw.log 2>&1
  ;;
sanitize)
  timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py -m gpu -q -x \
      -k "nms or roi_align or psm or mgd or conv_fprop or cta_pair or fused_backward_ops or rasterize or paste or topk or balanced_sample or padded_heads" > $OUT/sanitizer_memcheck.log 2>&1
  echo "memcheck exit code $?" >> $OUT/sanitizer_memcheck.log
  timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py -m gpu -q -x \
      -k "nms_golden or roi_align_forward_golden or roi_align_fpn or psm or mgd or sigmoid_focal or topk or balanced_sample" > $OUT/sanitizer_racecheck.log 2>&1
  echo "racecheck exit code $?" >> $OUT/sanitizer_racecheck.log
  ;;
esac; done
ls -la $OUT
