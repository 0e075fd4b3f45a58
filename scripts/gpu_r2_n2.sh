#!/bin/bash
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" PYTHONFAULTHANDLER=1 NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/n2_$tag.json 2> gpurun_out/n2_$tag.log
  echo "$tag rc=$? json_bytes=$(wc -c < gpurun_out/n2_$tag.json)"
  grep -v "Warning\|warn\|run_backward\|_set_allocator" gpurun_out/n2_$tag.log | tail -25
}
run thread MMT_TEACHER_THREAD=1
run nothread MMT_TEACHER_THREAD=0
