#!/bin/bash
# N = 2 on the final build: the data-parallel path (bucketed all-reduce, pinned touched-flag upload, replicas_identical)
mkdir -p gpurun_out
PYTHONFAULTHANDLER=1 NCCL_DEBUG=WARN timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-roofline > gpurun_out/n2b.json 2> gpurun_out/n2b.log
echo "rc=$? json_bytes=$(wc -c < gpurun_out/n2b.json)"
grep -v "Warning\|warn\|run_backward\|_set_allocator" gpurun_out/n2b.log | tail -8
python -c "
import json; j=json.load(open('gpurun_out/n2b.json')); print('N=2 value %.2f (%.2f ms) e2e %.2f (%.2f ms) replicas_identical %s' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['ms_per_step'], j.get('replicas_identical')), j['step_ms'])"
