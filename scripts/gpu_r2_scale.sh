#!/bin/bash
# usage: bash scripts/gpu_r2_scale.sh N
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.log
echo "rc=$?"
python - <<EOF
import json
for l in open("gpurun_out/scale_n$N.json"):
    if l.startswith("{"):
        j=json.loads(l); print("N=$N", round(j["value"],2), "img/s", round(j["ms_per_step"],2), "ms | e2e", round(j["e2e"]["value"],2), round(j["e2e"]["ms_per_step"],2), j["step_ms"]["dev"], "replicas_identical", j.get("replicas_identical"), j["allocator"])
EOF
grep -i -E "error|traceback" gpurun_out/scale_n$N.log | head -5
