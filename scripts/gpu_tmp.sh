#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 600 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline --no-roofline 2>/dev/null | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('mt', b['ms_per_step'], b['e2e']['ms_per_step'])"
