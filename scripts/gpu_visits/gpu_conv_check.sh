#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "conv" --timeout 300 2>&1 | tail -15
timeout 600 python scripts/conv_bench.py --all > gpurun_out/conv_bench.log 2>&1
cut -c1-175 gpurun_out/conv_bench.log | tail -50
