#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -8 > gpurun_out/s_pytest.log; tail -2 gpurun_out/s_pytest.log
timeout 600 python scripts/ab_bench.py --steps 30 --configs base,noshare 2>&1 | grep visit
