#!/bin/bash
# stream-priority / SM-reservation / GIL switch-interval A/B on the final build
mkdir -p gpurun_out
uptime
timeout 600 python scripts/ab_bench.py --steps 30 --configs hiprio_tcap132,hiprio_tcap116,hiprio_tcap100,hiprio_tcap132_sw 2>&1 | grep -v Warning | tail -14 | tee gpurun_out/t_ab2.log
uptime
