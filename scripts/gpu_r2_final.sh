#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1; tail -3 gpurun_out/z_smoke.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
