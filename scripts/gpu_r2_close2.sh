#!/bin/bash
# round-end sequence on the final commit (the CPU reference arm is unchanged since scripts/gpu_r2_close.sh ran it: not repeated)
mkdir -p gpurun_out
uptime
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > gpurun_out/final2_smoke.log; cat gpurun_out/final2_smoke.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/final2_pytest.log; tail -3 gpurun_out/final2_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.log
python - <<PY
import json
j=json.load(open("gpurun_out/final2_bench.json"))
print("ours value %.2f (%.2f ms) e2e %.2f (%.2f ms) frac %.3f traffic %s launches %d clocks %s step_ms %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["ms_per_step"], j["roofline"]["frac"], j["roofline"]["traffic"], j["gpu_launches"], j["clocks"], j["step_ms"]))
PY
uptime
