#!/bin/bash
# what the driver runs at round end, on the final commit
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > gpurun_out/final_smoke.log; cat gpurun_out/final_smoke.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/final_pytest.log; tail -1 gpurun_out/final_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.log
python - <<EOF
import json
j=json.load(open("gpurun_out/final_bench.json")); r=json.load(open("gpurun_out/final_bench_reference.json"))
print("ours value %.2f (%.2f ms) e2e %.2f (%.2f ms) frac %.3f traffic %s launches %d clocks %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["ms_per_step"], j["roofline"]["frac"], j["roofline"]["traffic"], j["gpu_launches"], j["clocks"]))
print("reference value %.3f img/s on %d cores; e2e ratio %.0fx" % (r["value"], r["cpu_baseline"]["cores"], j["e2e"]["value"] / r["value"]))
EOF
