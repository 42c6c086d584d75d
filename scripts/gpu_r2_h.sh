#!/bin/bash
# round 2: variants test + the default bench line (cpu baseline, reference CUDA, e2e_images) + profiles
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_variants_gpu.py tests/test_parity_gpu.py -m gpu -q --no-header -rf --timeout 240 > gpurun_out/pytest_r2h.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r2h.log
tail -5 gpurun_out/pytest_r2h.log
timeout 900 python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo "bench default exit $?"; tail -2 gpurun_out/bench_r2_n1.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r2_n1.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3))
print("cpu", d["cpu_baseline"]); print("refcuda", d.get("reference_cuda")); print("e2e_images", d.get("e2e_images")); print("clocks", d["clocks"])
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err; echo "reference arm exit $?"; head -c 700 gpurun_out/bench_r2_reference.json; echo
bash scripts/gpu_r2_profile.sh
