#!/bin/bash
# parity at BASELINE config sizes + bench lines for the other configs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_parity_gpu.py -m gpu -q --no-header -rf -s -k "baseline_configs" > gpurun_out/pytest_configs.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_configs.log
grep -E "face_idx exact|passed|failed|Error" gpurun_out/pytest_configs.log | tail -12
for w in c2 c3 c5; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --ref-cuda-views 1 > gpurun_out/bench_cfg_$w.json 2> gpurun_out/bench_cfg_$w.err; echo "bench $w exit $?"; tail -1 gpurun_out/bench_cfg_$w.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_cfg_$w.json").read().strip().splitlines()[-1])
    print("$w", "value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "fwd/bwd ms", round(d["roofline"]["phases"]["forward_ms"],3), round(d["roofline"]["phases"]["backward_ms"],3), "ref_cuda", d.get("reference_cuda"))
except Exception as e:
    print("$w failed", e)
PY
done
