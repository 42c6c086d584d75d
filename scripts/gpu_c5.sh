#!/bin/bash
# c5 (1.3M triangles, 2048^2), smooth and spiky: parity vs reference CUDA + bench + launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -q --no-header -rf -s -k "${C5_TESTS:-windowed or c5_ or overflow}" > gpurun_out/pytest_c5.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_c5.log
grep -E "face_idx exact|passed|failed|Error|exit" gpurun_out/pytest_c5.log | tail -12
for w in c5 c5_spiky; do
timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --ref-cuda-views 1 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench exit $?"; tail -2 gpurun_out/bench_$w.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$w.json").read().strip().splitlines()[-1])
print("$w value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "phases", {k: round(v, 3) for k, v in d["roofline"]["phases"].items()}, "ref_cuda", d.get("reference_cuda"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${w}_launches.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/${w}_launches.csv 2>/dev/null | head -20
done
