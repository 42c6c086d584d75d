#!/bin/bash
# round 2, first check of the row-walk rasterize backward: parity tests + A/B timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -x > gpurun_out/pytest_r2a.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r2a.log
tail -5 gpurun_out/pytest_r2a.log
for v in rows warp; do
  DIBR_B200_RASTER_BWD=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_r2a_$v.json 2> gpurun_out/bench_r2a_$v.err; echo "bench $v exit $?"; tail -2 gpurun_out/bench_r2a_$v.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2a_$v.json"))
print("$v value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), " e2e", round(d["e2e"]["value"]), "ms", round(d["e2e"]["ms_per_step"], 3))
print("phases", {k: round(v, 3) for k, v in d["roofline"]["phases"].items()})
print("raster_bwd ms", round(d["roofline"]["ms_per_launch"], 4), "GB/s", round(d["roofline"]["achieved"]), "frac", round(d["roofline"]["frac"], 3))
PY
done
