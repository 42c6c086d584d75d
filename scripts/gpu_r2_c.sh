#!/bin/bash
# round 2: after the bit-matrix overflow fix: parity (per-test timeout), bench, ncu of the v2 forward kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -x --timeout 240 > gpurun_out/pytest_r2c.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r2c.log
tail -6 gpurun_out/pytest_r2c.log
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-e2e-images > gpurun_out/bench_r2c_$tag.json 2> gpurun_out/bench_r2c_$tag.err; echo "bench $tag exit $?"; tail -2 gpurun_out/bench_r2c_$tag.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2c_$tag.json"))
print("$tag value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), " e2e ms", round(d["e2e"]["ms_per_step"], 3), "raster_bwd", round(d["roofline"]["ms_per_launch"], 4), "frac", round(d["roofline"]["frac"], 3))
print("   ", [(k["kernel"], k["ms"]) for k in d["roofline"]["kernels"] if k["ms"] > 0.02])
PY
}
run s2 DIBR_B200_FWD=s2
run s1 DIBR_B200_FWD=s1
NCU_SKIP=1 NCU_COUNT=2 timeout 600 bash scripts/gpu_check.sh r2fwd2 full:dibr_fwd2
