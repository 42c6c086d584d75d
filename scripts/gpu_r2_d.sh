#!/bin/bash
# round 2: all GPU tests (incl. pipeline / wrappers / fp64), bench with rows-kernel strip variants, launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf --timeout 240 > gpurun_out/pytest_r2d.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r2d.log
tail -12 gpurun_out/pytest_r2d.log
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-e2e-images > gpurun_out/bench_r2d_$tag.json 2> gpurun_out/bench_r2d_$tag.err; echo "bench $tag exit $?"; tail -2 gpurun_out/bench_r2d_$tag.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2d_$tag.json"))
print("$tag value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), " e2e ms", round(d["e2e"]["ms_per_step"], 3), "raster_bwd", round(d["roofline"]["ms_per_launch"], 4), "frac", round(d["roofline"]["frac"], 3))
print("   ", [(k["kernel"], k["ms"]) for k in d["roofline"]["kernels"] if k["ms"] > 0.01])
PY
}
run strip64 DIBR_B200_ROWS_STRIP=64
run strip128 DIBR_B200_ROWS_STRIP=128
run strip32 DIBR_B200_ROWS_STRIP=32
