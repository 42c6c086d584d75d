#!/bin/bash
# round 2: soft-mask kernels A/B (runs vs dense backward, enum occupancy), strip 16, failing tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py tests/test_reference_wrappers.py tests/test_parity_gpu.py -m gpu -q --no-header -rf --timeout 240 > gpurun_out/pytest_r2e.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r2e.log
tail -6 gpurun_out/pytest_r2e.log
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-e2e-images > gpurun_out/bench_r2e_$tag.json 2> gpurun_out/bench_r2e_$tag.err; echo "bench $tag exit $?"; tail -2 gpurun_out/bench_r2e_$tag.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2e_$tag.json"))
print("$tag value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), " e2e ms", round(d["e2e"]["ms_per_step"], 3), "raster_bwd", round(d["roofline"]["ms_per_launch"], 4), "frac", round(d["roofline"]["frac"], 3))
print("   ", [(k["kernel"], k["ms"]) for k in d["roofline"]["kernels"] if k["ms"] > 0.01])
PY
}
run base
run dense DIBR_B200_SOFT_BWD=dense
run strip16 DIBR_B200_ROWS_STRIP=16
run enum5 DIBR_B200_LIB=$PWD/kaolin_b200/csrc/libdibr_b200_enum5.so
run enum6 DIBR_B200_LIB=$PWD/kaolin_b200/csrc/libdibr_b200_enum6.so
