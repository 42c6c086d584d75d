#!/bin/bash
mkdir -p gpurun_out
DIBR_B200_LIB=$PWD/kaolin_b200/csrc/libdibr_b200_collect.so timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_variants_gpu.py -m gpu -q --no-header -rf --timeout 240 > gpurun_out/pytest_r2j.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r2j.log
tail -4 gpurun_out/pytest_r2j.log
run() {
  tag=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-e2e-images $BARGS > gpurun_out/bench_r2j_$tag.json 2> gpurun_out/bench_r2j_$tag.err; echo "bench $tag exit $?"; tail -2 gpurun_out/bench_r2j_$tag.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r2j_$tag.json").read().strip().splitlines()[-1])
print("$tag value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), " e2e ms", round(d["e2e"]["ms_per_step"], 3))
print("   ", [(k["kernel"], k["ms"]) for k in d["roofline"]["kernels"] if k["ms"] > 0.01])
PY
}
BARGS="" run base
BARGS="" run collect DIBR_B200_LIB=$PWD/kaolin_b200/csrc/libdibr_b200_collect.so
BARGS="--workload c5" run c5collect DIBR_B200_LIB=$PWD/kaolin_b200/csrc/libdibr_b200_collect.so
