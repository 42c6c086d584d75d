#!/bin/bash
# round 2: forward v2 kernel: parity + A/B timing, ncu of the row-walk backward
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -x > gpurun_out/pytest_r2b.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r2b.log
tail -4 gpurun_out/pytest_r2b.log
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-e2e-images > gpurun_out/bench_r2b_$tag.json 2> gpurun_out/bench_r2b_$tag.err; echo "bench $tag exit $?"; tail -2 gpurun_out/bench_r2b_$tag.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2b_$tag.json"))
print("$tag value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), " e2e ms", round(d["e2e"]["ms_per_step"], 3))
print("   ", [(k["kernel"], k["ms"]) for k in d["roofline"]["kernels"] if k["ms"] > 0.02])
PY
}
run old DIBR_B200_FWD=old
run s1 DIBR_B200_FWD=s1
run s2 DIBR_B200_FWD=s2
run s2minb5 DIBR_B200_FWD=s2 DIBR_B200_LIB=$PWD/kaolin_b200/csrc/libdibr_b200_minb5.so
for w in c2 c3 c5; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-e2e-images > gpurun_out/bench_r2b_$w.json 2> gpurun_out/bench_r2b_$w.err; echo "bench $w exit $?"
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2b_$w.json"))
print("$w value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), " e2e ms", round(d["e2e"]["ms_per_step"], 3))
print("   ", [(k["kernel"], k["ms"]) for k in d["roofline"]["kernels"] if k["ms"] > 0.02])
PY
done
NCU_SKIP=2 NCU_COUNT=2 bash scripts/gpu_check.sh r2rows full:raster_bwd_rows
