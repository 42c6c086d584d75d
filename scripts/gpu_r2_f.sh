#!/bin/bash
# round 2: soft_eval run kernel A/B + occupancy variants; parity subset
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rf --timeout 240 > gpurun_out/pytest_r2f.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r2f.log
tail -6 gpurun_out/pytest_r2f.log
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-e2e-images > gpurun_out/bench_r2f_$tag.json 2> gpurun_out/bench_r2f_$tag.err; echo "bench $tag exit $?"; tail -2 gpurun_out/bench_r2f_$tag.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2f_$tag.json"))
print("$tag value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), " e2e ms", round(d["e2e"]["ms_per_step"], 3), "raster_bwd", round(d["roofline"]["ms_per_launch"], 4), "frac", round(d["roofline"]["frac"], 3))
print("   ", [(k["kernel"], k["ms"]) for k in d["roofline"]["kernels"] if k["ms"] > 0.01])
PY
}
run base
run pairs DIBR_B200_SOFT_EVAL=pairs
run occ DIBR_B200_LIB=$PWD/kaolin_b200/csrc/libdibr_b200_occ.so
for w in c2 c3 c5; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-e2e-images > gpurun_out/bench_r2f_$w.json 2> gpurun_out/bench_r2f_$w.err; echo "bench $w exit $?"
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2f_$w.json"))
print("$w value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), " e2e ms", round(d["e2e"]["ms_per_step"], 3))
print("   ", [(k["kernel"], k["ms"]) for k in d["roofline"]["kernels"] if k["ms"] > 0.01])
PY
done
