#!/bin/bash
# round 2: prefetch variant of the rows kernel, graphed fast path
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --no-header -rf --timeout 240 -k "graphed or fused_dibr_vs_reference" > gpurun_out/pytest_r2i.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r2i.log
tail -5 gpurun_out/pytest_r2i.log
run() {  # tag, env..., -- args
  tag=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-e2e-images $BARGS > gpurun_out/bench_r2i_$tag.json 2> gpurun_out/bench_r2i_$tag.err; echo "bench $tag exit $?"; tail -2 gpurun_out/bench_r2i_$tag.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r2i_$tag.json").read().strip().splitlines()[-1])
print("$tag value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), " e2e ms", round(d["e2e"]["ms_per_step"], 3), "raster_bwd", round(d["roofline"]["ms_per_launch"], 4), "frac", round(d["roofline"]["frac"], 3), "graphed", d.get("e2e_graphed"))
print("   ", [(k["kernel"], k["ms"]) for k in d["roofline"]["kernels"] if k["ms"] > 0.01])
PY
}
BARGS="" run base
BARGS="" run pf DIBR_B200_LIB=$PWD/kaolin_b200/csrc/libdibr_b200_pf.so
BARGS="--workload c2 --e2e-graph" run c2graph
BARGS="--workload c3 --e2e-graph" run c3graph
