#!/bin/bash
# round 2: full GPU suite (deftet, bf16 rows, pipeline, binding) + bench lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf --timeout 240 > gpurun_out/pytest_r2g.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r2g.log
tail -8 gpurun_out/pytest_r2g.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
run() {  # tag, args...
  tag=$1; shift
  timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-e2e-images "$@" > gpurun_out/bench_r2g_$tag.json 2> gpurun_out/bench_r2g_$tag.err; echo "bench $tag exit $?"; tail -2 gpurun_out/bench_r2g_$tag.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r2g_$tag.json").read().strip().splitlines()[-1])
print("$tag value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), " e2e ms", round(d["e2e"]["ms_per_step"], 3), "raster_bwd", round(d["roofline"]["ms_per_launch"], 4), "frac", round(d["roofline"]["frac"], 3))
print("   ", [(k["kernel"], k["ms"]) for k in d["roofline"]["kernels"] if k["ms"] > 0.01])
PY
}
run c4
run c4bf16 --features bf16
run c2 --workload c2
run c3 --workload c3
run c5 --workload c5
