#!/bin/bash
# round 2, last call: the whole -m gpu suite, smoke(), the default bench line.
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -q --no-header -rf --timeout 120 -x > gpurun_out/pytest_r2_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r2_final.log
tail -4 gpurun_out/pytest_r2_final.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 120 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; echo "bench exit $?"
tail -c 600 gpurun_out/bench_r2_final.json | head -c 600; echo
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r2_final.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3), "cpu", d["cpu_baseline"]["value"])
PY
