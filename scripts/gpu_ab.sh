#!/bin/bash
# A/B of bench.py flag sets on one box: gpu_ab.sh "<flags A>" "<flags B>" ...
mkdir -p gpurun_out
i=0
for f in "$@"; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda $f > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err || tail -3 gpurun_out/ab_$i.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/ab_$i.json").read().strip().splitlines()[-1])
print("[$f]", "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 4), {k: round(v, 3) for k, v in d["roofline"]["phases"].items() if k.endswith("_ms")})
PY
  i=$((i+1))
done
