#!/bin/bash
# usage: gpu_multi.sh <tag> <ngpus>
tag=$1; n=$2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_${tag}_n$n.json 2> gpurun_out/bench_${tag}_n$n.err; echo "bench n=$n exit $?"; tail -3 gpurun_out/bench_${tag}_n$n.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${tag}_n$n.json").read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), " e2e", round(d["e2e"]["value"]), "ms", round(d["e2e"]["ms_per_step"], 3))
PY
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err; echo "bench n=1 exit $?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${tag}_n1.json").read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), " e2e", round(d["e2e"]["value"]), "ms", round(d["e2e"]["ms_per_step"], 3), d["e2e"].get("host_wall_ms_per_step"))
PY
