#!/bin/bash
# usage: gpu_multi.sh <tag> <ngpus>   (N-GPU bench with 1 and 4 view-chunks, then N=1 on the same box)
tag=$1; n=$2
mkdir -p gpurun_out
for ch in ${CHUNK_LIST:-1}; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 --chunks $ch --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_${tag}_n${n}_c$ch.json 2> gpurun_out/bench_${tag}_n${n}_c$ch.err; echo "bench n=$n chunks=$ch exit $?"; tail -3 gpurun_out/bench_${tag}_n${n}_c$ch.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${tag}_n${n}_c$ch.json").read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "chunks $ch value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), " e2e", round(d["e2e"]["value"]), "ms", round(d["e2e"]["ms_per_step"], 3), "phases", {k: round(v, 3) for k, v in d["roofline"]["phases"].items() if k.endswith("_ms")})
PY
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err; echo "bench n=1 exit $?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${tag}_n1.json").read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), " e2e", round(d["e2e"]["value"]), "ms", round(d["e2e"]["ms_per_step"], 3), d["e2e"].get("host_wall_ms_per_step"), "phases", {k: round(v, 3) for k, v in d["roofline"]["phases"].items() if k.endswith("_ms")})
PY
