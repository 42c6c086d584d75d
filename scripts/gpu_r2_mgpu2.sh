#!/bin/bash
# N GPUs: sharded parity (incl. pipelined backward) + bench with 1 and 2 backward chunks
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu_gpu.py tests/test_parity_gpu.py -m gpu -q --no-header -rf --timeout 300 -k "sharded or view_chunked" -s > gpurun_out/pytest_mgpu2_n$N.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mgpu2_n$N.log
grep -E "passed|failed|skipped|exit" gpurun_out/pytest_mgpu2_n$N.log | tail -4
for c in ${CHUNKS:-1 2}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --bwd-chunks $c > gpurun_out/bench_r2_n${N}_bc$c.json 2> gpurun_out/bench_r2_n${N}_bc$c.err; echo "bench N=$N bwd-chunks=$c exit $?"; tail -2 gpurun_out/bench_r2_n${N}_bc$c.err | cut -c1-300
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r2_n${N}_bc$c.json").read().strip().splitlines()[-1])
print("N=$N bwd-chunks=$c value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), " e2e ms", round(d["e2e"]["ms_per_step"], 3), "step_ms", {k: round(v, 3) for k, v in d["step_ms"].items() if isinstance(v, float)})
print("   phases", {k: round(v, 3) for k, v in d["roofline"]["phases"].items() if k.endswith("_ms")})
PY
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-e2e-images > gpurun_out/bench_r2_n1_samebox.json 2> /dev/null
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r2_n1_samebox.json").read().strip().splitlines()[-1])
print("N=1 (same box) value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), " e2e ms", round(d["e2e"]["ms_per_step"], 3))
PY
