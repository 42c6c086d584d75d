#!/bin/bash
# round 2, N GPUs (run with gpurun --gpus N): sharded == single-GPU parity test + scaling bench line
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu_gpu.py -q --no-header -rf -x -s > gpurun_out/pytest_mgpu_n$N.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mgpu_n$N.log
grep -E "MGPU_RESULT|passed|failed|skipped|exit" gpurun_out/pytest_mgpu_n$N.log | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err; echo "bench N=$N exit $?"; tail -3 gpurun_out/bench_r2_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r2_n$N.json").read().strip().splitlines()[-1])
print("N=$N value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), " e2e ms", round(d["e2e"]["ms_per_step"], 3))
print("step_ms", d["step_ms"])
PY
if [ "$N" != "1" ]; then
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-e2e-images > gpurun_out/bench_r2_n1_samebox.json 2> /dev/null
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2_n1_samebox.json"))
print("N=1 (same box) value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), " e2e ms", round(d["e2e"]["ms_per_step"], 3))
PY
fi
