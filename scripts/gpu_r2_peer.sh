#!/bin/bash
# N GPUs: the gradient all-gather by stores into peer memory vs NCCL - parity (N=2 test) and bench per transport.
N=${1:-2}
GATHERS=${GATHERS:-"nccl peer peer_sm"}
mkdir -p gpurun_out
if [ "${SKIP_TEST:-0}" != "1" ]; then
  timeout 600 python -m pytest tests/test_multi_gpu_gpu.py -m gpu -q --no-header -rf --timeout 400 -s > gpurun_out/pytest_peer_n$N.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_peer_n$N.log
  grep -E "MGPU_RESULT|passed|failed|skipped|exit|unavailable" gpurun_out/pytest_peer_n$N.log | cut -c1-1500 | tail -6
fi
for g in $GATHERS; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --gather $g --no-cpu-baseline --no-ref-cuda --no-e2e-images $BARGS > gpurun_out/bench_peer_n${N}_$g.json 2> gpurun_out/bench_peer_n${N}_$g.err; echo "bench N=$N gather=$g exit $?"; grep -E "transport|Error|error" gpurun_out/bench_peer_n${N}_$g.err | tail -3 | cut -c1-300
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_peer_n${N}_$g.json").read().strip().splitlines()[-1])
    print("N=$N gather=$g (", d["config"].get("gather_transport"), ") value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), " e2e ms", round(d["e2e"]["ms_per_step"], 3), "step_ms", {k: round(v, 3) for k, v in d["step_ms"].items() if isinstance(v, float)})
    print("   phases", {k: round(v, 3) for k, v in d["roofline"]["phases"].items() if k.endswith("_ms")})
except Exception as e:
    print("no line:", e)
PY
done
[ "${SKIP_N1:-0}" = "1" ] && exit 0
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-e2e-images > gpurun_out/bench_peer_n1_samebox_as_n$N.json 2> /dev/null
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_peer_n1_samebox_as_n$N.json").read().strip().splitlines()[-1])
print("N=1 (same box) value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), " e2e ms", round(d["e2e"]["ms_per_step"], 3))
PY
