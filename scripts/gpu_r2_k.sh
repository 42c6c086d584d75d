#!/bin/bash
# round 2, call k: the float64 instantiation against the reference's <double> kernels, then the
# whole -m gpu suite and the default bench line (regression check after adding the f64 entry points).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_f64_gpu.py tests/test_reference_wrappers.py -m gpu -q --no-header -rf -s --timeout 200 > gpurun_out/pytest_r2k_f64.log 2>&1; echo "f64 pytest exit $?" >> gpurun_out/pytest_r2k_f64.log
grep -E "f64|passed|failed|Error|assert" gpurun_out/pytest_r2k_f64.log | tail -30
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 300 -x > gpurun_out/pytest_r2k_all.log 2>&1; echo "all pytest exit $?" >> gpurun_out/pytest_r2k_all.log
tail -5 gpurun_out/pytest_r2k_all.log
timeout 600 python bench.py > gpurun_out/bench_r2k.json 2> gpurun_out/bench_r2k.err; echo "bench exit $?"; tail -2 gpurun_out/bench_r2k.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r2k.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), " e2e ms", round(d["e2e"]["ms_per_step"], 3), "frac", d["roofline"]["frac"])
PY
# f64 timing at the c4 shard shape (informational)
timeout 300 python - <<'PY'
import torch, time
from kaolin_b200 import synthetic
from kaolin_b200.render.mesh import dibr_rasterization
fvz, fvi, fnz = synthetic.icosphere_views(4, 5, seed=1)
ff = synthetic.random_features(4, fvz.shape[1], 3, seed=2)
D = lambda a: torch.from_numpy(a).cuda().double()
a = [D(fvz), D(fvi).requires_grad_(True), D(ff).requires_grad_(True), D(fnz)]
for H in (256, 1024):
    for it in range(3):
        torch.cuda.synchronize(); t = time.time()
        f, s, i = dibr_rasterization(H, H, *a)
        (f.sum() + s.sum()).backward()
        torch.cuda.synchronize(); dt = time.time() - t
    print(f"f64 fwd+bwd 4x{H}x{H}, {fvz.shape[1]} faces: {dt*1e3:.2f} ms")
PY
