#!/bin/bash
# round 2, call l (1 GPU): float64 + empty-mesh + peer-push tests first (fast signal), then the whole suite.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_f64_gpu.py tests/test_peer_push_gpu.py "tests/test_parity_gpu.py::test_empty_mesh" "tests/test_parity_gpu.py::test_errors_like_reference" -m gpu -q --no-header -rf -s --timeout 200 > gpurun_out/pytest_r2l_new.log 2>&1; echo "new pytest exit $?" >> gpurun_out/pytest_r2l_new.log
grep -E "f64|peer|passed|failed|Error|assert" gpurun_out/pytest_r2l_new.log | tail -30
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 300 > gpurun_out/pytest_r2l_all.log 2>&1; echo "all pytest exit $?" >> gpurun_out/pytest_r2l_all.log
tail -8 gpurun_out/pytest_r2l_all.log
