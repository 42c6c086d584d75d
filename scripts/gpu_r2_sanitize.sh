#!/bin/bash
# forward fuzz test (plain), then compute-sanitizer memcheck over small parity scenes and the new ops
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_variants_gpu.py tests/test_multi_gpu_gpu.py -m gpu -q --no-header -rf --timeout 300 > gpurun_out/pytest_r2k.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r2k.log
tail -5 gpurun_out/pytest_r2k.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 \
  python -m pytest tests/test_variants_gpu.py tests/test_deftet_gpu.py tests/test_pipeline_gpu.py -m gpu -q --no-header -x --timeout 600 \
  -k "(test_variant_equals_oracle and ico4_256) or golden or truncation or wide_faces or texture_mapping_vs or mask_iou or prepare_vertices_vs" \
  > gpurun_out/sanitize_r2.log 2>&1; echo "memcheck exit $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" gpurun_out/sanitize_r2.log | tail -12
