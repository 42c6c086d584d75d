#!/bin/bash
# compute-sanitizer memcheck over small parity scenes (default paths + the kept variants) and the new ops
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 \
  python -m pytest tests/test_variants_gpu.py tests/test_deftet_gpu.py tests/test_pipeline_gpu.py -m gpu -q --no-header -x --timeout 600 \
  -k "ico4_256 or soup2000 or golden or truncation or wide_faces or more_hits or texture or mask_iou or prepare_vertices_vs" \
  > gpurun_out/sanitize_r2.log 2>&1; echo "memcheck exit $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" gpurun_out/sanitize_r2.log | tail -12
