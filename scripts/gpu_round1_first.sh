mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -x --no-header -rf > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --workload tiny --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err; echo "bench tiny exit $?"; tail -3 gpurun_out/bench_tiny.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 exit $?"; tail -3 gpurun_out/bench_c4.err; cat gpurun_out/bench_c4.json | cut -c1-3000
