#!/bin/bash
# quick perf loop: launch lists of the given workloads (default c4_shard c5), no tests
mkdir -p gpurun_out
for w in ${@:-c4_shard c5}; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${w}_launches.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda > /dev/null 2>&1
echo "== $w"; python scripts/launch_summary.py gpurun_out/${w}_launches.csv 2>/dev/null | grep -v "at::"
done
