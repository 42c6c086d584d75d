#!/bin/bash
# round 2 evidence for profiles/: launch list + one ncu --set full capture of a whole step
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda --no-e2e-images > gpurun_out/ncu_launch_r2.log 2>&1; echo "ncu launches exit $?"
cp kaolin_b200/csrc/libdibr_b200.so gpurun_out/lib_r2final.so
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"dibr_|soft_|raster_|bin_faces|scan_bins" -s 11 -c 11 -o gpurun_out/prof_r2final -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda --no-e2e-images > gpurun_out/ncu_full_r2final.log 2>&1; echo "ncu full exit $?"
