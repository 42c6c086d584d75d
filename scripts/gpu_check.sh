#!/bin/bash
# usage: gpu_check.sh <tag> [tests] [bench] [launches] [full:<kernel regex>]
tag=$1; shift
mkdir -p gpurun_out
for what in "$@"; do
  case $what in
    tests)
      timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -s > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
      grep -E "rel err|passed|failed|Error|error" gpurun_out/pytest_$tag.log | tail -30
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ;;
    bench)
      timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?"; tail -2 gpurun_out/bench_$tag.err
      python - <<PY
import json
d = json.load(open("gpurun_out/bench_$tag.json"))
print("value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), " e2e", round(d["e2e"]["value"]), "ms", round(d["e2e"]["ms_per_step"], 3))
print("phases", {k: round(v, 3) for k, v in d["roofline"]["phases"].items()})
print("raster_bwd ms", round(d["roofline"]["ms_per_launch"], 4), "GB/s", round(d["roofline"]["achieved"]), "frac", round(d["roofline"]["frac"], 3))
print("clocks", d["clocks"], "ref_cuda", d.get("reference_cuda"), "cpu", d.get("cpu_baseline"))
PY
      ;;
    benchfast)
      timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?"; tail -2 gpurun_out/bench_$tag.err
      python - <<PY
import json
d = json.load(open("gpurun_out/bench_$tag.json"))
print("value", round(d["value"]), "Mpx/s  ms/step", round(d["ms_per_step"], 3), " e2e", round(d["e2e"]["value"]), "ms", round(d["e2e"]["ms_per_step"], 3))
print("phases", {k: round(v, 3) for k, v in d["roofline"]["phases"].items()})
print("raster_bwd ms", round(d["roofline"]["ms_per_launch"], 4), "GB/s", round(d["roofline"]["achieved"]), "frac", round(d["roofline"]["frac"], 3))
PY
      ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_launch_$tag.log 2>&1; echo "ncu launches exit $?" ;;
    full:*)
      rx=${what#full:}
      cp kaolin_b200/csrc/libdibr_b200.so gpurun_out/lib_$tag.so
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s ${NCU_SKIP:-3} -c ${NCU_COUNT:-4} -o gpurun_out/prof_$tag -f python bench.py ${BENCH_ARGS:-} --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full exit $?" ;;
  esac
done
