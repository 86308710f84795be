#!/bin/bash
# one GPU-box visit; usage: bash tools/visit.sh <tag> <steps...>
tag=$1; shift
mkdir -p gpurun_out
t0=$(date +%s)
for s in "$@"; do
  case $s in
    tiledtests) timeout 600 python -m pytest tests/test_gpu_tiled.py -x -q --durations=10 > gpurun_out/${tag}_pytest_tiled.txt 2>&1; echo "rc=$? t=$(( $(date +%s)-t0 ))" >> gpurun_out/${tag}_pytest_tiled.txt ;;
    sanitize) SCS_B200_TILED=1 timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_tiled.py -x -q -k "kkt and (case0 or case4)" > gpurun_out/${tag}_sanitize.txt 2>&1; echo "rc=$?" >> gpurun_out/${tag}_sanitize.txt ;;
    pytest) timeout 900 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/${tag}_pytest_gpu.txt 2>&1; echo "rc=$? t=$(( $(date +%s)-t0 ))" >> gpurun_out/${tag}_pytest_gpu.txt ;;
    profile) SCS_B200_TILED_VERBOSE=1 timeout 300 python tools/tiled_profile.py > gpurun_out/${tag}_tiled_profile.txt 2>&1; cat gpurun_out/${tag}_tiled_profile.txt ;;
    ncug) timeout 400 ncu --set full --clock-control none --import-source on -k regex:tiled -s 28 -c 4 -f -o gpurun_out/${tag}_tiled_g python tools/tiled_profile.py > gpurun_out/${tag}_ncug.log 2>&1; tail -2 gpurun_out/${tag}_ncug.log ;;
    benchq) timeout 300 python bench.py --scale 0.25 --steps 4 --warmup 3 --no-cpu-baseline --no-time-to-eps > gpurun_out/${tag}_bench_quarter.json 2> gpurun_out/${tag}_bench_quarter.err; echo "rc=$? t=$(( $(date +%s)-t0 ))" >> gpurun_out/${tag}_bench_quarter.err ;;
    benchq0) SCS_B200_TILED=0 timeout 300 python bench.py --scale 0.25 --steps 4 --warmup 3 --no-cpu-baseline --no-time-to-eps > gpurun_out/${tag}_bench_quarter_rowengine.json 2> gpurun_out/${tag}_bench_quarter_rowengine.err ;;
    bench) timeout 500 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "rc=$? t=$(( $(date +%s)-t0 ))" >> gpurun_out/${tag}_bench.err ;;
    bench0) SCS_B200_TILED=0 timeout 500 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_rowengine.json 2> gpurun_out/${tag}_bench_rowengine.err ;;
    benchref) timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err ;;
    configs) timeout 400 python tools/bench_configs.py --configs 5s,1,1lp,3,4,2s --no-ref > gpurun_out/${tag}_configs_b200.jsonl 2> gpurun_out/${tag}_configs_b200.err ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --scale 0.25 --no-cpu-baseline --no-time-to-eps > gpurun_out/${tag}_launches_bench.log 2>&1 ;;
    ncufull) timeout 600 ncu --set full --clock-control none --import-source on -k regex:tiled_kernel -s 40 -c 4 -f -o gpurun_out/${tag}_tiled_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-time-to-eps > gpurun_out/${tag}_ncufull.log 2>&1 ;;
  esac
done
for f in gpurun_out/${tag}_pytest*.txt gpurun_out/${tag}_sanitize.txt; do [ -f $f ] && tail -15 $f; done
for f in gpurun_out/${tag}_bench*.json; do [ -f $f ] && cut -c1-250 $f; done
for f in gpurun_out/${tag}_bench*.err; do [ -f $f ] && tail -5 $f; done
true
