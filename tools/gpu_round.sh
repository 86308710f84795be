#!/bin/bash
# One GPU-box visit: GPU test tier, the tiled-SpMV probe, the driver bench line, every-config table.
# Usage (from the repo root on the box): bash tools/gpu_round.sh <tag> [steps...]
tag=${1:-rX}; shift
steps=${@:-"pytest probe bench configs refs"}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.txt 2>&1
nproc >> gpurun_out/${tag}_gpus.txt
for s in $steps; do
  case $s in
    pytest) timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_gpu.txt ;;
    probe) timeout 300 tools/probes/_bin/tiled_probe > gpurun_out/${tag}_tiled_probe.txt 2>&1; echo "rc=$?" >> gpurun_out/${tag}_tiled_probe.txt ;;
    bench) timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "rc=$?" >> gpurun_out/${tag}_bench.err ;;
    benchref) timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err ;;
    configs) timeout 1200 python tools/bench_configs.py --configs 1,1lp,3,4,5s,5,2s --no-ref > gpurun_out/${tag}_configs_b200.jsonl 2> gpurun_out/${tag}_configs_b200.err ;;
    refs) timeout 900 python tools/bench_configs.py --configs 1,5s,3 --ref-batch-sample 16 --ref-time-limit 60 2> gpurun_out/${tag}_configs_ref.err | grep reference > gpurun_out/${tag}_configs_ref.jsonl ;;
    batchperf) timeout 600 python tests/dev_batch_check.py --perf-only --perf > gpurun_out/${tag}_batch_perf.txt 2>&1 ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --scale 0.25 --no-cpu-baseline --no-time-to-eps > gpurun_out/${tag}_launches_bench.log 2>&1 ;;
  esac
done
ls -la gpurun_out > /dev/null
