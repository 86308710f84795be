#!/bin/bash
# round-2 GPU visit A: GPU test tier, then the N=1 bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc > gpurun_out/r2a_nproc.txt
timeout 1500 python tests/dist_selftest.py > gpurun_out/r2a_dist_selftest.txt 2>&1; echo "selftest rc=$?" >> gpurun_out/r2a_dist_selftest.txt
timeout 2400 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_dist.py > gpurun_out/r2a_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest_gpu.txt
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?" >> gpurun_out/r2a_bench.err
tail -3 gpurun_out/r2a_dist_selftest.txt; tail -5 gpurun_out/r2a_pytest_gpu.txt; tail -c 600 gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err
