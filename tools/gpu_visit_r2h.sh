#!/bin/bash
# round-2 GPU visit H (1 GPU): greedy vs plain dealing, adaptive chunk size on the small configs, fixed thread test, parity regression
mkdir -p gpurun_out
timeout 600 python tools/spmv_variants.py --reps 30 > gpurun_out/r2h_spmv_deal.txt 2>&1
timeout 300 python tools/bench_configs.py --configs 1,1lp,3 --no-ref > gpurun_out/r2h_configs.jsonl 2>/dev/null
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tiled.py tests/test_gpu_full_size.py -q > gpurun_out/r2h_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.txt
timeout 300 python -m pytest tests/test_gpu_boundary.py -q -k "two_threads" > gpurun_out/r2h_pytest_threads.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest_threads.txt
cut -c1-250 gpurun_out/r2h_spmv_deal.txt; cut -c1-330 gpurun_out/r2h_configs.jsonl; tail -4 gpurun_out/r2h_pytest.txt; tail -4 gpurun_out/r2h_pytest_threads.txt
