#!/bin/bash
# round-2 GPU visit Q (1 GPU): batched entry-parallel epilogue vs the other forms
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tiled.py -m gpu -q -x > gpurun_out/r2q_pytest_tiled.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest_tiled.txt
timeout 500 python tools/spmv_variants.py --reps 30 --combos "greedy,0,8,-;greedy,0,8,0;greedy,0,8,1;greedy,0,8,4" > gpurun_out/r2q_spmv.txt 2>&1
tail -3 gpurun_out/r2q_pytest_tiled.txt | cut -c1-300; cut -c1-330 gpurun_out/r2q_spmv.txt
