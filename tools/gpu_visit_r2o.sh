#!/bin/bash
# round-2 GPU visit O (1 GPU): chunked epilogue pass (entry-parallel short rows) vs the row-parallel one, tile geometry
# 16384x4096 with 32 warps (cfg 2: 3 stages, cfg 4: 2 stages) vs the default 8192x8192
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tiled.py -m gpu -q -x > gpurun_out/r2o_pytest_tiled.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest_tiled.txt
timeout 500 python tools/spmv_variants.py --reps 30 --combos "greedy,0,8,0;greedy,0,8,4" > gpurun_out/r2o_spmv_default.txt 2>&1
SCS_B200_LIBPATH=$PWD/scs_python_b200/libscsb200_cfg2.so timeout 400 python tools/spmv_variants.py --reps 30 --combos "greedy,0,8,0" > gpurun_out/r2o_spmv_cfg2.txt 2>&1
SCS_B200_LIBPATH=$PWD/scs_python_b200/libscsb200_cfg4.so timeout 400 python tools/spmv_variants.py --reps 30 --combos "greedy,0,8,0" > gpurun_out/r2o_spmv_cfg4.txt 2>&1
tail -5 gpurun_out/r2o_pytest_tiled.txt | cut -c1-300; cut -c1-330 gpurun_out/r2o_spmv_default.txt gpurun_out/r2o_spmv_cfg2.txt gpurun_out/r2o_spmv_cfg4.txt
