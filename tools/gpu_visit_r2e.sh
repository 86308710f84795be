#!/bin/bash
# round-2 GPU visit E (1 GPU): tile-geometry A/B of the tiled SpMV engine, batch phase clocks, fixed tests, baseline table
mkdir -p gpurun_out
ALT=$PWD/scs_python_b200/libscsb200_cfg1.so
timeout 600 python tools/spmv_variants.py --reps 30 --combos "1,0,8" > gpurun_out/r2e_spmv_cfg0.txt 2>&1
SCS_B200_LIBPATH=$ALT timeout 600 python tools/spmv_variants.py --reps 30 --combos "1,0,8;1,0,6" > gpurun_out/r2e_spmv_cfg1.txt 2>&1
SCS_B200_LIBPATH=$ALT timeout 600 python -m pytest tests/test_gpu_tiled.py -q > gpurun_out/r2e_pytest_tiled_cfg1.txt 2>&1
timeout 300 python tools/batch_phases.py > gpurun_out/r2e_batch_phases.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_boundary.py -q -k "two_threads or linear_solver_b200" > gpurun_out/r2e_pytest_boundary.txt 2>&1
timeout 1500 python tools/bench_configs.py --configs 1,1lp,3,4,5s,2s --ref-time-limit 60 --ref-batch-sample 16 > gpurun_out/r2e_configs.jsonl 2> gpurun_out/r2e_configs.err
cat gpurun_out/r2e_spmv_cfg0.txt gpurun_out/r2e_spmv_cfg1.txt | cut -c1-260; tail -3 gpurun_out/r2e_pytest_tiled_cfg1.txt; cat gpurun_out/r2e_batch_phases.txt | cut -c1-500; tail -3 gpurun_out/r2e_pytest_boundary.txt; cut -c1-220 gpurun_out/r2e_configs.jsonl
