#!/bin/bash
# round-2 GPU visit C (1 GPU): new parity / boundary tests, tiled SpMV variants
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_full_size.py tests/test_gpu_tiled.py -x -q -s > gpurun_out/r2c_pytest_new.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest_new.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "psd or root_plus" > gpurun_out/r2c_pytest_psd.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest_psd.txt
timeout 900 python tools/spmv_variants.py --reps 30 > gpurun_out/r2c_spmv_variants.txt 2>&1; echo "variants rc=$?" >> gpurun_out/r2c_spmv_variants.txt
tail -25 gpurun_out/r2c_pytest_new.txt | cut -c1-300; tail -15 gpurun_out/r2c_pytest_psd.txt | cut -c1-300; cat gpurun_out/r2c_spmv_variants.txt | cut -c1-400
