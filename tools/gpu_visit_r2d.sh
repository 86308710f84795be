#!/bin/bash
# round-2 GPU visit D (1 GPU): remaining new tests, ncu full capture of the tiled streaming kernel (source view)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_full_size.py tests/test_gpu_tiled.py -q -s > gpurun_out/r2d_pytest_new.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest_new.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tiled_kernel -s 6 -c 1 -o gpurun_out/r2d_tiled_stream -f python tools/tiled_profile.py > gpurun_out/r2d_ncu_tiled.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r2d_ncu_tiled.log
tail -30 gpurun_out/r2d_pytest_new.txt | cut -c1-300; tail -5 gpurun_out/r2d_ncu_tiled.log
