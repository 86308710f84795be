#!/bin/bash
# round-2 GPU visit R (1 GPU): exp / power cones in the batch kernel; batch tests; cone parity (cone3.cuh move); batch phases
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch_dist.py tests/test_gpu_parity.py -m gpu -q -s -k "batch or cone or exp or pow" > gpurun_out/r2r_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.txt
timeout 300 python tools/batch_phases.py > gpurun_out/r2r_batch_phases.txt 2>&1
grep -v "^$" gpurun_out/r2r_pytest.txt | tail -12 | cut -c1-260; cut -c1-200 gpurun_out/r2r_batch_phases.txt | tail -3
