#!/bin/bash
# round-2 GPU visit T (1 GPU): small PSD cones in the batch kernel; all batch tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_batch_dist.py -m gpu -q -s -k "batch" > gpurun_out/r2t_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest.txt
grep -v "^$" gpurun_out/r2t_pytest.txt | tail -25 | cut -c1-300
