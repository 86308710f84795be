#!/bin/bash
# round-2 GPU visit G (1 GPU): thread reproducibility diagnostic, unrolled-CG graph on the small configs, rest of the GPU tier
mkdir -p gpurun_out
timeout 600 python tools/thread_diag.py > gpurun_out/r2g_thread_diag.txt 2>&1
timeout 300 python tools/bench_configs.py --configs 1,1lp,3 --no-ref > gpurun_out/r2g_configs_unroll.jsonl 2>/dev/null
SCS_B200_CG_UNROLL=0 timeout 300 python tools/bench_configs.py --configs 1,1lp,3 --no-ref > gpurun_out/r2g_configs_unroll0.jsonl 2>/dev/null
timeout 1500 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_parity.py tests/test_gpu_rw.py tests/test_gpu_tiled.py tests/test_gpu_batch_dist.py -q --durations=8 > gpurun_out/r2g_pytest_rest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest_rest.txt
cat gpurun_out/r2g_thread_diag.txt | cut -c1-600; tail -15 gpurun_out/r2g_pytest_rest.txt | cut -c1-200
