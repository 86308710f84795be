#!/bin/bash
# round-2 GPU visit I (2 GPUs): NCCL check of the row-partitioned mode against the compiled reference
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tests/dist_gpu_check.py > gpurun_out/r2i_dist_check_2gpu.txt 2>&1; echo "dist check rc=$?" >> gpurun_out/r2i_dist_check_2gpu.txt
grep -v "^\s*$" gpurun_out/r2i_dist_check_2gpu.txt | tail -12 | cut -c1-600
