#!/bin/bash
# round-2 GPU visit B (2 GPUs): NCCL check of the row-partitioned mode, then the strong-scaling bench line at N=2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2b_smi.txt 2>&1
export NCCL_DEBUG=WARN
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tests/dist_gpu_check.py > gpurun_out/r2b_dist_check_2gpu.txt 2>&1; echo "dist check rc=$?" >> gpurun_out/r2b_dist_check_2gpu.txt
tail -8 gpurun_out/r2b_dist_check_2gpu.txt | cut -c1-400
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err; echo "bench rc=$?" >> gpurun_out/r2b_bench_n2.err
tail -c 1500 gpurun_out/r2b_bench_n2.json; tail -5 gpurun_out/r2b_bench_n2.err
