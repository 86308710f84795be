#!/bin/bash
# round-2 GPU visit L (8 GPUs, peer-memory collectives): the strong-scaling bench line of the row-partitioned Cfg-2 at N = 8
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29657 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r2l_bench_n8_p2p.json 2> gpurun_out/r2l_bench_n8_p2p.err; echo "bench rc=$?" >> gpurun_out/r2l_bench_n8_p2p.err
tail -c 1200 gpurun_out/r2l_bench_n8_p2p.json; tail -5 gpurun_out/r2l_bench_n8_p2p.err
