#!/bin/bash
# round-2 GPU visit K (2 GPUs): peer-memory collectives of the row-partitioned mode: parity check, bench at N = 2 (P2P and NCCL)
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tests/dist_gpu_check.py > gpurun_out/r2k_dist_check_2gpu_p2p.txt 2>&1; echo "dist check rc=$?" >> gpurun_out/r2k_dist_check_2gpu_p2p.txt
grep -v "^\s*$" gpurun_out/r2k_dist_check_2gpu_p2p.txt | grep "case\|rc=\|ok\|libscs\|rror" | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 2 --steps 8 --warmup 3 --no-batch > gpurun_out/r2k_bench_n2_p2p.json 2> gpurun_out/r2k_bench_n2_p2p.err; echo "bench rc=$?" >> gpurun_out/r2k_bench_n2_p2p.err
python -c "
import json
d=json.loads(open('gpurun_out/r2k_bench_n2_p2p.json').read().strip().splitlines()[-1])
print('N=2 p2p value', d['value'], 'ms/step', d['ms_per_step'], 'agree', d['time_to_eps']['agreement_with_single_gpu'], 'tte', d['time_to_eps']['wall_s_incl_upload'], 'single', d['time_to_eps']['single_gpu']['wall_s_incl_upload'], d['time_to_eps']['single_gpu']['solve_ms'])
"
tail -3 gpurun_out/r2k_bench_n2_p2p.err
