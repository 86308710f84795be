#!/bin/bash
# round-2 GPU visit N (1 GPU): box cone in the batch kernel, the full-size SOCP / SDP fixtures, the reference arm with the
# driver's arguments (budget logic), ncu full capture of the G product in the 8192x8192 geometry (profiles/traffic.json)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch_dist.py tests/test_gpu_full_size.py -m gpu -q -s -k "box or mpc_vs or cfg3 or cfg4 or large_count" > gpurun_out/r2n_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.txt
timeout 300 python tools/batch_phases.py > gpurun_out/r2n_batch_phases.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tiled -c 48 -f -o gpurun_out/r2n_tiled_g python tools/tiled_profile.py --reps 2 > gpurun_out/r2n_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r2n_ncu.log
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2n_bench_reference.json 2> gpurun_out/r2n_bench_reference.err ) 2> gpurun_out/r2n_bench_reference.time; echo "ref rc=$?" >> gpurun_out/r2n_bench_reference.err
grep -v "^$" gpurun_out/r2n_pytest.txt | tail -14 | cut -c1-260; cut -c1-300 gpurun_out/r2n_batch_phases.txt | tail -6; tail -4 gpurun_out/r2n_ncu.log; tail -c 900 gpurun_out/r2n_bench_reference.json; cat gpurun_out/r2n_bench_reference.time
