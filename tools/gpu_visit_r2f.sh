#!/bin/bash
# round-2 GPU visit F (1 GPU): regression of the whole GPU tier on the new defaults, batch / PSD / epilogue measurements,
# Cfg-1 graph vs stream, ncu full capture of the tiled product in the new geometry
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_dist.py::test_row_partitioned_two_gpus > gpurun_out/r2f_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_gpu.txt
timeout 300 python tools/batch_phases.py > gpurun_out/r2f_batch_phases.txt 2>&1
timeout 300 python tools/psd_bench.py > gpurun_out/r2f_psd_bench.txt 2>&1
timeout 600 python tools/spmv_variants.py --reps 30 --combos "1,0,8" > gpurun_out/r2f_spmv_eb4.txt 2>&1
SCS_B200_TILED_EB=8 timeout 600 python tools/spmv_variants.py --reps 30 --combos "1,0,8" > gpurun_out/r2f_spmv_eb8.txt 2>&1
timeout 300 python tools/bench_configs.py --configs 1,1lp,3,4 --no-ref > gpurun_out/r2f_configs_graph.jsonl 2>/dev/null
SCS_B200_NO_GRAPH=1 timeout 300 python tools/bench_configs.py --configs 1,1lp --no-ref > gpurun_out/r2f_configs_nograph.jsonl 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tiled -s 12 -c 2 -o gpurun_out/r2f_tiled_g -f python tools/tiled_profile.py > gpurun_out/r2f_ncu_tiled.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r2f_ncu_tiled.log
tail -5 gpurun_out/r2f_pytest_gpu.txt | cut -c1-300; cut -c1-420 gpurun_out/r2f_batch_phases.txt; head -8 gpurun_out/r2f_psd_bench.txt; cut -c1-200 gpurun_out/r2f_spmv_eb4.txt gpurun_out/r2f_spmv_eb8.txt; cut -c1-250 gpurun_out/r2f_configs_graph.jsonl gpurun_out/r2f_configs_nograph.jsonl; tail -3 gpurun_out/r2f_ncu_tiled.log
