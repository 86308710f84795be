#!/bin/bash
# round-2 GPU visit P (1 GPU): default geometry 16384x4096 / 32 warps + chunked epilogue: whole GPU tier, bench line with the
# driver's arguments, variants (U, dealing, epilogue form, cfg 1 library), ncu full capture of the epilogue kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2p_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest_gpu.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err; echo "bench rc=$?" >> gpurun_out/r2p_bench_n1.err
timeout 500 python tools/spmv_variants.py --reps 30 --combos "greedy,0,8,0;greedy,0,6,0;perm,0,8,0;greedy,0,8,4" > gpurun_out/r2p_spmv_default.txt 2>&1
SCS_B200_LIBPATH=$PWD/scs_python_b200/libscsb200_cfg1.so timeout 400 python tools/spmv_variants.py --reps 30 --combos "greedy,0,8,0" > gpurun_out/r2p_spmv_cfg1.txt 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:epilogue -c 8 -f -o gpurun_out/r2p_epilogue python tools/tiled_profile.py --reps 2 > gpurun_out/r2p_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r2p_ncu.log
SCS_B200_TILED_EB=4 timeout 500 ncu --set full --clock-control none --import-source on -k regex:epilogue -c 8 -f -o gpurun_out/r2p_epilogue_eb4 python tools/tiled_profile.py --reps 2 > gpurun_out/r2p_ncu_eb4.log 2>&1
tail -14 gpurun_out/r2p_pytest_gpu.txt | cut -c1-200; tail -c 1500 gpurun_out/r2p_bench_n1.json; echo; cut -c1-330 gpurun_out/r2p_spmv_default.txt gpurun_out/r2p_spmv_cfg1.txt; tail -2 gpurun_out/r2p_ncu.log
