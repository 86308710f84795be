#!/bin/bash
# round-2 GPU visit S (1 GPU): final defaults (16384x4096 tiles, chunked Gp epilogue): tiled + dist self-test + smoke, bench line
# with the driver's arguments, ncu full capture of both products for profiles/traffic.json
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_tiled.py tests/test_gpu_dist.py -m gpu -q > gpurun_out/r2s_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r2s_smoke.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2s_bench_n1.json 2> gpurun_out/r2s_bench_n1.err; echo "bench rc=$?" >> gpurun_out/r2s_bench_n1.err
timeout 500 ncu --set full --clock-control none --import-source on -k regex:tiled -c 16 -f -o gpurun_out/r2s_tiled python tools/tiled_profile.py --reps 2 > gpurun_out/r2s_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r2s_ncu.log
tail -4 gpurun_out/r2s_pytest.txt | cut -c1-200; tail -2 gpurun_out/r2s_smoke.txt; tail -c 600 gpurun_out/r2s_bench_n1.json; echo; tail -2 gpurun_out/r2s_ncu.log
