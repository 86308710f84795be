#!/bin/bash
# round-2 GPU visit M (1 GPU): what the driver runs at round end -- the whole GPU tier, smoke(), the bench line with the
# driver's arguments, the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r2m_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r2m_smoke.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err; echo "bench rc=$?" >> gpurun_out/r2m_bench_n1.err
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2m_bench_reference.json 2> gpurun_out/r2m_bench_reference.err ) 2> gpurun_out/r2m_bench_reference.time; echo "ref rc=$?" >> gpurun_out/r2m_bench_reference.err
tail -22 gpurun_out/r2m_pytest_gpu.txt | cut -c1-200; tail -2 gpurun_out/r2m_smoke.txt; tail -c 700 gpurun_out/r2m_bench_n1.json; echo; tail -c 900 gpurun_out/r2m_bench_reference.json; cat gpurun_out/r2m_bench_reference.time
