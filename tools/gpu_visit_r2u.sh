#!/bin/bash
# round-2 GPU visit U (1 GPU): bench.py with the other_configs key on a small-scale workload (plumbing check)
mkdir -p gpurun_out
timeout 280 python bench.py --scale 0.05 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench_small.json 2> gpurun_out/r2u_bench_small.err; echo "bench rc=$?" >> gpurun_out/r2u_bench_small.err
tail -3 gpurun_out/r2u_bench_small.err | cut -c1-300; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2u_bench_small.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac']); print(json.dumps(d['other_configs'])[:1800])
P
