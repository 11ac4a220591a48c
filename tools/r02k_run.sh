#!/bin/bash
# Round-2 late check: templated max-pools, tap-parallel wgrad reduce, resident-weight 64-wide tap GEMM.
set -x
mkdir -p gpurun_out
DMC_RESIDENT_B=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02k_suite.log 2>&1; tail -5 gpurun_out/r02k_suite.log
timeout 200 python tools/time_i3d.py 32 > gpurun_out/r02k_time_i3d_b32.txt 2>&1; head -12 gpurun_out/r02k_time_i3d_b32.txt
DMC_RESIDENT_B=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02k_bench_rb.json 2> gpurun_out/r02k_bench_rb.err
DMC_RESIDENT_B=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02k_bench_norb.json 2> gpurun_out/r02k_bench_norb.err
python - <<'P'
import json
for f in ('gpurun_out/r02k_bench_rb.json', 'gpurun_out/r02k_bench_norb.json'):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['frac'], {k: v for k, v in d['kernel_breakdown_ms_per_step'].items() if v > 0.5})
    except Exception as e:
        print(f, 'ERR', e)
P
