#!/bin/bash
# Round-2 late check: templated max-pools, tap-parallel wgrad reduce.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02l_suite.log 2>&1; tail -5 gpurun_out/r02l_suite.log
timeout 200 python tools/time_i3d.py 32 > gpurun_out/r02l_time_i3d_b32.txt 2>&1; head -14 gpurun_out/r02l_time_i3d_b32.txt; grep maxpool gpurun_out/r02l_time_i3d_b32.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err
python - <<'P'
import json
for f in ('gpurun_out/r02l_bench.json',):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], {k: v for k, v in d['kernel_breakdown_ms_per_step'].items() if v > 0.5})
        for k, v in d.get('configs', {}).items():
            print(k, {a: v[a] for a in ('value', 'ms_per_step') if a in v}, v.get('e2e', {}).get('value'))
    except Exception as e:
        print(f, 'ERR', e)
P
