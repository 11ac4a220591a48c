#!/bin/bash
# elect.sync for the TMA-producer / MMA-issuing lanes (uniform-datapath descriptors instead of a per-MMA
# uniformisation loop): full GPU suite, bench, I3D per-family times.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02q_suite.log 2>&1; tail -3 gpurun_out/r02q_suite.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err
timeout 200 python tools/time_i3d.py 32 > gpurun_out/r02q_time_i3d_b32.txt 2>&1; head -16 gpurun_out/r02q_time_i3d_b32.txt
python - <<'P'
import json
d = json.loads([l for l in open('gpurun_out/r02q_bench.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks'], {k: v for k, v in d['kernel_breakdown_ms_per_step'].items() if v > 0.25})
for k, v in d.get('configs', {}).items():
    print(k, {a: v[a] for a in ('value', 'ms_per_step') if a in v}, v.get('e2e', {}).get('value'))
P
