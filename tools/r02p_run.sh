#!/bin/bash
# A/B of the TMA L2 prefetch in the row-window / one-tap tap GEMMs (DMC_TMA_PREFETCH = distance in tiles).
set -x
mkdir -p gpurun_out
DMC_TMA_PREFETCH=2 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_disc_tc.py -x -q > gpurun_out/r02p_tests.log 2>&1; tail -3 gpurun_out/r02p_tests.log
for pf in 0 2 4; do
  DMC_TMA_PREFETCH=$pf timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02p_c2_pf$pf.json 2> gpurun_out/r02p_c2_pf$pf.err
  DMC_TMA_PREFETCH=$pf timeout 300 python bench.py --config gan --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02p_c3_pf$pf.json 2> gpurun_out/r02p_c3_pf$pf.err
done
DMC_TMA_PREFETCH=2 timeout 200 python tools/time_i3d.py 32 > gpurun_out/r02p_time_i3d_pf2.txt 2>&1; head -14 gpurun_out/r02p_time_i3d_pf2.txt
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02p_c*_pf*.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        kb = d['kernel_breakdown_ms_per_step']
        print(f, round(d['value'], 1), round(d['ms_per_step'], 3), round(d['roofline']['frac'], 4), {k: kb[k] for k in ('tc_tap_gemm', 'tc_tap_gemm_act', 'tc_wgrad') if k in kb}, d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
P
