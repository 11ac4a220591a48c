#!/bin/bash
# A/B: four vs eight epilogue warps in the 64-wide row-window tap GEMM (DMC_EPILOGUE_WARPS).
set -x
mkdir -p gpurun_out
DMC_EPILOGUE_WARPS=8 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_disc_tc.py tests/test_gpu_backward_exact.py tests/test_gpu_parity.py -x -q > gpurun_out/ab_epw_tests.log 2>&1; tail -3 gpurun_out/ab_epw_tests.log
for w in 4 8; do
  DMC_EPILOGUE_WARPS=$w timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ab_epw_c2_epw$w.json 2> gpurun_out/ab_epw_c2_epw$w.err
  DMC_EPILOGUE_WARPS=$w timeout 300 python bench.py --config gan --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ab_epw_c3_epw$w.json 2> gpurun_out/ab_epw_c3_epw$w.err
done
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/ab_epw_c*_epw*.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        kb = d['kernel_breakdown_ms_per_step']
        print(f, round(d['value'], 1), round(d['ms_per_step'], 3), round(d['roofline']['frac'], 4), {k: kb[k] for k in ('tc_tap_gemm', 'tc_tap_gemm_act', 'tc_wgrad') if k in kb}, d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
P
