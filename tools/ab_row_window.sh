#!/bin/bash
# A/B of the row-window schedule of the K = 64 tap GEMMs (DMC_NO_ROW_WINDOW=1 disables it):
# kernel tests first, then config 2 and config 3 with and without it.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_disc_tc.py tests/test_gpu_backward_exact.py -q -p no:cacheprovider > gpurun_out/r02_rw_tests.log 2>&1
grep -n "^FAILED\|passed\|failed\|^E   " gpurun_out/r02_rw_tests.log | cut -c1-300 | head -20
for v in 0 1; do
  if [ $v = 1 ]; then export DMC_NO_ROW_WINDOW=1; fi
  python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_rw_cfg2_$v.json 2>/dev/null
  python bench.py --config gan --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_rw_gan_$v.json 2>/dev/null
  python - <<EOF
import json
for c in ("cfg2", "gan"):
    d = json.load(open("gpurun_out/r02_rw_%s_$v.json" % c)); b = d["kernel_breakdown_ms_per_step"]
    print("NO_RW=$v", c, round(d["ms_per_step"], 3), {k: b[k] for k in ("tc_tap_gemm", "tc_tap_gemm_act", "tc_wgrad") if k in b}, round(d["roofline"]["frac"], 4))
EOF
done
