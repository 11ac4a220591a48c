#!/bin/bash
# End-of-round records with the final kernels (templated max-pools, tap-parallel wgrad reduce):
# GPU suite, both bench arms, ncu launch lists of config 2 and of the I3D step, --set full capture of the pools.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_suite_final4.log 2>&1; tail -3 gpurun_out/r02_gpu_suite_final4.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_final5.json 2> gpurun_out/r02_bench_final5.err; cut -c1-200 gpurun_out/r02_bench_final5.json
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference3.json 2>/dev/null; cut -c1-300 gpurun_out/r02_bench_reference3.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 1200 --csv --log-file gpurun_out/r02_launches_dmcnet_b64.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-extras > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_dmcnet_b64.csv > gpurun_out/r02_launches_dmcnet_b64.txt
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 800 --csv \
    --log-file gpurun_out/r02_launches_i3d_b16.csv python tools/time_i3d.py 16 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_i3d_b16.csv > gpurun_out/r02_launches_i3d_b16.txt
timeout 500 ncu --set full --clock-control none -k regex:"maxpool3d" -s 4 -c 20 -f -o /tmp/i3d_pools python tools/time_i3d.py 16 > /dev/null 2>&1
ncu -i /tmp/i3d_pools.ncu-rep --page raw --csv > gpurun_out/r02_ncu_i3d_pools_raw.csv 2>/dev/null
python tools/summarize_ncu_raw.py gpurun_out/r02_ncu_i3d_pools_raw.csv > gpurun_out/r02_ncu_i3d_pools_summary.txt
head -12 gpurun_out/r02_launches_dmcnet_b64.txt; head -14 gpurun_out/r02_launches_i3d_b16.txt; head -24 gpurun_out/r02_ncu_i3d_pools_summary.txt
