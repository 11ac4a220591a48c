#!/bin/bash
# Pools: compact CTA shape for the strided forward pools only; ncu source capture of the generator weight gradient.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_i3d.py -x -q -k "maxpool or forward" > gpurun_out/r02n_i3d_tests.log 2>&1; tail -3 gpurun_out/r02n_i3d_tests.log
timeout 200 python tools/time_i3d.py 32 > gpurun_out/r02n_time_i3d_b32.txt 2>&1; head -12 gpurun_out/r02n_time_i3d_b32.txt; grep maxpool gpurun_out/r02n_time_i3d_b32.txt | head -8
REPS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_wgrad_v3" -c 6 -f -o gpurun_out/r02n_wgrad_v3 python tools/time_gen.py wgrad > gpurun_out/r02n_ncu.log 2>&1
ls -la gpurun_out/r02n_wgrad_v3.ncu-rep
timeout 100 python tools/time_gen.py wgrad,fwd
