#!/bin/bash
# Pool kernels with the CTA = (channels x 2 column blocks x 4 rows x 2 frames) decomposition.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_i3d.py -x -q > gpurun_out/r02m_i3d_tests.log 2>&1; tail -3 gpurun_out/r02m_i3d_tests.log
timeout 200 python tools/time_i3d.py 32 > gpurun_out/r02m_time_i3d_b32.txt 2>&1; head -14 gpurun_out/r02m_time_i3d_b32.txt; grep maxpool gpurun_out/r02m_time_i3d_b32.txt
