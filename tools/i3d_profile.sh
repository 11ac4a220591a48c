# ncu evidence for the I3D step (BASELINE config 5): launch list with per-launch device times, and a
# --set full capture of the dominant kernels.  Run on a GPU box: bash tools/i3d_profile.sh
set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 800 --csv \
    --log-file gpurun_out/r02_launches_i3d_b16.csv python tools/time_i3d.py 16 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_i3d_b16.csv > gpurun_out/r02_launches_i3d_b16.txt
timeout 900 ncu --set full --clock-control none \
    -k regex:"tap_gemm_ws|wgrad_gemm|maxpool3d|i3d_stem" -s 40 -c 36 -o /tmp/i3d python tools/time_i3d.py 16 > /dev/null 2>&1
ncu -i /tmp/i3d.ncu-rep --page raw --csv > gpurun_out/r02_ncu_i3d_raw.csv 2>/dev/null
python tools/summarize_ncu_raw.py gpurun_out/r02_ncu_i3d_raw.csv > gpurun_out/r02_ncu_i3d_summary.txt
head -50 gpurun_out/r02_ncu_i3d_summary.txt
