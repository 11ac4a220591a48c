#!/bin/bash
# Last records of the round (eight epilogue warps on the 64-wide row-window GEMMs, elect.sync issuing threads):
# GPU suite, both bench arms, ncu launch lists of config 2 / 3, --set full capture of the classifier GEMMs.
set -x
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_suite_final6.log 2>&1; tail -3 gpurun_out/r02_gpu_suite_final6.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_final7.json 2> gpurun_out/r02_bench_final7.err; cut -c1-200 gpurun_out/r02_bench_final7.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 1200 --csv --log-file gpurun_out/r02_launches_dmcnet_b64.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-extras > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_dmcnet_b64.csv > gpurun_out/r02_launches_dmcnet_b64.txt
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 2400 --csv --log-file gpurun_out/r02_launches_gan_b64.csv \
    python bench.py --config gan --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-extras > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_gan_b64.csv > gpurun_out/r02_launches_gan_b64.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tap_gemm_ws|wgrad_gemm|wgrad64" -s 70 -c 16 -o /tmp/r02_gemms \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-extras > /dev/null 2>&1
ncu -i /tmp/r02_gemms.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_ncu_gemms_raw.csv 2>/dev/null
python tools/summarize_ncu_raw.py gpurun_out/r02_ncu_gemms_raw.csv --traffic gpurun_out/r02_tap_gemm_traffic.json > gpurun_out/r02_ncu_gemms_summary.txt
head -12 gpurun_out/r02_launches_dmcnet_b64.txt; head -12 gpurun_out/r02_launches_gan_b64.txt; head -22 gpurun_out/r02_ncu_gemms_summary.txt
