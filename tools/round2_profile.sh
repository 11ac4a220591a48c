#!/bin/bash
# Round-2 measurement call: bench lines (both arms), ncu launch lists and --set full captures.
#   gpurun --timeout 2400 -- 'bash tools/round2_profile.sh'
set -x
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; cut -c1-300 gpurun_out/r02_bench_final.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_bench_reference.json 2>/dev/null; cut -c1-400 gpurun_out/r02_bench_reference.json
# launch lists (cold-cache, serialised: shares only)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 1200 --csv --log-file gpurun_out/r02_launches_dmcnet_b64.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-extras > /dev/null 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 2400 --csv --log-file gpurun_out/r02_launches_gan_b64.csv \
    python bench.py --config gan --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-extras > /dev/null 2>&1
# --set full: the classifier GEMMs and the generator convs (config 2)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tap_gemm_ws|wgrad_gemm|wgrad64" -s 70 -c 16 -o /tmp/r02_gemms \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-extras > /dev/null 2>&1
ncu -i /tmp/r02_gemms.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_ncu_gemms_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:"conv3x3_fwd_v3|conv3x3_wgrad_v3" -s 18 -c 17 -o /tmp/r02_gen \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-extras > /dev/null 2>&1
ncu -i /tmp/r02_gen.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_ncu_generator_raw.csv 2>/dev/null
# --set full: the discriminator plan (config 3)
timeout 600 ncu --set full --clock-control none -k regex:"tap_gemm_ws|wgrad|pm_act_bwd|bn_apply_kernel|planar_to_s2d4|planar_to_pm_ring" -s 260 -c 40 -o /tmp/r02_disc \
    python bench.py --config gan --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-extras > /dev/null 2>&1
ncu -i /tmp/r02_disc.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_ncu_disc_raw.csv 2>/dev/null
ls -la gpurun_out | grep r02_
