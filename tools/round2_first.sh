#!/bin/bash
# First GPU call of the next round: run what round 1 wrote after its GPU budget was spent.
#   gpurun --timeout 900 -- 'bash tools/round2_first.sh'
# 1. the whole extension test file including the opt-in (never executed) paths,
# 2. the e2e leg of the bench fed with the uint8 sample stack,
# 3. an ncu capture of the two input-stage kernels (achieved HBM GB/s vs 35 B/pixel algorithmic).
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_z_extensions.py -v -p no:cacheprovider 2>&1 | tee gpurun_out/r02_ext_tests.log | tail -40
python examples/train_synthetic.py --epochs 2 --batch-size 8 --model-prefix /tmp/r02_example/hmdb51 \
    > gpurun_out/r02_example.log 2>&1; tail -5 gpurun_out/r02_example.log
python bench.py --steps 10 --warmup 3 --input u8 --no-cpu-baseline > gpurun_out/r02_bench_u8.json 2> gpurun_out/r02_bench_u8.err
tail -c 600 gpurun_out/r02_bench_u8.json
ncu --set full --clock-control none -k regex:'unpack_normalize_u8|flow_block_mean_u8' -c 4 \
    -o gpurun_out/r02_input_stage python -m pytest tests/test_gpu_z_extensions.py -q -p no:cacheprovider \
    -k "full_size" > gpurun_out/r02_ncu_input_stage.log 2>&1
ncu -i gpurun_out/r02_input_stage.ncu-rep --page raw --csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum > gpurun_out/r02_input_stage_raw.csv 2>&1
