set -x
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/r01_bench_final.json 2> gpurun_out/r01_bench_final.err; cut -c1-200 gpurun_out/r01_bench_final.json
timeout 300 python bench.py --config gan --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_gan.json 2> gpurun_out/r01_bench_gan.err; cut -c1-200 gpurun_out/r01_bench_gan.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 1000 --csv --log-file gpurun_out/r01_launches_v3.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none -k regex:"tap_gemm_ws|wgrad_gemm|wgrad64|stem_conv_tc" -s 60 -c 14 -o /tmp/gemms python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > /dev/null 2>&1
ncu -i /tmp/gemms.ncu-rep --page raw --csv > gpurun_out/r01_ncu_full_gemms_raw.csv 2>/dev/null
ls -la gpurun_out/ | grep r01_
