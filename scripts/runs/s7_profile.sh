# new-feature GPU tests, dispersion stage timing at cfg 3, launch list (time + DRAM bytes) of one full stage
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q -k "driver or raypath" ) > gpurun_out/s7_pytest.log 2>&1
tail -n 15 gpurun_out/s7_pytest.log
timeout 600 python bench.py --step-mode type --steps 1 --warmup 1 --no-cpu --lsmr-iters 0 > gpurun_out/s7_bench_disp.json 2> gpurun_out/s7_bench_disp.err
tail -c 700 gpurun_out/s7_bench_disp.json; tail -n 5 gpurun_out/s7_bench_disp.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/r01_launches_v3.csv python bench.py --steps 1 --warmup 1 --no-cpu --lsmr-iters 5 --no-dispersion \
  > gpurun_out/s7_ncu_bench.json 2> gpurun_out/s7_ncu_bench.err
tail -n 3 gpurun_out/s7_ncu_bench.err; wc -l gpurun_out/r01_launches_v3.csv
