# round 2, call 1: lane-per-sweep eikonal pipeline -- parity subset, then one cfg-3 stage step
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sweep_bit_exact or calsurfg or full_size or variants or heap_slab" ) > gpurun_out/r2s1_pytest.log 2>&1
tail -n 15 gpurun_out/r2s1_pytest.log
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu --lsmr-iters 0 --no-dispersion > gpurun_out/r2s1_bench.json 2> gpurun_out/r2s1_bench.err
cat gpurun_out/r2s1_bench.json; tail -n 5 gpurun_out/r2s1_bench.err
