# round 2, call 11: new bench.py code paths on the tiny configuration + parity subset after the refactors
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "calsurfg or device_glue or lsmr_matches or outer_iteration or sweep_bit_exact" ) > gpurun_out/r2s11_pytest.log 2>&1
tail -n 6 gpurun_out/r2s11_pytest.log
timeout 600 python bench.py --config 0 --steps 2 --warmup 1 > gpurun_out/r2s11_bench0.json 2> gpurun_out/r2s11_bench0.err
tail -n 5 gpurun_out/r2s11_bench0.err; cut -c1-1500 gpurun_out/r2s11_bench0.json
