# round 2, call 13: measured parity budgets (ulp flips, pattern mismatches, Vs deviations) + the widened parity tests
mkdir -p gpurun_out
timeout 900 python scripts/parity_stats.py cfg2 > gpurun_out/r2s13_parity_stats.json 2> gpurun_out/r2s13_parity_stats.err
cat gpurun_out/r2s13_parity_stats.json; tail -n 3 gpurun_out/r2s13_parity_stats.err
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "depthkernel or outer_iteration or cfg2" ) > gpurun_out/r2s13_pytest.log 2>&1
tail -n 12 gpurun_out/r2s13_pytest.log
