# round 2, call 9: fixed round-trip schedule (complete chain prefetch, tracked last element) -- parity subset, cfg-3 stage step, metrics at full residency,
# per-line instruction counts
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sweep_bit_exact or calsurfg_small or full_size" ) > gpurun_out/r2s9_pytest.log 2>&1
tail -n 6 gpurun_out/r2s9_pytest.log
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --lsmr-iters 0 --no-dispersion > gpurun_out/r2s9_bench.json 2> gpurun_out/r2s9_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s9_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','stage_ms_per_step','gpu_launches')}, d['e2e']['value'])
PY
tail -n 3 gpurun_out/r2s9_bench.err
timeout 900 ncu --metrics $(cat scripts/ncu_eik_metrics.txt) --clock-control none -k regex:k_march_lps --csv --log-file gpurun_out/r2s9_lps_257_metrics.csv python scripts/profile_eikonal.py 35 3072 8 > gpurun_out/r2s9_prof.log 2>&1
tail -n 2 gpurun_out/r2s9_prof.log
timeout 600 ncu --section SourceCounters --section WarpStateStats --import-source on --clock-control none -k regex:k_march_lps -c 1 -f -o gpurun_out/r2s9_lps_src python scripts/profile_eikonal.py 35 512 8 > gpurun_out/r2s9_prof2.log 2>&1
