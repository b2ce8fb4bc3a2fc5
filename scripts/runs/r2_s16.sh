# round 2, call 16: first GPU run of the fast-iterative eikonal (DSURF_EIKONAL=fim): parity tests, deviation table on
# cfg 1 / small / cfg 2, timing at cfg 2 and cfg 3
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_fim.py -m gpu -x -q ) > gpurun_out/r2s16_pytest.log 2>&1
tail -n 15 gpurun_out/r2s16_pytest.log
timeout 600 python scripts/fim_parity.py taipei small cfg2 > gpurun_out/r2s16_fim_parity.jsonl 2> gpurun_out/r2s16_fim_parity.err
cat gpurun_out/r2s16_fim_parity.jsonl; tail -n 5 gpurun_out/r2s16_fim_parity.err
DSURF_EIKONAL=fim timeout 300 python bench.py --config 2 --steps 2 --warmup 1 --no-cpu --no-calsurfg-e2e --no-dispersion > gpurun_out/r2s16_bench_cfg2_fim.json 2> gpurun_out/r2s16_bench_cfg2_fim.err
tail -n 3 gpurun_out/r2s16_bench_cfg2_fim.err
DSURF_EIKONAL=fim timeout 500 python bench.py --steps 1 --warmup 1 --no-cpu --no-calsurfg-e2e --no-dispersion > gpurun_out/r2s16_bench_cfg3_fim.json 2> gpurun_out/r2s16_bench_cfg3_fim.err
tail -n 3 gpurun_out/r2s16_bench_cfg3_fim.err
python - <<'PY'
import json
for f in ("cfg2_fim","cfg3_fim"):
    try:
        d=json.load(open(f"gpurun_out/r2s16_bench_{f}.json"))
        print(f, {k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["stage_ms_per_step"], "e2e", d["e2e"]["value"])
    except Exception as e:
        print(f, "ERR", e)
PY
