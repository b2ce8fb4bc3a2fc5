# round 2, call 20: whole GPU test suite; bench with both eikonal pipelines in one line (cfg 2); fast-iterative headline
# at cfg 3 (slowness prefetch)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2s20_pytest.log 2>&1
tail -n 8 gpurun_out/r2s20_pytest.log
timeout 400 python bench.py --config 2 --steps 2 --warmup 1 --no-cpu --no-dispersion --no-calsurfg-e2e > gpurun_out/r2s20_bench_cfg2.json 2> gpurun_out/r2s20_bench_cfg2.err
tail -n 3 gpurun_out/r2s20_bench_cfg2.err
timeout 400 python bench.py --eikonal fim --no-both --steps 3 --warmup 3 --no-cpu --no-dispersion --no-calsurfg-e2e > gpurun_out/r2s20_bench_cfg3_fim.json 2> gpurun_out/r2s20_bench_cfg3_fim.err
tail -n 3 gpurun_out/r2s20_bench_cfg3_fim.err
python - <<'PY'
import json
for f in ("cfg2","cfg3_fim"):
    try:
        d=json.load(open(f"gpurun_out/r2s20_bench_{f}.json"))
        print(f, d["eikonal_pipeline"], {k:d[k] for k in ("value","ms_per_step")}, d["stage_ms_per_step"], "e2e", d["e2e"]["value"], d["coo"]["digest"], "roofline", d["roofline"]["frac"])
        o=d.get("other_pipeline")
        if o: print("   other:", o["pipeline"], o["value"], o["e2e"]["value"], o["coo"], o["roofline"]["frac"])
        if d.get("lsmr"): print("   lsmr", d["lsmr"]["iters_per_s"], d["lsmr"]["nnz"], d["lsmr"]["roofline"]["frac"])
    except Exception as e:
        print(f, "ERR", e)
PY
