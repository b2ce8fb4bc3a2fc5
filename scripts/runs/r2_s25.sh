# round 2, call 25: whole GPU suite on the final kernels; fast-iterative headline at cfg 3; the default command (exact
# headline + other_pipeline) at 1 step / 1 warm-up with the CPU legs
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2s25_pytest.log 2>&1
tail -n 6 gpurun_out/r2s25_pytest.log
timeout 400 python bench.py --eikonal fim --no-both --steps 3 --warmup 3 --no-cpu --no-dispersion --no-calsurfg-e2e --lsmr-iters 0 > gpurun_out/r2s25_bench_cfg3_fim.json 2> gpurun_out/r2s25_bench_cfg3_fim.err
tail -n 2 gpurun_out/r2s25_bench_cfg3_fim.err | cut -c1-300
timeout 900 python bench.py --steps 1 --warmup 1 > gpurun_out/r2s25_bench_default.json 2> gpurun_out/r2s25_bench_default.err
tail -n 2 gpurun_out/r2s25_bench_default.err | cut -c1-300
python - <<'PY'
import json
for f in ("cfg3_fim","default"):
    try:
        d=json.load(open(f"gpurun_out/r2s25_bench_{f}.json"))
        print(f, d["eikonal_pipeline"], {k:d[k] for k in ("value","ms_per_step")}, d["stage_ms_per_step"], "e2e", d["e2e"]["value"], d["coo"]["digest"], "roofline", round(d["roofline"]["frac"],5), d["roofline"]["traffic"])
        o=d.get("other_pipeline")
        if o: print("   other:", o["pipeline"], o["value"], o["e2e"]["value"], o["coo"], round(o["roofline"]["frac"],5))
        if d.get("lsmr"): print("   lsmr", d["lsmr"]["iters_per_s"], d["lsmr"]["nnz"], d["lsmr"]["roofline"]["frac"], d["lsmr"]["to_convergence"])
        if d.get("cpu_baseline"): print("   cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"].get("reference_threading",{}).get("value"))
        if d.get("e2e_calsurfg"): print("   e2e_calsurfg", d["e2e_calsurfg"])
        if d.get("dispersion"): print("   disp", d["dispersion"]["ms"], d["dispersion"].get("dp_gflops_est"))
    except Exception as e:
        print(f, "ERR", e)
PY
