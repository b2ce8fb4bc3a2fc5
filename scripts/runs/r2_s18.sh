# round 2, call 18: fast-iterative eikonal v2 (one warp per sweep, cached quadrant solutions, TMA bulk tile loads):
# parity tests, cfg-3 stage timing for the register / TMA variants, ncu metrics at one resident wave
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_fim.py -m gpu -x -q ) > gpurun_out/r2s18_pytest.log 2>&1
tail -n 6 gpurun_out/r2s18_pytest.log
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-calsurfg-e2e --no-dispersion --lsmr-iters 0"
DSURF_EIKONAL=fim timeout 400 $B > gpurun_out/r2s18_bench_fim_default.json 2> gpurun_out/r2s18_bench_fim_default.err
DSURF_EIKONAL=fim DSURF_FIM_TMA=0 timeout 400 $B > gpurun_out/r2s18_bench_fim_notma.json 2> gpurun_out/r2s18_bench_fim_notma.err
DSURF_EIKONAL=fim DSURF_FIM_MINB=4 timeout 400 $B > gpurun_out/r2s18_bench_fim_minb4.json 2> gpurun_out/r2s18_bench_fim_minb4.err
DSURF_EIKONAL=fim DSURF_FIM_MINB=8 timeout 400 $B > gpurun_out/r2s18_bench_fim_minb8.json 2> gpurun_out/r2s18_bench_fim_minb8.err
python - <<'PY'
import json
for f in ("default","notma","minb4","minb8"):
    try:
        d=json.load(open(f"gpurun_out/r2s18_bench_fim_{f}.json"))
        print(f, {k:d[k] for k in ("value","ms_per_step")}, d["stage_ms_per_step"], "e2e", d["e2e"]["value"], d["coo"]["digest"])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/r2s18_bench_fim_{f}.err").read()[-1500:])
PY
DSURF_EIKONAL=fim timeout 600 ncu --metrics $(cat scripts/ncu_eik_metrics.txt) --clock-control none -k regex:"k_fim|k_refine" --csv --log-file gpurun_out/r2s18_fim_metrics.csv python scripts/profile_eikonal.py 131 3552 1 > gpurun_out/r2s18_prof.log 2>&1
tail -n 2 gpurun_out/r2s18_prof.log
DSURF_EIKONAL=fim timeout 600 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --import-source on --clock-control none -k regex:k_fim_march -c 1 -f -o gpurun_out/r2s18_fim_src python scripts/profile_eikonal.py 131 3552 1 > gpurun_out/r2s18_prof2.log 2>&1
tail -n 2 gpurun_out/r2s18_prof2.log
