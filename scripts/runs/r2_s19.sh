# round 2, call 19: fast-iterative eikonal v3 (straight-line rule for the common case + generic fallback): parity tests,
# register variants on two resident waves at 1025^2, ncu metrics + per-line counts on a quarter wave
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_fim.py -m gpu -x -q ) > gpurun_out/r2s19_pytest.log 2>&1
tail -n 6 gpurun_out/r2s19_pytest.log
for v in 6 5 4 8; do echo "MINB $v"; DSURF_EIKONAL=fim DSURF_FIM_MINB=$v timeout 300 python scripts/profile_eikonal.py 131 7104 1 2>&1 | tail -n 1 | cut -c1-330; done | tee gpurun_out/r2s19_variants.log
DSURF_EIKONAL=fim timeout 600 ncu --metrics $(cat scripts/ncu_eik_metrics.txt) --clock-control none -k regex:"k_fim_march" --csv --log-file gpurun_out/r2s19_fim_metrics.csv python scripts/profile_eikonal.py 131 888 1 > gpurun_out/r2s19_prof.log 2>&1
tail -n 2 gpurun_out/r2s19_prof.log | cut -c1-300
DSURF_EIKONAL=fim timeout 600 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --import-source on --clock-control none -k regex:k_fim_march -c 1 -f -o gpurun_out/r2s19_fim_src python scripts/profile_eikonal.py 131 888 1 > gpurun_out/r2s19_prof2.log 2>&1
tail -n 2 gpurun_out/r2s19_prof2.log | cut -c1-300
