# round 2, call 24: where the time of the 2049^2 sweep stage goes (fast-iterative pipeline): stage breakdown at 1024 and 3552 sweeps
mkdir -p gpurun_out
for n in 1024 3552; do DSURF_EIKONAL=fim timeout 600 python scripts/profile_eikonal.py 259 $n 1 2>&1 | tail -n 1 | cut -c1-420; done | tee gpurun_out/r2s24_2049.log
DSURF_EIKONAL=fim timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s24_launches_2049.csv python scripts/profile_eikonal.py 259 1024 1 > /dev/null 2>&1
grep -E "k_fim|k_refine|k_rays|k_rows|k_list|k_clear|Memset|memset" gpurun_out/r2s24_launches_2049.csv | cut -d, -f5,15- | cut -c1-160 | head -40
