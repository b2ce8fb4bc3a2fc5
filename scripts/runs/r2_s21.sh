# round 2, call 21: ncu launch list (gpu__time_duration) of the default bench command at 1 step / 1 warm-up, and one
# --set full capture of k_fim_march at one resident wave (3552 sweeps, 1025^2)
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2s21_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-dispersion --no-calsurfg-e2e --lsmr-iters 2 > gpurun_out/r2s21_bench_under_ncu.json 2> gpurun_out/r2s21_bench_under_ncu.err
tail -n 2 gpurun_out/r2s21_bench_under_ncu.err | cut -c1-300
wc -l gpurun_out/r2s21_launches.csv
DSURF_EIKONAL=fim timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_fim_march -c 1 -f -o gpurun_out/r2s21_fim_full python scripts/profile_eikonal.py 131 3552 1 > gpurun_out/r2s21_prof.log 2>&1
tail -n 2 gpurun_out/r2s21_prof.log | cut -c1-300
gzip -f gpurun_out/r2s21_launches.csv
