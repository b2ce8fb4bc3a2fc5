# round 2, call 7: per-source-line instruction counts of k_march_lps (SourceCounters only; 257^2 grid, 4096 sweeps)
mkdir -p gpurun_out
timeout 600 ncu --section SourceCounters --section WarpStateStats --import-source on --clock-control none -k regex:k_march_lps -c 1 -f -o gpurun_out/r2s7_lps_src python scripts/profile_eikonal.py 35 512 8 > gpurun_out/r2s7_prof.log 2>&1
tail -n 3 gpurun_out/r2s7_prof.log
ls -la gpurun_out/r2s7_lps_src.ncu-rep
