# round 2, call 5: ncu --set full with source correlation of k_march_lps (257^2 grid, 24576 resident sweeps)
mkdir -p gpurun_out
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:k_march_lps -c 1 -f -o gpurun_out/r2s5_lps_257 python scripts/profile_eikonal.py 35 3072 8 > gpurun_out/r2s5_prof.log 2>&1
tail -n 3 gpurun_out/r2s5_prof.log
ls -la gpurun_out/r2s5_lps_257.ncu-rep
