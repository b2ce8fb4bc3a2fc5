mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_eikonal3 -c 1 -o gpurun_out/r01_eikonal_v3_257 -f python scripts/profile_eikonal.py 35 64 16 > gpurun_out/s5_ncu.log 2>&1
tail -n 3 gpurun_out/s5_ncu.log
ls -la gpurun_out/*.ncu-rep
