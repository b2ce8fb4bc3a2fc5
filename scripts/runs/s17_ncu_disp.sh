mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_disp_columns -c 1 -o gpurun_out/r01_disp_cfg3 -f python scripts/profile_disp.py > gpurun_out/s17_ncu.log 2>&1
tail -n 4 gpurun_out/s17_ncu.log; ls -la gpurun_out/r01_disp_cfg3.ncu-rep
