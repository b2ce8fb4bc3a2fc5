# ncu --set full of the LSMR iteration kernels (stand-alone probe system, cfg-3 shape)
mkdir -p gpurun_out
DSURF_LSMR_NO_GRAPH=1 timeout 600 ncu --set full --import-source on --clock-control none \
  -k regex:"k_bspmv|k_fused" --launch-skip 24 -c 6 -o gpurun_out/r01_lsmr_iter -f \
  python scripts/lsmr_bench.py --iters 12 > gpurun_out/s10_ncu.log 2>&1
tail -n 4 gpurun_out/s10_ncu.log; ls -la gpurun_out/*.ncu-rep
