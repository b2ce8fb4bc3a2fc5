# round 2, call 10: how does the time per acceptance depend on the number of resident sweeps and on the footprint?
mkdir -p gpurun_out
for cfg in "35 128 8" "35 512 8" "35 1024 8" "35 3072 8" "131 128 8" "131 512 8" "131 1024 8"; do
  echo "== $cfg"; timeout 300 python scripts/profile_eikonal.py $cfg 2>&1 | tail -n 1
done > gpurun_out/r2s10_scaling.log 2>&1
cat gpurun_out/r2s10_scaling.log
