# round 2, call 26: A/B of k_fim_march source variants (two resident waves, 7104 sweeps at 1025^2, eikonal ms)
mkdir -p gpurun_out
for rep in 1 2; do for v in v0_97b9907 v1_refactor v2_current v3_divdefer_oldmark v4_olddiv_newmark; do
  echo -n "$v: "; DSURF_B200_LIB=$PWD/scripts/ab/lib_$v.so DSURF_EIKONAL=fim timeout 300 python scripts/profile_eikonal.py 131 7104 1 2>&1 | tail -n 1 | grep -o "'eikonal_ms': np.float64([0-9.]*)"
done; done | tee gpurun_out/r2s26_ab.log
