# round 2, call 30: A/B -- marking without bounds tests (out-of-grid words flagged at tile load) and 5 / 6 / 7 resident CTAs per SM;
# then the parity tests and the cfg-3 COO digest with the no-bounds build (must equal 870cc4991b8bbf64)
mkdir -p gpurun_out
for v in base nobounds; do for m in 6 7 5; do
  echo -n "$v minb=$m: "; DSURF_B200_LIB=$PWD/scripts/ab/lib_$v.so DSURF_FIM_MINB=$m DSURF_EIKONAL=fim timeout 200 python scripts/profile_eikonal.py 131 7104 1 2>&1 | tail -n 1 | grep -o "'eikonal_ms': np.float64([0-9.]*)"
done; done | tee gpurun_out/r2s30_ab.log
DSURF_B200_LIB=$PWD/scripts/ab/lib_nobounds.so timeout 300 python -m pytest tests/test_gpu_fim.py -m gpu -x -q 2>&1 | tail -n 2
DSURF_B200_LIB=$PWD/scripts/ab/lib_nobounds.so timeout 300 python bench.py --eikonal fim --no-both --steps 1 --warmup 1 --no-cpu --no-dispersion --no-calsurfg-e2e --lsmr-iters 0 2> /dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('nobounds bench', d['value'], d['coo']['digest'])"
