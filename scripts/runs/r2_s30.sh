# round 2, call 30: A/B -- marking without bounds tests (out-of-grid words flagged at tile load), stencil words reused by the
# marking loop, 5 / 6 / 7 resident CTAs per SM; then the cfg-3 COO digest of the variant builds (must equal 870cc4991b8bbf64)
# and the parity tests
mkdir -p gpurun_out
t() { echo -n "$1 minb=$2: "; DSURF_B200_LIB=$PWD/scripts/ab/lib_$1.so DSURF_FIM_MINB=$2 DSURF_EIKONAL=fim timeout 200 python scripts/profile_eikonal.py 131 7104 1 2>&1 | tail -n 1 | grep -o "'eikonal_ms': np.float64([0-9.]*)"; }
( t base 6; t nobounds 6; t nbreuse 6; t base 7; t nbreuse 7; t nbreuse 5 ) | tee gpurun_out/r2s30_ab.log
for v in nbreuse nobounds; do DSURF_B200_LIB=$PWD/scripts/ab/lib_$v.so timeout 300 python bench.py --eikonal fim --no-both --steps 1 --warmup 1 --no-cpu --no-dispersion --no-calsurfg-e2e --lsmr-iters 0 2> /dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v bench', d['value'], d['coo']['digest'])"; done | tee -a gpurun_out/r2s30_ab.log
DSURF_B200_LIB=$PWD/scripts/ab/lib_nbreuse.so timeout 300 python -m pytest tests/test_gpu_fim.py -m gpu -x -q 2>&1 | tail -n 2 | tee -a gpurun_out/r2s30_ab.log
