mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python scripts/lsmr_bench.py > gpurun_out/s12_$name.json 2> gpurun_out/s12_$name.err; python - <<PY
import json
d=json.load(open('gpurun_out/s12_$name.json')); print('$name', round(d['iters_per_s']), round(d['us_per_iter'],1), 'spmv', round(d['spmv_us'],1), 'spmtv', round(d['spmtv_us'],1), d['x_checksum'])
PY
tail -n 1 gpurun_out/s12_$name.err; }
run default A=1
run rows0 DSURF_LSMR_GRID_ROWS=0
run rows8 DSURF_LSMR_GRID_ROWS=8
run cols4 DSURF_LSMR_GRID_COLS=4
run cols8 DSURF_LSMR_GRID_COLS=8
run old DSURF_LSMR_OLD_KERNELS=1
( timeout 600 python -m pytest tests -m gpu -q -k "lsmr or aprod or glue or outer" ) > gpurun_out/s12_pytest.log 2>&1
tail -n 4 gpurun_out/s12_pytest.log
