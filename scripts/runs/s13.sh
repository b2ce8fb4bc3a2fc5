mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python scripts/lsmr_bench.py > gpurun_out/s13_$name.json 2> gpurun_out/s13_$name.err; python - <<PY
import json
d=json.load(open('gpurun_out/s13_$name.json')); print('$name', round(d['iters_per_s']), round(d['us_per_iter'],1), 'spmv', round(d['spmv_us'],1), 'spmtv', round(d['spmtv_us'],1), d['x_checksum'])
PY
tail -n 1 gpurun_out/s13_$name.err; }
run default A=1
run w128 DSURF_LSMR_W128=1
run oldcols DSURF_LSMR_OLD_COLS=1
run oldcols_w128 DSURF_LSMR_OLD_COLS=1 DSURF_LSMR_W128=1
run oldboth DSURF_LSMR_OLD_COLS=1 DSURF_LSMR_OLD_ROWS=1
