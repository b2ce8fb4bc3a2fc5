mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -k "lsmr or aprod or glue or outer or driver" ) > gpurun_out/s11_pytest.log 2>&1
tail -n 8 gpurun_out/s11_pytest.log
for v in persist nopersist; do
  e=""; [ $v = nopersist ] && e="DSURF_LSMR_NO_PERSIST=1"
  env $e timeout 300 python scripts/lsmr_bench.py > gpurun_out/s11_lsmr_$v.json 2> gpurun_out/s11_lsmr_$v.err
  cat gpurun_out/s11_lsmr_$v.json; tail -n 2 gpurun_out/s11_lsmr_$v.err
done
