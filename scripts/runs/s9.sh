# full parity suite (no -x) with fused LSMR + driver/raypath tests, LSMR probe A/B
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/s9_pytest.log 2>&1
tail -n 25 gpurun_out/s9_pytest.log
for v in fused nofuse; do
  e=""; [ $v = nofuse ] && e="DSURF_LSMR_NO_FUSE=1"
  env $e timeout 300 python scripts/lsmr_bench.py > gpurun_out/s9_lsmr_$v.json 2> gpurun_out/s9_lsmr_$v.err
  cat gpurun_out/s9_lsmr_$v.json; tail -n 2 gpurun_out/s9_lsmr_$v.err
done
