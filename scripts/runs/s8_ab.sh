# parity suite with the lazy-back-pointer eikonal + fused LSMR, then A/B timings
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s8_pytest.log 2>&1
tail -n 12 gpurun_out/s8_pytest.log
for v in fused nofuse nofork; do
  e=""; [ $v = nofuse ] && e="DSURF_LSMR_NO_FUSE=1"; [ $v = nofork ] && e="DSURF_LSMR_NO_FORK=1"
  env $e timeout 300 python scripts/lsmr_bench.py > gpurun_out/s8_lsmr_$v.json 2> gpurun_out/s8_lsmr_$v.err
  cat gpurun_out/s8_lsmr_$v.json; tail -n 2 gpurun_out/s8_lsmr_$v.err
done
for v in lazy eager lazy8; do
  e=""; [ $v = eager ] && e="DSURF_EIKONAL_EAGER=1"; [ $v = lazy8 ] && e="DSURF_EIKONAL_G=8"
  env $e timeout 300 python bench.py --step-mode type --steps 2 --warmup 1 --no-cpu --lsmr-iters 0 --no-dispersion > gpurun_out/s8_eik_$v.json 2> gpurun_out/s8_eik_$v.err
  python -c "import json;d=json.load(open('gpurun_out/s8_eik_$v.json'));print('$v',d['value'],d['ms_per_step'],d['stage_ms_per_step'])"; tail -n 2 gpurun_out/s8_eik_$v.err
done
