mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "sweep_bit_exact or calsurfg or variants or full_size or heap_slab" ) > gpurun_out/s15_pytest.log 2>&1
tail -n 4 gpurun_out/s15_pytest.log
timeout 300 python bench.py --step-mode type --steps 2 --warmup 1 --no-cpu --lsmr-iters 0 --no-dispersion > gpurun_out/s15_eik.json 2> gpurun_out/s15_eik.err
python -c "import json;d=json.load(open('gpurun_out/s15_eik.json'));print('posthoc+lds128',d['value'],d['ms_per_step'],d['stage_ms_per_step'])"; tail -n 2 gpurun_out/s15_eik.err
