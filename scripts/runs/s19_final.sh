mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/s19_pytest.log 2>&1
tail -n 6 gpurun_out/s19_pytest.log
timeout 300 python bench.py --step-mode type --steps 1 --warmup 1 --lsmr-iters 20 --no-dispersion > gpurun_out/s19_bench_type.json 2> gpurun_out/s19_bench_type.err
python -c "
import json
l=[x for x in open('gpurun_out/s19_bench_type.json').read().splitlines() if x.strip()]
print('stdout lines', len(l)); d=json.loads(l[0]); print(d['value'], d['e2e']['value'], d['lsmr']['iters_per_s'], d['rays'], d['cpu_baseline']['value'], d['roofline']['traffic'])"
tail -n 3 gpurun_out/s19_bench_type.err
