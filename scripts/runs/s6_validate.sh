# round-1 validation call: full GPU parity suite, default bench (our arm), short reference arm
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s6_pytest.log 2>&1
tail -n 6 gpurun_out/s6_pytest.log
timeout 900 python bench.py > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err
tail -c 600 gpurun_out/s6_bench.json; tail -n 5 gpurun_out/s6_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/s6_ref.json 2> gpurun_out/s6_ref.err
cat gpurun_out/s6_ref.json; tail -n 5 gpurun_out/s6_ref.err
