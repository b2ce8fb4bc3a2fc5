# round-1 final validation: full GPU parity suite, default bench (both arms)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/s14_pytest.log 2>&1
tail -n 6 gpurun_out/s14_pytest.log
timeout 900 python bench.py > gpurun_out/s14_bench.json 2> gpurun_out/s14_bench.err
tail -c 1500 gpurun_out/s14_bench.json; tail -n 5 gpurun_out/s14_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/s14_ref.json 2> gpurun_out/s14_ref.err
cat gpurun_out/s14_ref.json; tail -n 5 gpurun_out/s14_ref.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s14_smoke.log 2>&1; tail -n 3 gpurun_out/s14_smoke.log
