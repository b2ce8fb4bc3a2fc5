# round 2, call 28: final tree -- whole GPU suite + smoke
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2s28_pytest.log 2>&1
tail -n 6 gpurun_out/r2s28_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
