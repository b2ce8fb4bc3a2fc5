# round 2, call 29 (2 GPUs): the multi-GPU pytest module with the fast-iterative case
mkdir -p gpurun_out
( time timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/r2s29_pytest_multi.log 2>&1
tail -n 5 gpurun_out/r2s29_pytest_multi.log
