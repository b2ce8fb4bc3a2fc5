# round 2, call 14 (2 GPUs): multi-GPU pytest + the cfg-5 path (2049^2 grid, per-rank row blocks never gathered, device
# glue + distributed LSMR + model update in a loop), scaled down to 4 periods x 64 sources, 2 outer iterations
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/r2s14_pytest_multi.log 2>&1
tail -n 5 gpurun_out/r2s14_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 scripts/outer_loop_dist.py --nxy 259 --periods 4 --sources 64 --iters 2 > gpurun_out/r2s14_cfg5_scaled.jsonl 2> gpurun_out/r2s14_cfg5_scaled.err
cat gpurun_out/r2s14_cfg5_scaled.jsonl; tail -n 4 gpurun_out/r2s14_cfg5_scaled.err
