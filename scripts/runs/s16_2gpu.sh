mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/dist_lsmr_check.py > gpurun_out/s16_dist_check.log 2>&1
tail -n 3 gpurun_out/s16_dist_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 1 --warmup 1 > gpurun_out/s16_bench_2gpu.json 2> gpurun_out/s16_bench_2gpu.err
tail -c 1200 gpurun_out/s16_bench_2gpu.json; tail -n 4 gpurun_out/s16_bench_2gpu.err
