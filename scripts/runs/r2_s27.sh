# round 2, call 27 (8 GPUs): BASELINE configs[4] -- 2049^2 propagation grid, 32 periods x 1024 sources (32 768 sweeps,
# 524 288 rays, 528 392 unknowns), full outer loop on 8 B200 with the fast-iterative eikonal: every rank keeps its row block
# (never gathered), device glue, distributed LSMR with the peer-memory exchange, model update; 5 outer iterations
mkdir -p gpurun_out
DSURF_EIKONAL=fim timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 scripts/outer_loop_dist.py --nxy 259 --periods 32 --sources 1024 --iters 5 > gpurun_out/r2s27_cfg5_8gpu_fim.jsonl 2> gpurun_out/r2s27_cfg5_8gpu_fim.err
cat gpurun_out/r2s27_cfg5_8gpu_fim.jsonl | cut -c1-1000; tail -n 6 gpurun_out/r2s27_cfg5_8gpu_fim.err | cut -c1-300
