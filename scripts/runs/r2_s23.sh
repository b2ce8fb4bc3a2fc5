# round 2, call 23 (2 GPUs): the cfg-5 path with the fast-iterative eikonal at the PER-GPU load of the full configuration
# (2049^2 grid, 1024 sources x 8 periods = 8192 sweeps, 4096 per GPU -- cfg 5 on 8 GPUs is 32 periods, the same 4096 per GPU):
# device glue + distributed LSMR (peer exchange) + model update, 3 outer iterations, row blocks never gathered
mkdir -p gpurun_out
DSURF_EIKONAL=fim timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 scripts/outer_loop_dist.py --nxy 259 --periods 8 --sources 1024 --iters 3 > gpurun_out/r2s23_cfg5_pergpu_fim.jsonl 2> gpurun_out/r2s23_cfg5_pergpu_fim.err
cat gpurun_out/r2s23_cfg5_pergpu_fim.jsonl | cut -c1-900; tail -n 6 gpurun_out/r2s23_cfg5_pergpu_fim.err | cut -c1-300
