# round 2, call 15 (2 GPUs): peer-memory exchange of the distributed LSMR (CUDA IPC over NVLink, iteration in a CUDA
# graph) -- correctness against the single-GPU solve (dist_check), then the weak-scaling probe, peer vs NCCL
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555"
for w in taipei small; do timeout 600 $TR scripts/dist_check.py $w 2>&1 | grep -E "gather:|lsmr:|DIST_CHECK|Error|error" ; done | tee gpurun_out/r2s15_dist_check.log
timeout 300 python scripts/lsmr_bench.py --iters 60 2>&1 | tail -n 1 | tee gpurun_out/r2s15_lsmr_n1.json
timeout 400 $TR scripts/dist_lsmr_bench.py --iters 60 2>&1 | grep -E "^\{|rror" | tee gpurun_out/r2s15_lsmr_n2_peer.json
DSURF_LSMR_NCCL_ONLY=1 timeout 400 $TR scripts/dist_lsmr_bench.py --iters 60 2>&1 | grep -E "^\{|rror" | tee gpurun_out/r2s15_lsmr_n2_nccl.json
timeout 400 $TR scripts/dist_lsmr_bench.py --iters 60 --strong 2>&1 | grep -E "^\{|rror" | tee gpurun_out/r2s15_lsmr_n2_peer_strong.json
