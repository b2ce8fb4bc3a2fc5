# round 2, call 22 (4 GPUs): smoke; sharded stage + NCCL gather + distributed LSMR (peer-memory exchange) against the
# single-GPU results; bench with both eikonal pipelines at N = 4 (cfg 2); LSMR weak-scaling probe, peer vs NCCL
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29566"
timeout 600 $TR scripts/dist_check.py small 2>&1 | grep -E "gather:|lsmr:|DIST_CHECK|Error|error" | tee gpurun_out/r2s22_dist_check.log
timeout 400 $TR bench.py --gpus 4 --config 2 --steps 2 --warmup 1 --no-cpu > gpurun_out/r2s22_bench_cfg2_n4.json 2> gpurun_out/r2s22_bench_cfg2_n4.err
tail -n 3 gpurun_out/r2s22_bench_cfg2_n4.err | cut -c1-300
timeout 400 $TR scripts/dist_lsmr_bench.py --iters 60 2>&1 | grep -E "^\{|rror" | tee gpurun_out/r2s22_lsmr_n4_peer.json
DSURF_LSMR_NCCL_ONLY=1 timeout 400 $TR scripts/dist_lsmr_bench.py --iters 60 2>&1 | grep -E "^\{|rror" | tee gpurun_out/r2s22_lsmr_n4_nccl.json
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2s22_bench_cfg2_n4.json"))
    print(d["eikonal_pipeline"], {k:d[k] for k in ("value","ms_per_step","n_gpus","scaling")}, d["coo"], "e2e", d["e2e"]["value"])
    o=d.get("other_pipeline"); print("  other", o["pipeline"], o["value"], o["coo"]) if o else None
    print("  lsmr", d["lsmr"]["iters_per_s"], d["lsmr"]["per_rank"], d["lsmr"]["to_convergence"])
except Exception as e:
    print("ERR", e)
PY
