# round 2, call 12 (2 GPUs): sharded sweep stage + NCCL gather == single GPU bit for bit; distributed LSMR on the row
# partition; bench at N = 1 and N = 2 (cfg 2 digests must agree), then cfg 3 at N = 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
for w in taipei small; do timeout 600 $TR scripts/dist_check.py $w 2>&1 | grep -E "gather:|lsmr:|DIST_CHECK|Error|error" ; done | tee gpurun_out/r2s12_dist_check.log
timeout 300 python bench.py --config 2 --steps 2 --warmup 1 --no-cpu --no-calsurfg-e2e > gpurun_out/r2s12_bench_cfg2_n1.json 2> gpurun_out/r2s12_bench_cfg2_n1.err
timeout 300 $TR bench.py --gpus 2 --config 2 --steps 2 --warmup 1 --no-cpu > gpurun_out/r2s12_bench_cfg2_n2.json 2> gpurun_out/r2s12_bench_cfg2_n2.err
tail -n 3 gpurun_out/r2s12_bench_cfg2_n2.err
timeout 900 $TR bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2s12_bench_cfg3_n2.json 2> gpurun_out/r2s12_bench_cfg3_n2.err
tail -n 3 gpurun_out/r2s12_bench_cfg3_n2.err
python - <<'PY'
import json
for f in ("cfg2_n1","cfg2_n2","cfg3_n2"):
    try:
        d=json.load(open(f"gpurun_out/r2s12_bench_{f}.json"))
        print(f, {k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["coo"], d["stage_ms_per_step"], "e2e", d["e2e"]["value"])
        print("   lsmr", {k:d["lsmr"][k] for k in ("iters_per_s","nnz","m","per_rank","to_convergence")} if d.get("lsmr") else None, d["lsmr"]["roofline"]["frac"] if d.get("lsmr") else None)
    except Exception as e:
        print(f, "ERR", e)
PY
