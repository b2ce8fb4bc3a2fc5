# round 2, call 31: final build (global proxy fence): fast-iterative parity tests + cfg-3 COO digest
timeout 100 python -m pytest tests/test_gpu_fim.py -m gpu -x -q 2>&1 | tail -n 2
timeout 100 python bench.py --eikonal fim --no-both --steps 1 --warmup 1 --no-cpu --no-dispersion --no-calsurfg-e2e --lsmr-iters 0 2> /dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('final bench', d['value'], d['e2e']['value'], d['coo']['digest'])"
