#!/bin/bash
# 2 GPUs after the merge_lists change: world-2 test + bench vector / bm25 / parity
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/c25_pytest_multi.log 2>&1; echo "pytest multi rc=$? $(tail -1 gpurun_out/c25_pytest_multi.log)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --cpu-seconds 0 --sections vector,bm25,parity --vector-kernel filt256p > gpurun_out/c25_bench_n2.json 2> gpurun_out/c25_bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/c25_bench_n2.json") if l.startswith("{")][-1])
    print("n=2", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parity", d["parity_check"]["mismatches"], "of", d["parity_check"]["n"], "bm25", round(d["bm25"]["value"]))
except Exception as e: print("parse", e)
PY
