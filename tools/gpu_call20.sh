#!/bin/bash
# 2 GPUs at HEAD: world-2 test through the C-ABI + the bench's vector / bm25 / parity sections (what the driver's scaling run exercises)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r02_pytest_multi_n2_final.log 2>&1; echo "pytest multi rc=$? $(tail -1 gpurun_out/r02_pytest_multi_n2_final.log)"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --cpu-seconds 0 --sections vector,int8,bm25,hybrid,parity > gpurun_out/r02_bench_n2_final.json 2> gpurun_out/r02_bench_n2_final.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_n2_final.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02_bench_n2_final.json") if l.startswith("{")][-1])
    print("n=2", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["config"]["kernel"][:30], "parity", d["parity_check"]["mismatches"], "of", d["parity_check"]["n"])
    b = d["bm25"]; print("bm25", round(b["value"]), round(b["e2e"]["value"]), b["roofline"]["kernel_ms"], {k: round(v["value"]) for k, v in b["variants"].items()})
    print("int8", round(d["int8"]["value"]), "hybrid", round(d["hybrid"]["value"]))
except Exception as e: print("parse", e)
PY
