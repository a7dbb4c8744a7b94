#!/bin/bash
# 2 GPUs: threshold exchange of the sharded BM25 path — world-2 test through the C-ABI, bm25 + parity sections of the bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r02_pytest_multi_n2b.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/r02_pytest_multi_n2b.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --cpu-seconds 0 --sections vector,bm25,parity > gpurun_out/r02_bench_n2b.json 2> gpurun_out/r02_bench_n2b.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_n2b.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_n2b.json"))
print("n=2", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parity", d["parity_check"]["mismatches"], "of", d["parity_check"]["n"])
b = d["bm25"]; print("bm25", round(b["value"]), round(b["e2e"]["value"]), b["roofline"]["kernel_ms"], {k: (round(v["value"]), v["kernel_ms"]) for k, v in b["variants"].items()}, b["roofline"]["postings_visited"])
PY
