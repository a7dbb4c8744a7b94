#!/bin/bash
# Multi-GPU validation (run under `gpurun --gpus N`): the world-N test through the C-ABI, then bench.py at N GPUs with NCCL_DEBUG=INFO
N=${NGPU:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r02_pytest_multi_n$N.log 2>&1; echo "pytest multi rc=$?" | tee -a gpurun_out/r02_pytest_multi_n$N.log; tail -3 gpurun_out/r02_pytest_multi_n$N.log
for G in ${SWEEP:-$N}; do
  NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/nccl_n${G}_%p.log timeout 1800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $G --steps 10 --warmup 3 --cpu-seconds 0 > gpurun_out/r02_bench_n$G.json 2> gpurun_out/r02_bench_n$G.err; echo "bench n=$G rc=$?"; tail -2 gpurun_out/r02_bench_n$G.err
  cat gpurun_out/nccl_n${G}_*.log 2>/dev/null | grep -i "NVLS\|Connected all\|via P2P\|Channel 00/\|comm .* rank 0 " | sort | uniq -c | sort -rn | head -12 > gpurun_out/r02_nccl_n$G.summary.txt
  rm -f gpurun_out/nccl_n${G}_*.log
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_bench_n$G.json"))
    print("n=$G", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parity", d["parity_check"]["mismatches"], "of", d["parity_check"]["n"])
    for k in ("int8", "bm25", "hybrid", "c5"):
        if k in d: print("   ", k, round(d[k]["value"]), (d[k].get("config") or {}).get("workload", "")[:90])
except Exception as e: print("n=$G parse", e)
PY
done
head -12 gpurun_out/r02_nccl_n$N.summary.txt
