#!/bin/bash
# scan_tc2 (filter scan on CTA pairs): parity first (short timeouts: a wrong barrier protocol hangs), then timing
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tcgen05_parity and 9" > gpurun_out/c14_pytest_a.log 2>&1; echo "parity[9] rc=$?"; tail -5 gpurun_out/c14_pytest_a.log | cut -c1-200
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "(fallback and 9) or c2_full_size_vector" > gpurun_out/c14_pytest_b.log 2>&1; echo "fallback/full rc=$?"; tail -5 gpurun_out/c14_pytest_b.log | cut -c1-200
for K in filt256 filt256p; do
  timeout 200 python bench.py --sections vector --cpu-seconds 0 --vector-kernel $K > gpurun_out/c14_$K.json 2> gpurun_out/c14_$K.err || tail -3 gpurun_out/c14_$K.err
done
python - <<'PY'
import json
for k in ("filt256", "filt256p"):
    try:
        d = json.load(open(f"gpurun_out/c14_{k}.json")); r = d["roofline"]
        print(k, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 4), r["kernel_ms"], r["frac"], list(d["kernels"].values())[0].get("filter_fallbacks"))
    except Exception as e: print(k, "parse", e)
PY
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
