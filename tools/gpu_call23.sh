#!/bin/bash
# merge_lists heads-first: every vector test + the vector / int8 / hybrid bench sections
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py tests/test_gpu_ivf.py tests/test_gpu_fullsize.py tests/test_gpu_loader.py -m gpu -q -k "vector or hybrid or ivf or int8 or turboquant or affine or c2 or c4 or kernel or stream or delete or mirror" > gpurun_out/c23_pytest.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/c23_pytest.log)"
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/c23_pytest.log | head -20
timeout 600 python bench.py --sections vector,int8,hybrid,parity --cpu-seconds 0 > gpurun_out/c23_bench.json 2> gpurun_out/c23_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/c23_bench.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/c23_bench.json") if l.startswith("{")][-1])
    print("C2", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["ms_per_step"], d["clocks"]["sm_mhz"], "parity", d.get("parity_check", {}).get("mismatches"))
    print({k: (round(v["value"]), v["roofline"]["kernel_ms"]) for k, v in d["kernels"].items()})
    print("int8", round(d["int8"]["value"]), {k: round(v["value"]) for k, v in d["int8"].get("variants", {}).items() if "value" in v}, "hybrid", round(d["hybrid"]["value"]), d["batch_sweep_e2e"])
except Exception as e: print("parse", e)
PY
