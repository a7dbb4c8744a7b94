#!/bin/bash
# lex_score L1-policy A/B (streams without L1 allocation = main; alloc = old; keep = evict_last on the coarse bytes) + vector sanity
mkdir -p gpurun_out
timeout 300 python bench.py --sections vector --cpu-seconds 0 --vector-kernel filt256 > gpurun_out/c10_vec.json 2> gpurun_out/c10_vec.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/c10_vec.json"))
print("filt256", round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["roofline"]["kernel_ms"], d["kernels"]["scan_tc_f16_filter_n256"].get("filter_fallbacks"), d["batch_sweep_e2e"])
PY
VARIANTS="main alloc keep" bash tools/gpu_variants.sh
SSB_LIB=$PWD/seekstorm_b200/libseekstorm_b200.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py -m gpu -q -x -k "lex or bm25 or hybrid or delete or not_ or many or paging or stats or count" > gpurun_out/c10_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c10_pytest.log)"
