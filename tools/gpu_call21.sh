#!/bin/bash
# facet-filtered Topk queries on the record path (lex_score<.., HAS_NOT>): parity + the bench variant
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_filters.py tests/test_gpu_phrase.py tests/test_gpu_multifield.py tests/test_gpu_abi.py tests/test_gpu_fullsize.py -m gpu -q -k "filter or phrase or not_lists or delete or paging or multifield" > gpurun_out/c21_pytest.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/c21_pytest.log)"
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/c21_pytest.log | head -20
timeout 600 python bench.py --sections bm25 --vector-kernel filt256p --cpu-seconds 0 --steps 10 > gpurun_out/c21_bench.json 2> gpurun_out/c21_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/c21_bench.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/c21_bench.json") if l.startswith("{")][-1])
    b = d["bm25"]; print("bm25", round(b["value"]), b["roofline"]["kernel_ms"], json.dumps(b["variants"])[:1200])
except Exception as e: print("parse", e)
PY
