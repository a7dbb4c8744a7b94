#!/bin/bash
# new bench sections (int8 variants, facet-filter variants, phrase) + the new larger tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_phrase.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "phrase or config_errors or facet_filter" > gpurun_out/c19_pytest.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/c19_pytest.log)"
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/c19_pytest.log | head -20
timeout 900 python bench.py --sections vector,int8,bm25,phrase --vector-kernel filt256p --cpu-seconds 0 > gpurun_out/c19_bench.json 2> gpurun_out/c19_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/c19_bench.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/c19_bench.json") if l.startswith("{")][-1])
    print("int8 variants", json.dumps(d["int8"].get("variants"))[:1500])
    print("bm25 variants", json.dumps(d["bm25"]["variants"])[:1500])
    print("phrase", json.dumps(d.get("phrase"))[:1500])
except Exception as e: print("parse", e)
PY
