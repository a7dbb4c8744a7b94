#!/bin/bash
# BM25 iteration call: lexical parity tests, C3 full-size identity, bm25 bench, optional ncu of lex_score (NCU=1)
mkdir -p gpurun_out
T=${TAG:-v6}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py -m gpu -q -x -k "lex or bm25 or hybrid or delete or not_ or many or paging or stats or count" > gpurun_out/${T}_pytest.log 2>&1
echo "pytest lex rc=$?"; tail -4 gpurun_out/${T}_pytest.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "c3 or c4" > gpurun_out/${T}_pytest_full.log 2>&1
echo "pytest full rc=$?"; tail -4 gpurun_out/${T}_pytest_full.log
timeout 600 python bench.py --sections bm25 --rows 65536 --cpu-seconds 0 --steps 10 > gpurun_out/${T}_bm25.json 2> gpurun_out/${T}_bm25.err
echo "bench rc=$?"; tail -3 gpurun_out/${T}_bm25.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bm25.json"))
    b = d.get("bm25") or d
    print(json.dumps({k: b[k] for k in b if k in ("value","e2e","variants","roofline")})[:1500])
except Exception as e: print("parse", e)
PY
if [ -n "$NCU" ]; then
  sed -n '/^cat > \/tmp\/prof_driver.py/,/^PY$/p' tools/gpu_call2.sh > /tmp/mk_driver.sh; bash /tmp/mk_driver.sh
  case "$NCU" in *score*|1) timeout 900 ncu --set full --clock-control none --import-source on -k regex:lex_score -s 2 -c 1 -f -o gpurun_out/r02_lex_score_${T} python /tmp/prof_driver.py lex_or > gpurun_out/${T}_ncu_lex.log 2>&1; echo "ncu lex rc=$?";; esac
  case "$NCU" in *count*) timeout 900 ncu --set full --clock-control none --import-source on -k regex:lex_count -s 2 -c 1 -f -o gpurun_out/r02_lex_count_${T} python /tmp/prof_driver.py lex_count > gpurun_out/${T}_ncu_lexc.log 2>&1; echo "ncu lexc rc=$?";; esac
fi
