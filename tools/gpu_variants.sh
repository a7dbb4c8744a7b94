#!/bin/bash
# bm25 bench over experiment builds: VARIANTS="main nopf m4 ..." (libseekstorm_b200_<name>.so; main = the shipped library)
mkdir -p gpurun_out
for V in ${VARIANTS:-main}; do
  L=$PWD/seekstorm_b200/libseekstorm_b200_$V.so; [ "$V" = main ] && L=$PWD/seekstorm_b200/libseekstorm_b200.so
  if [ -n "$PYTEST" ]; then
    SSB_LIB=$L timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py -m gpu -q -x -k "lex or bm25 or hybrid or delete or not_ or many or paging or stats or count" > gpurun_out/var_${V}_pytest.log 2>&1
    echo "$V pytest rc=$? $(tail -1 gpurun_out/var_${V}_pytest.log)"
  fi
  SSB_LIB=$L timeout 400 python bench.py --sections bm25 --rows 65536 --cpu-seconds 0 --steps 8 > gpurun_out/var_$V.json 2> gpurun_out/var_$V.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/var_$V.json")); b = d.get("bm25") or d
    r = b["roofline"]
    print("$V", round(b["value"]), {k: round(x["value"]) for k, x in b["variants"].items()}, "ms", r["kernel_ms"], "visited", r["postings_visited"], "probes", r["probes"])
except Exception as e: print("$V parse", e)
PY
done
