#!/bin/bash
# lex_score: coarse bytes of all terms requested before use (main) vs one term at a time (serial)
mkdir -p gpurun_out
VARIANTS="main serial main serial" bash tools/gpu_variants.sh
SSB_LIB=$PWD/seekstorm_b200/libseekstorm_b200.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py -m gpu -q -x -k "lex or bm25 or hybrid or delete or not_ or many or paging or stats or count" > gpurun_out/c13_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c13_pytest.log)"
