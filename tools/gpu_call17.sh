#!/bin/bash
# field filter (multi-field), phrase queries, regression of everything lexical (QueryPlan / lex_plan changed)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_phrase.py -m gpu -q > gpurun_out/c17_pytest_phrase.log 2>&1; echo "phrase rc=$? $(tail -1 gpurun_out/c17_pytest_phrase.log)"
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/c17_pytest_phrase.log | head -30
timeout 600 python -m pytest tests/test_gpu_multifield.py tests/test_gpu_filters.py -m gpu -q > gpurun_out/c17_pytest_filters.log 2>&1; echo "multifield+filters rc=$? $(tail -1 gpurun_out/c17_pytest_filters.log)"
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/c17_pytest_filters.log | head -30
timeout 900 python -m pytest tests/test_gpu_abi.py tests/test_gpu_parity.py tests/test_gpu_loader.py tests/test_cpp_mirror.py -m gpu -q -k "not turboquant and not int8 and not vector" > gpurun_out/c17_pytest_regress.log 2>&1; echo "regress rc=$? $(tail -1 gpurun_out/c17_pytest_regress.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/c17_pytest_regress.log | head -20
