#!/bin/bash
# affine Euclidean SQ parity + int8 regression (scan_tc signature changed), smoke() with the new paths, compute-sanitizer memcheck on smoke()
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c18_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/c18_smoke.log)"; grep -E "Error|assert" gpurun_out/c18_smoke.log | head -5
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "int8 or affine or turboquant" > gpurun_out/c18_pytest_int8.log 2>&1; echo "int8 rc=$? $(tail -1 gpurun_out/c18_pytest_int8.log)"
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/c18_pytest_int8.log | head -20
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck_v3.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r02_sanitizer_memcheck_v3.log; tail -4 gpurun_out/r02_sanitizer_memcheck_v3.log
