#!/bin/bash
# new rows: facet filters (lex_generic predicate) and TurboQuantI8 parity; ncu --set full of scan_tc2 (the headline kernel)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_filters.py -m gpu -q > gpurun_out/c16_pytest_filters.log 2>&1; echo "filters rc=$? $(tail -1 gpurun_out/c16_pytest_filters.log)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "turboquant or int8" > gpurun_out/c16_pytest_turbo.log 2>&1; echo "turbo rc=$? $(tail -1 gpurun_out/c16_pytest_turbo.log)"
grep -E "^(FAILED|ERROR)|Error|assert" gpurun_out/c16_pytest_filters.log | head -20
grep -E "^(FAILED|ERROR)" gpurun_out/c16_pytest_turbo.log | head -20
timeout 600 python -m pytest tests/test_gpu_abi.py tests/test_gpu_multifield.py tests/test_gpu_parity.py -m gpu -q -x -k "not turboquant and not int8" > gpurun_out/c16_pytest_regress.log 2>&1; echo "regress rc=$? $(tail -1 gpurun_out/c16_pytest_regress.log)"
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:scan_tc2 -s 2 -c 1 -f -o gpurun_out/r02_scan_tc2_filter_pair python tools/prof_driver.py filt256p > gpurun_out/c16_ncu_tc2.log 2>&1; echo "ncu tc2 rc=$?"; tail -2 gpurun_out/c16_ncu_tc2.log
