#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_loader.py tests/test_gpu_zloader_phrase.py -m gpu -q > gpurun_out/c26_pytest_loader.log 2>&1; echo "loader rc=$? $(tail -1 gpurun_out/c26_pytest_loader.log)"
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/c26_pytest_loader.log | head -10
