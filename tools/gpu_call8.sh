#!/bin/bash
# IVF probe + multi-field BM25F parity, then the whole GPU suite (everything behind the first failure of call 7 did not run)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ivf.py tests/test_gpu_multifield.py -m gpu -q > gpurun_out/r02_pytest_ivf_mf.log 2>&1; echo "ivf/mf rc=$?"; tail -25 gpurun_out/r02_pytest_ivf_mf.log | cut -c1-220
timeout 1500 python -m pytest tests -m gpu -q --durations=5 --deselect tests/test_gpu_ivf.py --deselect tests/test_gpu_multifield.py > gpurun_out/r02_pytest_gpu_c.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest_gpu_c.log; tail -12 gpurun_out/r02_pytest_gpu_c.log | cut -c1-220
