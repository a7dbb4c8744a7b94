#!/bin/bash
# GPU call 1 (round 2): full gpu test-suite incl. the new C3/C4 full-size identity tests on the round-1 kernels,
# plus the A/B of the never-run SSB_FFMA_GROUPMAX switch.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/c1_gpu.txt 2>&1
free -g > gpurun_out/c1_host.txt; nproc >> gpurun_out/c1_host.txt
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
# FFMA group-max seeding A/B: parity tests + batch sweep with the variant library
SSB_LIB=$PWD/seekstorm_b200/libseekstorm_b200_gm.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py -m gpu -q -x -k "vector_parity_small or vector_doc_ids or vector_paging" > gpurun_out/c1_gm_pytest.log 2>&1
echo "gm pytest rc=$?" >> gpurun_out/c1_gm_pytest.log
tail -3 gpurun_out/c1_gm_pytest.log
timeout 300 python bench.py --sections "" --vector-kernel ffma --cpu-seconds 0 --batch 16 --steps 20 > gpurun_out/c1_ffma_base.json 2> gpurun_out/c1_ffma_base.err
SSB_LIB=$PWD/seekstorm_b200/libseekstorm_b200_gm.so timeout 300 python bench.py --sections "" --vector-kernel ffma --cpu-seconds 0 --batch 16 --steps 20 > gpurun_out/c1_ffma_gm.json 2> gpurun_out/c1_ffma_gm.err
echo done
