#!/bin/bash
# GPU call 2a: remaining gpu tests (loader, C/C++ mirrors), ncu --set full of lex_score / lex_count / scan_tc variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_loader.py tests/test_cpp_mirror.py tests/test_gpu_multi.py -m gpu -q --durations=5 > gpurun_out/c2_pytest_rest.log 2>&1
echo "pytest rest rc=$?" | tee -a gpurun_out/c2_pytest_rest.log
tail -8 gpurun_out/c2_pytest_rest.log
sed -n '/^cat > \/tmp\/prof_driver.py/,/^PY$/p' tools/gpu_call2.sh > /tmp/mk_driver.sh; bash /tmp/mk_driver.sh
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:lex_score -s 2 -c 1 -f -o gpurun_out/r02_lex_score python /tmp/prof_driver.py lex_or > gpurun_out/c2_ncu_lex.log 2>&1; echo "ncu lex rc=$?"
timeout 900 $NCU -k regex:lex_count -s 2 -c 1 -f -o gpurun_out/r02_lex_count python /tmp/prof_driver.py lex_count > gpurun_out/c2_ncu_lexc.log 2>&1; echo "ncu lexc rc=$?"
timeout 600 $NCU -k regex:scan_tc -s 4 -c 2 -f -o gpurun_out/r02_scan_tc_bf16 python /tmp/prof_driver.py tcb > gpurun_out/c2_ncu_tcb.log 2>&1; echo "ncu tcb rc=$?"
SSB_PROF_KERNEL=6 timeout 600 $NCU -k regex:scan_tc -s 2 -c 2 -f -o gpurun_out/r02_scan_tc_bf16_n256 python /tmp/prof_driver.py tcb > gpurun_out/c2_ncu_tcb256.log 2>&1; echo "ncu tcb256 rc=$?"
SSB_LIB=$PWD/seekstorm_b200/libseekstorm_b200_gm.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py -m gpu -q -x -k "vector_parity_small or vector_doc_ids or vector_paging" > gpurun_out/c1_gm_pytest.log 2>&1
echo "gm pytest rc=$?" | tee -a gpurun_out/c1_gm_pytest.log
SSB_LIB=$PWD/seekstorm_b200/libseekstorm_b200_gm.so timeout 300 python bench.py --sections "" --vector-kernel ffma --cpu-seconds 0 --batch 16 --steps 20 > gpurun_out/c1_ffma_gm.json 2> gpurun_out/c1_ffma_gm.err
python - <<'PY'
import json
for n in ("base", "gm"):
    try:
        d = json.load(open(f"gpurun_out/c1_ffma_{n}.json")); print(n, d["value"], d["batch_sweep_e2e"], d["roofline"]["frac"])
    except Exception as e: print(n, "parse", e)
PY
