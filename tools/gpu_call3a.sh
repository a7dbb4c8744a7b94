#!/bin/bash
# Evidence pass A (round 2): the whole GPU test suite, compute-sanitizer on the smoke shapes, ncu --set full of every shipped kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" | tee -a gpurun_out/r02_pytest_gpu.log; tail -14 gpurun_out/r02_pytest_gpu.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r02_sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?" | tee -a gpurun_out/r02_sanitizer_synccheck.log
sed -n '/^cat > \/tmp\/prof_driver.py/,/^PY$/p' tools/gpu_call2.sh > /tmp/mk_driver.sh; bash /tmp/mk_driver.sh
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:lex_score -s 2 -c 1 -f -o gpurun_out/r02_lex_score_v6 python /tmp/prof_driver.py lex_or > gpurun_out/c3_ncu_lex.log 2>&1; echo "ncu lex rc=$?"
timeout 900 $NCU -k regex:lex_score -s 2 -c 1 -f -o gpurun_out/r02_lex_score_and_v6 python /tmp/prof_driver.py lex_and > gpurun_out/c3_ncu_lexa.log 2>&1; echo "ncu lex and rc=$?"
timeout 900 $NCU -k regex:lex_count -s 2 -c 1 -f -o gpurun_out/r02_lex_count_v2 python /tmp/prof_driver.py lex_count > gpurun_out/c3_ncu_lexc.log 2>&1; echo "ncu lexc rc=$?"
timeout 600 $NCU -k regex:scan_tc -s 4 -c 2 -f -o gpurun_out/r02_scan_tc_i8 python /tmp/prof_driver.py i8 > gpurun_out/c3_ncu_i8.log 2>&1; echo "ncu i8 rc=$?"
timeout 600 $NCU -k regex:scan_ffma -s 2 -c 2 -f -o gpurun_out/r02_scan_ffma python /tmp/prof_driver.py ffma > gpurun_out/c3_ncu_ffma.log 2>&1; echo "ncu ffma rc=$?"
