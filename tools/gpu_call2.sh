#!/bin/bash
# GPU call 2 (round 2): evidence pass — ncu --set full of the shipped kernels, launch list of a bench run, compute-sanitizer
# on the smoke shapes, full default bench (N=1).
mkdir -p gpurun_out
cat > /tmp/prof_driver.py <<'PY'
import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
from seekstorm_b200 import Index, QueryType, ResultType, VectorSimilarity, synth
what = sys.argv[1]
if what in ("tcb", "ffma", "i8"):
    n, d = 1_000_000, 768
    ix = Index(0, vector_dims=d, vector_similarity=VectorSimilarity.Cosine, max_batch=1024, vector_quantization=1 if what == "i8" else 0)
    for lv in range((n + 65535) // 65536):
        ix.add_vector_level(lv, synth.gen_vectors(min(65536, n - lv * 65536), d, 1002000 + lv, "cuda"))
    nq = {"tcb": 256, "ffma": 16, "i8": 1024}[what]
    q = synth.gen_vectors(nq, d, 2002, "cuda")
    keys = torch.zeros((nq, 32), dtype=torch.int64, device="cuda")
    ix.set_vector_kernel(int(os.environ.get("SSB_PROF_KERNEL", {"tcb": 4, "ffma": 1, "i8": 0}[what])))
    for _ in range(3):
        ix.search_vector_keys(q, 10, keys); torch.cuda.synchronize()
else:
    from seekstorm_b200._lib import SsbLexBatch
    n = 10_000_000
    ix = Index(0, max_batch=4096)
    ls = 0
    for lv in synth.gen_lexical_corpus(n, 1_000_000, 1003, "cuda"):
        ix.add_synth_level(lv); ls += lv.len_sum_normalized
    ix.commit(n, ls)
    qs = synth.gen_queries(4096, 2003, 20, 100000, (2, 3, 4), (0.4, 0.4, 0.2))
    qk = [[int(k) for k in synth.term_keys_np(np.array(q, dtype=np.int64))] for q in qs]
    qt = QueryType.Intersection if what == "lex_and" else QueryType.Union
    rt = ResultType.TopkCount if what == "lex_count" else ResultType.Topk
    b, keep = ix._lex_batch(qk, qt)
    keys = torch.zeros((4096, 32), dtype=torch.int64, device="cuda"); cnt = torch.zeros(4096, dtype=torch.int64, device="cuda")
    for _ in range(3):
        ix.search_lexical_keys(b, 10, rt, keys, cnt); torch.cuda.synchronize()
    print(ix.last_stats())
PY
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:scan_tc -s 4 -c 2 -f -o gpurun_out/r02_scan_tc_bf16 python /tmp/prof_driver.py tcb > gpurun_out/c2_ncu_tcb.log 2>&1; echo "ncu tcb rc=$?"
timeout 900 $NCU -k regex:lex_score -s 2 -c 1 -f -o gpurun_out/r02_lex_score python /tmp/prof_driver.py lex_or > gpurun_out/c2_ncu_lex.log 2>&1; echo "ncu lex rc=$?"
timeout 900 $NCU -k regex:lex_count -s 2 -c 1 -f -o gpurun_out/r02_lex_count python /tmp/prof_driver.py lex_count > gpurun_out/c2_ncu_lexc.log 2>&1; echo "ncu lexc rc=$?"
timeout 600 $NCU -k regex:scan_tc -s 4 -c 2 -f -o gpurun_out/r02_scan_tc_i8 python /tmp/prof_driver.py i8 > gpurun_out/c2_ncu_i8.log 2>&1; echo "ncu i8 rc=$?"
timeout 600 $NCU -k regex:scan_ffma -s 4 -c 2 -f -o gpurun_out/r02_scan_ffma python /tmp/prof_driver.py ffma > gpurun_out/c2_ncu_ffma.log 2>&1; echo "ncu ffma rc=$?"
# compute-sanitizer on the smoke shapes (one tiny invocation of every hot path)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r02_sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?" | tee -a gpurun_out/r02_sanitizer_synccheck.log
# full default bench + launch list
timeout 1500 python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err; echo "bench rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --cpu-seconds 0 --sections vector,int8,bm25,hybrid > gpurun_out/c2_launch_bench.log 2>&1; echo "launch list rc=$?"
