#!/bin/bash
# GPU call (round 2): new BM25 engine (records + multi-level items + vectorised stream) — memcheck on small shapes first,
# then the full gpu suite incl. the C3/C4 full-size identity tests, then a bm25-only bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/c1_gpu.txt 2>&1
free -g > gpurun_out/c1_host.txt; nproc >> gpurun_out/c1_host.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lexical_hand_corpus or lexical_reference_fixture or device_pointers" > gpurun_out/c1_memcheck_lex.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/c1_memcheck_lex.log
tail -4 gpurun_out/c1_memcheck_lex.log
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py -m gpu -q -x --durations=10 > gpurun_out/c1_pytest_small.log 2>&1
echo "pytest small rc=$?" | tee -a gpurun_out/c1_pytest_small.log
tail -15 gpurun_out/c1_pytest_small.log
timeout 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_multi.py -m gpu -q -x --durations=10 > gpurun_out/c1_pytest_full.log 2>&1
echo "pytest full rc=$?" | tee -a gpurun_out/c1_pytest_full.log
tail -15 gpurun_out/c1_pytest_full.log
timeout 600 python bench.py --sections bm25 --rows 65536 --cpu-seconds 0 --steps 10 > gpurun_out/c1_bench_bm25.json 2> gpurun_out/c1_bench_bm25.err
echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/c1_bench_bm25.json'))
    b=d['bm25']; print('bm25', b.get('value'), b.get('e2e'), b.get('variants'), b.get('roofline'))
except Exception as e: print('bench parse', e)
PY
# FFMA group-max seeding A/B (compile-time switch SSB_FFMA_GROUPMAX=1 in libseekstorm_b200_gm.so): parity + batch-1/16 latency
SSB_LIB=$PWD/seekstorm_b200/libseekstorm_b200_gm.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py -m gpu -q -x -k "vector_parity_small or vector_doc_ids or vector_paging" > gpurun_out/c1_gm_pytest.log 2>&1
echo "gm pytest rc=$?" | tee -a gpurun_out/c1_gm_pytest.log
timeout 300 python bench.py --sections "" --vector-kernel ffma --cpu-seconds 0 --batch 16 --steps 20 > gpurun_out/c1_ffma_base.json 2> gpurun_out/c1_ffma_base.err
SSB_LIB=$PWD/seekstorm_b200/libseekstorm_b200_gm.so timeout 300 python bench.py --sections "" --vector-kernel ffma --cpu-seconds 0 --batch 16 --steps 20 > gpurun_out/c1_ffma_gm.json 2> gpurun_out/c1_ffma_gm.err
python - <<'PY'
import json
for n in ("base", "gm"):
    try:
        d = json.load(open(f"gpurun_out/c1_ffma_{n}.json")); print(n, d["value"], d["batch_sweep_e2e"], d["roofline"]["frac"])
    except Exception as e: print(n, "parse", e)
PY
# item-shape sweep of the BM25 engine (env knobs read once per process)
for W in 1024 2048 8192 16384; do
  SSB_LEX_ITEM_W=$W timeout 400 python bench.py --sections bm25 --rows 65536 --cpu-seconds 0 --steps 6 > gpurun_out/c1_bm25_w$W.json 2> gpurun_out/c1_bm25_w$W.err
done
for G in 2 4; do
  SSB_LEX_GRID=$G timeout 400 python bench.py --sections bm25 --rows 65536 --cpu-seconds 0 --steps 6 > gpurun_out/c1_bm25_g$G.json 2> gpurun_out/c1_bm25_g$G.err
done
SSB_LEX_FIRST=1 timeout 400 python bench.py --sections bm25 --rows 65536 --cpu-seconds 0 --steps 6 > gpurun_out/c1_bm25_f1.json 2> gpurun_out/c1_bm25_f1.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c1_bm25_*.json")):
    try:
        b = json.load(open(f))["bm25"]; print(f, round(b["value"]), {k: round(v["value"]) for k, v in b["variants"].items()}, b["roofline"]["kernel_ms"])
    except Exception as e: print(f, "parse", e)
PY
# vector kernels A/B: FP32, bf16 128-query tile, bf16 256-query tile
timeout 400 python bench.py --sections "" --vector-kernel all --cpu-seconds 0 --steps 20 > gpurun_out/c1_vec_all.json 2> gpurun_out/c1_vec_all.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c1_vec_all.json"))
    for k, v in d["kernels"].items(): print(k, round(v["value"]), round(v["e2e"]["value"]), v["roofline"]["kernel_ms"], v["roofline"]["frac"])
    print("sweep", d.get("batch_sweep_e2e"))
except Exception as e: print("vec parse", e)
PY
