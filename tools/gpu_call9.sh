#!/bin/bash
# Evidence pass (round 2, second half): default bench N=1 + reference arm, ncu --set full of lex_score (source-level), sanitizers on the
# smoke shapes (filter scan / refine / fallback included), launch list of a short bench run
mkdir -p gpurun_out
timeout 1800 python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_full.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "bench ref rc=$?"
python - <<'PY'
import json
for f in ("r02_bench_full", "r02_bench_reference"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d.get("value"), d.get("e2e", {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("parity_check", {}).get("mismatches"))
        for k in ("bm25", "hybrid", "int8", "c5"):
            if k in d: print("  ", k, d[k].get("value"), d[k].get("e2e", {}).get("value") if isinstance(d[k].get("e2e"), dict) else None, d[k].get("ms_per_step"))
        print("  sweep", d.get("batch_sweep_e2e"))
    except Exception as e: print(f, "parse", e)
PY
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:lex_score -s 2 -c 1 -f -o gpurun_out/r02_lex_score_b python tools/prof_driver.py lex_or > gpurun_out/c9_ncu_lex.log 2>&1; echo "ncu lex rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r02_sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?" | tee -a gpurun_out/r02_sanitizer_synccheck.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --cpu-seconds 0 --sections vector,int8,bm25,hybrid > gpurun_out/c9_launch_bench.log 2>&1; echo "launch list rc=$?"
