#!/bin/bash
# Evidence pass B (round 2): default bench (N=1), the reference arm, launch list of a short bench run
mkdir -p gpurun_out
timeout 1800 python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_full.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "bench ref rc=$?"; tail -3 gpurun_out/r02_bench_reference.err
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --cpu-seconds 0 --sections vector,int8,bm25,hybrid > gpurun_out/c3_launch_bench.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import json
for f in ("r02_bench_full", "r02_bench_reference"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d.get("value"), d.get("e2e", {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("parity_check"))
        for k in ("bm25", "hybrid", "int8", "c5"):
            if k in d: print("  ", k, d[k].get("value"), d[k].get("e2e", {}).get("value") if isinstance(d[k].get("e2e"), dict) else None)
    except Exception as e: print(f, "parse", e)
PY
