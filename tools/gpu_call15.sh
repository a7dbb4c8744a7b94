#!/bin/bash
# Evidence pass at HEAD (round 2, final state): whole GPU suite, default bench (both arms), ncu --set full of scan_tc2, launch list
mkdir -p gpurun_out
cp tools/prof_driver.py /tmp/prof_driver.py
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02_pytest_gpu_final.log)"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/r02_smoke.log)"
timeout 1200 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_final.err | cut -c1-300
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref arm rc=$?"
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:scan_tc2 -s 2 -c 1 -f -o gpurun_out/r02_scan_tc2_filter_pair python /tmp/prof_driver.py filt256p > gpurun_out/c15_ncu_tc2.log 2>&1; echo "ncu tc2 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 3 --warmup 3 --cpu-seconds 0 --sections vector,int8,bm25,hybrid --vector-kernel filt256p > gpurun_out/c15_launch_bench.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02_bench_final.json") if l.startswith("{")][-1])
    print("C2", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["config"]["kernel"][:40], "frac", d["roofline"]["frac"], "parity", d.get("parity_check", {}).get("mismatches"))
    print({k: (round(v["value"]), v["roofline"]["kernel_ms"], round(v["roofline"]["frac"], 3)) for k, v in d["kernels"].items()})
    b = d["bm25"]; print("bm25", round(b["value"]), round(b["e2e"]["value"]), b["roofline"]["kernel_ms"], b["roofline"]["frac"], {k: (round(v["value"]), v["kernel_ms"]) for k, v in b["variants"].items()})
    print("int8", round(d["int8"]["value"]), "hybrid", round(d["hybrid"]["value"]), "c5", round(d["c5"]["value"]), "cpu", d["cpu_baseline"]["value"])
except Exception as e: print("parse", e)
PY
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
