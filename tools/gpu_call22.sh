#!/bin/bash
# FINAL evidence pass of round 2: whole GPU suite, smoke, default bench (both arms), launch list, racecheck on smoke()
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final2.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02_pytest_gpu_final2.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/r02_pytest_gpu_final2.log | head -10
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final2.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/r02_smoke_final2.log)"
timeout 1200 python bench.py > gpurun_out/r02_bench_final2.json 2> gpurun_out/r02_bench_final2.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_final2.err | cut -c1-300
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm2.json 2> gpurun_out/r02_bench_reference_arm2.err; echo "ref arm rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02_bench_final2.json") if l.startswith("{")][-1])
    print("C2", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["config"]["kernel"][:40], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "parity", d.get("parity_check", {}).get("mismatches"))
    b = d["bm25"]; print("bm25", round(b["value"]), round(b["e2e"]["value"]), b["roofline"]["kernel_ms"], b["roofline"]["frac"], {k: (round(v["value"]), v["kernel_ms"]) for k, v in b["variants"].items() if "value" in v})
    print("int8", round(d["int8"]["value"]), {k: round(v["value"]) for k, v in d["int8"].get("variants", {}).items() if "value" in v}, "hybrid", round(d["hybrid"]["value"]), "c5", round(d["c5"]["value"]), "phrase", d.get("phrase", {}).get("topk", {}).get("value"), "cpu", d["cpu_baseline"]["value"])
except Exception as e: print("parse", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_final2.csv python bench.py --steps 3 --warmup 3 --cpu-seconds 0 --sections bm25 --bm25-docs 2000000 > gpurun_out/c22_launch_bench.log 2>&1; echo "launch list rc=$?"
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_racecheck_v3.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r02_sanitizer_racecheck_v3.log; tail -3 gpurun_out/r02_sanitizer_racecheck_v3.log
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
