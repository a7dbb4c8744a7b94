#!/bin/bash
# Full GPU suite (filter scan, 12-warp epilogues, IVF probe), vector bench, ncu --set full of the filter kernels, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/r02_pytest_gpu_b.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest_gpu_b.log; tail -14 gpurun_out/r02_pytest_gpu_b.log
timeout 900 python bench.py --sections vector,parity --cpu-seconds 0 > gpurun_out/r02_bench_vec_b.json 2> gpurun_out/r02_bench_vec_b.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_vec_b.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_vec_b.json"))
print("headline", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["config"]["kernel"][:60], d.get("parity_check", {}).get("mismatches"))
for k, v in d["kernels"].items():
    r = v["roofline"]
    print(f"{k:26s} value {v['value']:10.0f} e2e {v['e2e']['value']:10.0f} ms/step {v['ms_per_step']:.3f} kern_ms {r['kernel_ms']:.3f} frac {r['frac']:.3f} fb {v.get('filter_fallbacks')}")
print(d["batch_sweep_e2e"])
PY
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:scan_tc -s 4 -c 2 -f -o gpurun_out/r02_scan_tc_filter_n256 python tools/prof_driver.py filt256 > gpurun_out/c7_ncu_f256.log 2>&1; echo "ncu filt256 rc=$?"
timeout 600 $NCU -k regex:scan_tc -s 4 -c 2 -f -o gpurun_out/r02_scan_tc_filter python tools/prof_driver.py filt > gpurun_out/c7_ncu_f.log 2>&1; echo "ncu filt rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_filter.csv python bench.py --steps 2 --warmup 1 --cpu-seconds 0 --sections vector --vector-kernel filt256 > gpurun_out/launch_filter.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02_launches_filter.csv")) if len(r) > 5]
hdr = None
for i, r in enumerate(rows):
    if "Kernel Name" in r: hdr = r; rows = rows[i + 1:]; break
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = [(r[ki][:70], float(r[vi].replace(",", ""))) for r in rows if r[vi].replace(",", "").replace(".", "").isdigit()]
for k, v in seq[-11:]: print(f"{v/1000:9.1f} us  {k}")
PY
