#!/bin/bash
# Filter scan, 12 epilogue warps: parity tests, sample-size A/B, launch list of one filter step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py tests/test_gpu_fullsize.py -m gpu -x -q -k "tcgen05_parity or filter or stats_and_kernel or c2_full_size_vector or delete_set or hybrid_parity or multi_chunk or threshold or int8" > gpurun_out/r02_pytest_filter.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_pytest_filter.log
for T in 1 4 8; do
  for K in filt filt256; do
    SSB_TC_SAMPLE_TILES=$T timeout 300 python bench.py --sections vector --cpu-seconds 0 --vector-kernel $K > gpurun_out/ab_${K}_t$T.json 2> gpurun_out/ab_${K}_t$T.err || tail -3 gpurun_out/ab_${K}_t$T.err
  done
done
timeout 300 python bench.py --sections vector --cpu-seconds 0 --vector-kernel tcb > gpurun_out/ab_tcb.json 2> gpurun_out/ab_tcb.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f"{f:36s} value {d['value']:10.0f} e2e {d['e2e']['value']:10.0f} ms/step {d['ms_per_step']:.3f} kern_ms {r['kernel_ms']:.3f} frac {r['frac']:.3f}")
    except Exception as e: print(f, "parse", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_filter.csv python bench.py --steps 2 --warmup 1 --cpu-seconds 0 --sections vector --vector-kernel filt > gpurun_out/launch_filter.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_launches_filter.csv")) if len(r) > 5]
hdr = None
for i, r in enumerate(rows):
    if "Kernel Name" in r: hdr = r; rows = rows[i + 1:]; break
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = [(r[ki][:70], float(r[vi].replace(",", ""))) for r in rows if r[vi].replace(",", "").replace(".", "").isdigit()]
# last 12 launches = one step
for k, v in seq[-14:]: print(f"{v/1000:9.1f} us  {k}")
PY
