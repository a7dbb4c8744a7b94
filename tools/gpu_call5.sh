#!/bin/bash
# Filter vector scan (fp16 plane + exact f32 refine): parity tests, then the vector section of the bench with every kernel side by side
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py tests/test_gpu_fullsize.py -m gpu -x -q -k "tcgen05_parity or filter or stats_and_kernel or c2_full_size_vector or delete_set or hybrid_parity or multi_chunk or threshold" > gpurun_out/r02_pytest_filter.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_pytest_filter.log
timeout 900 python bench.py --sections vector,parity --cpu-seconds 0 > gpurun_out/r02_bench_vec_filter.json 2> gpurun_out/r02_bench_vec_filter.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_vec_filter.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_vec_filter.json"))
print("headline", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["config"]["kernel"][:60], d.get("parity_check", {}).get("mismatches"))
for k, v in d["kernels"].items():
    r = v["roofline"]
    print(f"{k:26s} value {v['value']:10.0f} e2e {v['e2e']['value']:10.0f} ms/step {v['ms_per_step']:.3f} kern_ms {r['kernel_ms']:.3f} frac {r['frac']:.3f} f32eq {r.get('f32_equivalent_gbs')} fb {v.get('filter_fallbacks')}")
print(d["batch_sweep_e2e"])
PY
