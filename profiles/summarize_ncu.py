#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python profiles/summarize_ncu.py rep  gpurun_out/prof_x.ncu-rep  > profiles/rNN_x.summary.txt
  python profiles/summarize_ncu.py list gpurun_out/launches.csv     > profiles/rNN_launches.summary.txt
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_active.avg",
    "sm__inst_executed_pipe_lsu.sum", "smsp__warps_eligible.avg.per_cycle_active",
    # round 2: spills / local memory, shared-memory conflicts, global request efficiency, launch shape
    "launch__local_size_bytes", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "sm__maximum_warps_per_active_cycle_pct",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "smsp__inst_executed_op_global_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
    "sm__inst_executed_pipe_uniform.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== kernel:", r[hdr.index("Kernel Name")][:100], "| grid", r[hdr.index("Grid Size")], "| block", r[hdr.index("Block Size")])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:75s} {r[hdr.index(k)]:>20s} {units[hdr.index(k)]}")
        st = []
        for i, h in enumerate(hdr):
            if "average_warp_latency_issue_stalled" in h or ("warp_issue_stalled" in h and h.endswith("_per_warp_active.pct")):
                try:
                    st.append((float(r[i].replace(",", "")), h))
                except ValueError:
                    pass
        for v, h in sorted(st, reverse=True)[:8]:
            print(f"  stall {h:85s} {v:10.3f}")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    d = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui].strip(), 1e-6)
        d[r[ki][:90]][0] += 1
        d[r[ki][:90]][1] += v * scale
    tot = sum(v[1] for v in d.values())
    print(f"# {path}: {sum(v[0] for v in d.values())} launches, {tot:.3f} ms total (ncu serialised, cold cache: compare SHARES)")
    for n, (c, t) in sorted(d.items(), key=lambda x: -x[1][1]):
        print(f"{t:12.3f} ms {100 * t / tot:6.2f}%  x{c:5d}  {n}")


if __name__ == "__main__":
    {"rep": rep, "list": launches}[sys.argv[1]](sys.argv[2])
