#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the seekstorm_b200 hot path.

Metric (BASELINE.json): queries/sec at top-10.  N=1 workload = configs[1]: brute-force cosine kNN over
1M x 768 f32 (C2).  A "step" = one call of the hot path over one batch of synthetic queries (batch = --batch
queries = batch/16 corpus passes).  `value` = device-resident QPS (queries already in HBM, packed keys left
in HBM); `e2e` = the same through the reference-facing C-ABI call ssb_search_vector with HOST buffers (H2D of
the queries and D2H of the results inside the timed region).  A second section ("bm25") measures C3
(BM25 OR top-10 over a 10M-doc Zipfian index) the same way.

N>1 (torchrun, one rank per GPU): the corpus is sharded by contiguous 64K-row level ranges (strong scaling);
every rank scans its shard for the whole batch, then one NCCL all-gather of the packed top-k keys and a
G*k -> k merge (seekstorm_b200/parallel.py).

--impl reference: times the CPU restatement of the reference path (oracle/, kind "port": the Rust reference
cannot be built here) on the host cores for the same metric / config.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

C2_ROWS, C2_DIMS, TOPK = 1_000_000, 768, 10
C3_DOCS, C3_VOCAB = 10_000_000, 1_000_000


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--batch", type=int, default=256, help="vector queries per step")
    p.add_argument("--rows", type=int, default=C2_ROWS)
    p.add_argument("--dims", type=int, default=C2_DIMS)
    p.add_argument("--sections", default="vector,int8,bm25,hybrid")
    p.add_argument("--int8-batch", type=int, default=1024, help="queries per step of the int8 (ScalarQuantizationI8) section")
    p.add_argument("--bm25-docs", type=int, default=C3_DOCS)
    p.add_argument("--bm25-batch", type=int, default=4096, help="lexical queries per step")
    p.add_argument("--hybrid-docs", type=int, default=5_000_000)
    p.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of each cpu_baseline sample")
    p.add_argument("--vector-kernel", default="both", choices=["both", "ffma", "tc", "tc64", "tcb", "tcb64"],
                   help="FP32 FFMA2 scan, tcgen05 3xTF32 scan (128 / 64 queries per pass) or both (headline = the faster)")
    return p.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  In-process NVML polling
    (nvidia_ml_py) every 10 ms; falls back to an `nvidia-smi -lms` child process.  (The first version polled nvidia-smi with
    power.draw in the query: each sample stalled kernel launches for milliseconds and the device-resident `value`, measured
    with the sampler running, came out slower than the e2e number measured without it.)"""
    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, gpu_index: int):
        self.sm, self.mx, self.reasons = [], [], set()
        self._stop = False
        self.p = None
        self.th = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(gpu_index).uuid))
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def poll():
                while not self._stop:
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.mx.append(mx)
                        r = int(get_reasons(h))
                        for name, bit in self.REASONS:
                            if r & bit:
                                self.reasons.add(name)
                    except Exception:
                        pass
                    time.sleep(0.01)
            self.th = threading.Thread(target=poll, daemon=True)
            self.th.start()
        except Exception:
            self._start_smi(gpu_index)

    def _start_smi(self, gpu_index):
        q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def has_samples(self):
        if self.th is not None:
            return len(self.sm) > 0
        try:
            return self.p is None or os.path.getsize(self.f.name) > 0
        except OSError:
            return True

    def stop(self):
        if self.th is not None:
            self._stop = True
            self.th.join(timeout=2)
        elif self.p is not None:
            time.sleep(0.15)
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
            self.f.flush()
            self.f.seek(0)
            for line in self.f:
                c = [x.strip() for x in line.split(",")]
                if len(c) < 7:
                    continue
                try:
                    self.sm.append(float(c[1])); self.mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            os.unlink(self.f.name)
        if not self.sm:
            return None
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": float(max(self.mx)), "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


def dist_setup(n):
    if n <= 1:
        return 0, 1
    os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", n))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    import datetime
    dist.init_process_group("nccl", timeout=datetime.timedelta(seconds=180))
    return rank, world


def timed_steps(fn, steps, warmup, world, sampler=None):
    """W untimed + exactly K timed steps, barrier + synchronize on both sides, device time, max over ranks.
    With a clock sampler, extra untimed warm-up steps keep the GPU under load until nvidia-smi delivers its first sample."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if sampler is not None and world == 1:
        t_end = time.perf_counter() + 3.0
        while not sampler.has_samples() and time.perf_counter() < t_end:
            fn()
            torch.cuda.synchronize()
    elif world > 1:
        # every rank must issue the same number of collectives: a FIXED number of extra untimed steps keeps the GPUs under
        # load while rank 0's nvidia-smi sampler starts (a rank-dependent loop here deadlocks the all-gather)
        for _ in range(100):
            fn()
        torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    return ms


# ----------------------------------------------------------------------------------------------------------------
def vector_levels(rows, rank, world):
    from seekstorm_b200.parallel import level_range
    n_levels = (rows + 65535) // 65536
    return n_levels, level_range(n_levels, rank, world)


def gen_vector_level(level, rows, dims, device):
    from seekstorm_b200 import synth
    n = min(65536, rows - level * 65536)
    return synth.gen_vectors(n, dims, 1002 * 1000 + level, device)


KERNELS = {"ffma": (1, 16, "scan_ffma", "scan_ffma (TMA + packed FP32 FFMA2 + warp top-k)"),
           "tc": (2, 128, "scan_tc", "scan_tc (TMA + tcgen05 3xTF32 split, TMEM accumulators, TMEM-epilogue top-k)"),
           "tc64": (3, 64, "scan_tc", "scan_tc<64> (tcgen05 3xTF32, 64 queries per pass)"),
           "tcb": (4, 128, "scan_tc", "scan_tc (TMA + tcgen05 3xBF16 split, TMEM accumulators, TMEM-epilogue top-k)"),
           "tcb64": (5, 64, "scan_tc", "scan_tc<64> (tcgen05 3xBF16, 64 queries per pass)")}
# DRAM traffic per corpus pass (dram__bytes_read.sum + dram__bytes_write.sum of ONE ncu --set full capture, divided by
# the passes in that launch, 1M x 768 corpus) from the committed captures under profiles/: traffic ~= algorithmic bytes
# (3.072 GB), i.e. no re-reads.
NCU = {"scan_ffma": {"traffic_per_pass": 24.608e9 / 8, "source": "profiles/r01_scan_ffma_v3.summary.txt"},
       "scan_tc": {"traffic_per_pass": 3.1149e9, "source": "profiles/r01_scan_tc_bf16.summary.txt"},
       # int8 full scan of 1M x 768: dram read 777.6 MB + write 29.3 MB (per-warp list scratch)
       "scan_tc_i8": {"traffic_per_pass": 0.8069e9, "source": "profiles/r01_scan_tc_i8_v1.summary.txt"}}


def measure_vector_kernel(a, ix, sh, kname, q_host, q_dev, keys, local_rows, rank, world, dev, want_clocks):
    kid, qt, kshort, klong = KERNELS[kname]
    ix.set_vector_kernel(kid)
    # ---- value: device-resident hot path (per rank scan + (N>1) NCCL all-gather of the packed keys) ----
    if world == 1:
        def step_dev():
            ix.search_vector_keys(q_dev, TOPK, keys)
    else:
        def step_dev():
            ix.search_vector_keys(q_dev, TOPK, keys)
            sh.gather_keys(keys)
    step_dev(); torch.cuda.synchronize()
    sampler = ClockSampler(dev.index) if want_clocks else None
    ms = timed_steps(step_dev, a.steps, a.warmup, world, sampler)
    clocks = sampler.stop() if sampler else None
    kern_ns = []
    for _ in range(5):       # duration of the dominant kernel: CUDA events the library records around that launch
        step_dev(); torch.cuda.synchronize()
        kern_ns.append(ix.last_stats()["dominant_kernel_ns"])
    launches = ix.last_stats()["kernel_launches"] + (1 if world > 1 else 0)
    passes = (a.batch + qt - 1) // qt
    # ---- e2e: the reference-facing call with HOST buffers (H2D queries, D2H hits inside the timed region) ----
    q_np = q_host.numpy()
    hits_buf, nh_buf = ix.hits_buffer(a.batch * TOPK), np.zeros(a.batch, dtype=np.uint32)
    if world == 1:
        def step_e2e():
            ix.search_vector_raw(q_np, TOPK, hits_buf, nh_buf)     # ssb_search_vector: host queries in, host hits out
    else:
        import torch.distributed as dist

        def step_e2e():
            qd = q_host.to(dev, non_blocking=True) if rank == 0 else q_dev
            dist.broadcast(qd, 0)
            sh.search_vector(qd, TOPK, raw_out=(hits_buf, nh_buf))
    ms_e2e = timed_steps(step_e2e, a.steps, a.warmup, world)
    peak, peak_kind = peaks()
    kern_ms = float(np.median(kern_ns)) / 1e6 if kern_ns and min(kern_ns) > 0 else None
    alg_bytes = float(local_rows) * a.dims * 4 * passes          # per launch (one launch = all passes of the batch)
    achieved = alg_bytes / (kern_ms / 1e3) / 1e9 if kern_ms else None
    ncu = NCU.get(kshort, {})
    return {
        "value": a.batch * a.steps / (ms / 1e3), "unit": "queries/s", "ms_per_step": ms / a.steps,
        "e2e": {"value": a.batch * a.steps / (ms_e2e / 1e3), "unit": "queries/s", "ms_per_step": ms_e2e / a.steps,
                "h2d_bytes_per_step": a.batch * a.dims * 4, "d2h_bytes_per_step": a.batch * 32 * 8},
        "gpu_launches": int(launches) * a.steps, "queries_per_pass": qt, "passes_per_step": passes, "kernel_desc": klong,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None,
                     "traffic": (ncu["traffic_per_pass"] * passes * local_rows / 1e6) if (ncu and a.dims == C2_DIMS) else None,
                     "traffic_source": ncu.get("source"), "peak_kind": f"of {peak_kind}", "kernel": kshort, "kernel_ms": kern_ms,
                     "algorithmic_bytes_per_launch": alg_bytes},
        "clocks": clocks,
    }


def bench_vector(a, rank, world, out):
    from seekstorm_b200 import Index, VectorSimilarity, synth
    from seekstorm_b200.parallel import ShardedSearcher
    dev = torch.device("cuda", torch.cuda.current_device())
    ix = Index(dev.index, vector_dims=a.dims, vector_similarity=VectorSimilarity.Cosine, max_batch=max(a.batch, 16))
    ix.set_stream(torch.cuda.current_stream().cuda_stream)
    n_levels, mine = vector_levels(a.rows, rank, world)
    local_rows = 0
    for lv in mine:
        r = gen_vector_level(lv, a.rows, a.dims, dev)
        ix.add_vector_level(lv, r)
        local_rows += r.shape[0]
        del r
    q_host = synth.gen_vectors(a.batch, a.dims, 2002, "cpu").pin_memory()
    q_dev = q_host.to(dev)
    keys = torch.zeros((a.batch, 32), dtype=torch.int64, device=dev)
    sh = ShardedSearcher(ix)
    names = ["ffma", "tcb"] if a.vector_kernel == "both" else [a.vector_kernel]
    res = {k: measure_vector_kernel(a, ix, sh, k, q_host, q_dev, keys, local_rows, rank, world, dev, rank == 0) for k in names}
    # batch-size sweep through the reference-facing call (host buffers, AUTO kernel choice): latency at batch 1 .. 256
    sweep = {}
    if world == 1:
        ix.set_vector_kernel(0)
        for bs in (1, 8, 64, 256):
            if bs > a.batch:
                continue
            qn = q_host.numpy()[:bs].copy()
            hb, nb = ix.hits_buffer(bs * TOPK), np.zeros(bs, dtype=np.uint32)

            def step_b():
                ix.search_vector_raw(qn, TOPK, hb, nb)
            msb = timed_steps(step_b, max(5, a.steps // 2), 2, world)
            per = msb / max(5, a.steps // 2)
            sweep[str(bs)] = {"ms_per_call": per, "queries_per_s": bs / (per / 1e3)}
    best = max(names, key=lambda k: res[k]["value"])      # headline = what SSB_VEC_KERNEL_AUTO picks for this batch size
    r = res[best]
    out.update({
        "metric": "queries/sec at top-10 (1M x 768 f32 cosine brute-force kNN)", "value": r["value"], "unit": "queries/s",
        "ms_per_step": r["ms_per_step"], "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C2 brute-force cosine kNN: {a.rows} x {a.dims} f32, top-{TOPK}, batch {a.batch} queries/step "
                               f"({r['passes_per_step']} corpus passes of {r['queries_per_pass']} queries)",
                   "l2": "inputs larger than L2 (corpus %.2f GB per GPU)" % (local_rows * a.dims * 4 / 1e9),
                   "parallelism": f"64K-row levels sharded over {world} GPU(s)", "kernel": r["kernel_desc"]},
        "e2e": r["e2e"], "gpu_launches": r["gpu_launches"], "roofline": r["roofline"], "clocks": r["clocks"],
        "batch_sweep_e2e": sweep,
        "kernels": {{"ffma": "scan_ffma", "tc": "scan_tc_tf32", "tc64": "scan_tc_tf32_n64", "tcb": "scan_tc_bf16", "tcb64": "scan_tc_bf16_n64"}[k]:
                    {kk: vv for kk, vv in res[k].items() if kk != "kernel_desc"} for k in names},
    })
    return ix, q_host


def bench_vector_int8(a, rank, world):
    """C2 corpus with Cosine + ScalarQuantizationI8: int8 corpus, tcgen05 kind::i8 scan, exact scores."""
    from seekstorm_b200 import Index, VectorSimilarity, synth
    from seekstorm_b200.parallel import ShardedSearcher
    dev = torch.device("cuda", torch.cuda.current_device())
    nb = a.int8_batch
    ix = Index(dev.index, vector_dims=a.dims, vector_similarity=VectorSimilarity.Cosine, max_batch=max(nb, 16), vector_quantization=1)
    ix.set_stream(torch.cuda.current_stream().cuda_stream)
    n_levels, mine = vector_levels(a.rows, rank, world)
    local_rows = 0
    for lv in mine:
        r = gen_vector_level(lv, a.rows, a.dims, dev)
        ix.add_vector_level(lv, r)
        local_rows += r.shape[0]
        del r
    q_host = synth.gen_vectors(nb, a.dims, 2002, "cpu").pin_memory()
    q_dev = q_host.to(dev)
    keys = torch.zeros((nb, 32), dtype=torch.int64, device=dev)
    sh = ShardedSearcher(ix)

    def step_dev():
        ix.search_vector_keys(q_dev, TOPK, keys)
        if world > 1:
            sh.gather_keys(keys)
    step_dev(); torch.cuda.synchronize()
    ms = timed_steps(step_dev, a.steps, a.warmup, world)
    kern_ns = []
    for _ in range(5):
        step_dev(); torch.cuda.synchronize()
        kern_ns.append(ix.last_stats()["dominant_kernel_ns"])
    launches = ix.last_stats()["kernel_launches"] + (1 if world > 1 else 0)
    q_np = q_host.numpy()
    hits_buf, nh_buf = ix.hits_buffer(nb * TOPK), np.zeros(nb, dtype=np.uint32)
    if world == 1:
        def step_e2e():
            ix.search_vector_raw(q_np, TOPK, hits_buf, nh_buf)
    else:
        import torch.distributed as dist

        def step_e2e():
            qd = q_host.to(dev, non_blocking=True) if rank == 0 else q_dev
            dist.broadcast(qd, 0)
            sh.search_vector(qd, TOPK, raw_out=(hits_buf, nh_buf))
    ms_e2e = timed_steps(step_e2e, a.steps, a.warmup, world)
    sweep = {}
    if world == 1:
        for bs in (1, 128, 256):
            qn = q_np[:bs].copy()
            hb, nbuf = ix.hits_buffer(bs * TOPK), np.zeros(bs, dtype=np.uint32)

            def step_b():
                ix.search_vector_raw(qn, TOPK, hb, nbuf)
            n_it = max(5, a.steps // 2)
            per = timed_steps(step_b, n_it, 2, world) / n_it
            sweep[str(bs)] = {"ms_per_call": per, "queries_per_s": bs / (per / 1e3)}
    peak, peak_kind = peaks()
    passes = (nb + 127) // 128
    kern_ms = float(np.median(kern_ns)) / 1e6 if kern_ns and min(kern_ns) > 0 else None
    alg_bytes = float(local_rows) * a.dims * 1 * passes
    achieved = alg_bytes / (kern_ms / 1e3) / 1e9 if kern_ms else None
    ix.close()
    return {
        "metric": "queries/sec at top-10 (1M x 768 cosine, ScalarQuantizationI8 brute-force kNN)",
        "value": nb * a.steps / (ms / 1e3), "unit": "queries/s", "ms_per_step": ms / a.steps, "dtype": "i8 (int32 accumulate, exact)",
        "config": {"workload": f"C2 corpus quantised to int8 (Cosine + ScalarQuantizationI8): {a.rows} x {a.dims}, top-{TOPK}, "
                               f"batch {nb} queries/step ({passes} corpus passes of 128 queries)",
                   "kernel": "scan_tc<128, i8> (tcgen05 kind::i8, TMEM s32 accumulators)"},
        "e2e": {"value": nb * a.steps / (ms_e2e / 1e3), "unit": "queries/s", "ms_per_step": ms_e2e / a.steps,
                "h2d_bytes_per_step": nb * a.dims * 4, "d2h_bytes_per_step": nb * 32 * 8},
        "gpu_launches": int(launches) * a.steps, "batch_sweep_e2e": sweep,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                     "traffic": (NCU["scan_tc_i8"]["traffic_per_pass"] * passes * local_rows / 1e6) if a.dims == C2_DIMS else None,
                     "traffic_source": NCU["scan_tc_i8"]["source"],
                     "peak_kind": f"of {peak_kind}", "kernel": "scan_tc_i8", "kernel_ms": kern_ms,
                     "algorithmic_bytes_per_launch": alg_bytes},
    }


def cpu_vector_int8_baseline(a, seconds):
    """Restated reference CPU path for Cosine + SQ-I8 (dot_i8 over the int8 corpus, linear top-k), one query per thread."""
    from oracle import oracle as O
    from seekstorm_b200 import synth
    cores = os.cpu_count() or 1
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    rows = np.empty((a.rows, a.dims), dtype=np.int8)
    sl = 65536

    def quant(lv):
        r = gen_vector_level(lv, a.rows, a.dims, dev).cpu().numpy()
        rows[lv * sl: lv * sl + r.shape[0]] = O.quantize_rows_i8(r)
    for lv in range((a.rows + sl - 1) // sl):
        quant(lv)
    qs = synth.gen_vectors(64, a.dims, 2002, "cpu").numpy()
    q8 = O.quantize_rows_i8(qs)
    O.search_vector_i8(rows, q8[0], TOPK)
    done = [0] * cores
    stop = time.perf_counter() + seconds
    nxt = [0]
    lock = threading.Lock()

    def work(i):
        while time.perf_counter() < stop:
            with lock:
                j = nxt[0]; nxt[0] += 1
            O.search_vector_i8(rows, q8[j % len(q8)], TOPK)
            done[i] += 1
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
    n_done = sum(done)
    return {"value": n_done / dt, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"{n_done} queries, full {a.rows}x{a.dims} int8 corpus, {cores} threads (one query each), {dt:.1f}s"}


def cpu_vector_baseline(a, seconds):
    """The restated reference CPU path (oracle: exhaustive scan, 8-lane FMA dot as dot_f32_avx2, linear top-k), all
    host threads, on a bounded sample of the C2 queries."""
    from oracle import oracle as O
    from seekstorm_b200 import synth
    cores = os.cpu_count() or 1
    rows = np.empty((a.rows, a.dims), dtype=np.float32)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    for lv in range((a.rows + 65535) // 65536):
        r = gen_vector_level(lv, a.rows, a.dims, dev)
        r = r / r.norm(dim=1, keepdim=True)
        rows[lv * 65536: lv * 65536 + r.shape[0]] = r.cpu().numpy()
    # spread the corpus pages over the NUMA nodes: re-copy it with one first-touching worker per slice (the single
    # allocating thread above would otherwise place all 3 GB on its own node and cap the scan at one socket's bandwidth)
    rows2 = np.empty_like(rows)
    sl = max(1, (a.rows + cores - 1) // cores)

    def touch(i):
        rows2[i * sl:(i + 1) * sl] = rows[i * sl:(i + 1) * sl]
    tt = [threading.Thread(target=touch, args=(i,)) for i in range(cores)]
    [t.start() for t in tt]; [t.join() for t in tt]
    rows = rows2
    qs = synth.gen_vectors(a.batch, a.dims, 2002, "cpu").numpy()
    qn = [O.normalize(q) for q in qs]
    O.search_vector(rows, qn[0], TOPK, O.SIM_COSINE, lanes8=True, n_threads=cores)  # warm (page in the corpus)
    # one worker per core, each answering whole queries single-threaded (8-lane FMA dot, linear top-k): the
    # throughput-optimal arrangement of the reference's per-shard scan on this host
    done = [0] * cores
    stop = time.perf_counter() + seconds
    nxt = [0]
    lock = threading.Lock()

    def work(i):
        while time.perf_counter() < stop:
            with lock:
                j = nxt[0]; nxt[0] += 1
            O.search_vector(rows, qn[j % len(qn)], TOPK, O.SIM_COSINE, lanes8=True, n_threads=1)
            done[i] += 1
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
    n_done = sum(done)
    return {"value": n_done / dt, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"{n_done} queries (cycling the {a.batch} C2 queries), full {a.rows}x{a.dims} corpus, {cores} threads (one query each), {dt:.1f}s"}


# ----------------------------------------------------------------------------------------------------------------
def build_bm25(a, rank, world, dev, want_host_copy):
    from seekstorm_b200 import Index, synth
    from seekstorm_b200.parallel import level_range, allreduce_global_df
    ix = Index(dev.index if dev.type == "cuda" else 0, max_batch=a.bm25_batch)
    ix.set_stream(torch.cuda.current_stream().cuda_stream)
    n_levels = (a.bm25_docs + 65535) // 65536
    mine = level_range(n_levels, rank, world)
    len_sum = torch.zeros(1, dtype=torch.int64, device=dev)
    host_levels = []
    for lv in synth.gen_lexical_corpus(a.bm25_docs, C3_VOCAB, 1003, dev, level_ids=mine):
        ix.add_synth_level(lv)
        len_sum += lv.len_sum_normalized
        if want_host_copy:
            host_levels.append(lv.to_numpy())
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(len_sum)
    ix.commit(a.bm25_docs, int(len_sum.item()))
    if world > 1:
        allreduce_global_df(ix)
    return ix, host_levels, int(len_sum.item())


def bm25_queries(n):
    from seekstorm_b200 import synth
    qs = synth.gen_queries(n, 2003, 20, 100000, (2, 3, 4), (0.4, 0.4, 0.2))
    return [[int(k) for k in synth.term_keys_np(np.array(q, dtype=np.int64))] for q in qs]


def bench_bm25(a, rank, world):
    from seekstorm_b200 import QueryType, ResultType
    from seekstorm_b200.parallel import ShardedSearcher
    dev = torch.device("cuda", torch.cuda.current_device())
    t0 = time.perf_counter()
    ix, _, _ = build_bm25(a, rank, world, dev, False)
    build_s = time.perf_counter() - t0
    qk = bm25_queries(a.bm25_batch)
    b, keep = ix._lex_batch(qk, QueryType.Union)
    offs_dev = torch.from_numpy(keep[0].view(np.int32)).to(dev)
    keys_dev = torch.from_numpy(keep[1].view(np.int64)).to(dev)
    from seekstorm_b200._lib import SsbLexBatch
    b_dev = SsbLexBatch(len(qk), int(QueryType.Union), offs_dev.data_ptr(), keys_dev.data_ptr())
    out_keys = torch.zeros((len(qk), 32), dtype=torch.int64, device=dev)
    sh = ShardedSearcher(ix)

    def step_dev():
        ix.search_lexical_keys(b_dev, TOPK, ResultType.Topk, out_keys)
        if world > 1:
            sh.gather_keys(out_keys)
    steps = max(3, a.steps // 2)
    ms = timed_steps(step_dev, steps, a.warmup, world)
    kern_ns = []
    for _ in range(3):
        step_dev(); torch.cuda.synchronize()
        kern_ns.append(ix.last_stats()["dominant_kernel_ns"])

    hits_buf, nh_buf, cnt_buf = ix.hits_buffer(len(qk) * TOPK), np.zeros(len(qk), dtype=np.uint32), np.zeros(len(qk), dtype=np.uint64)

    def step_e2e():
        if world == 1:
            ix.search_lexical_raw(b, TOPK, ResultType.Topk, hits_buf, nh_buf, cnt_buf)   # ssb_search_lexical, host buffers
        else:
            sh.search_lexical(b, len(qk), TOPK, ResultType.Topk, dev, raw_out=(hits_buf, nh_buf))
    ms_e2e = timed_steps(step_e2e, steps, a.warmup, world)
    st = ix.last_stats() if world == 1 else {}
    # secondary modes on the same index / queries (device-resident, same timing rules): exact counts and AND
    variants = {}
    if world == 1:
        for name, qt_, rt_ in (("or_topkcount", QueryType.Union, ResultType.TopkCount), ("and_topkcount", QueryType.Intersection, ResultType.TopkCount),
                               ("and_topk", QueryType.Intersection, ResultType.Topk)):
            bv = SsbLexBatch(len(qk), int(qt_), offs_dev.data_ptr(), keys_dev.data_ptr())
            cnt_dev = torch.zeros(len(qk), dtype=torch.int64, device=dev)

            def step_v():
                ix.search_lexical_keys(bv, TOPK, rt_, out_keys, cnt_dev)
            msv = timed_steps(step_v, max(2, steps // 2), 2, world)
            variants[name] = {"value": len(qk) * max(2, steps // 2) / (msv / 1e3), "unit": "queries/s"}
    peak, peak_kind = peaks()
    kern_ms = float(np.median(kern_ns)) / 1e6 if kern_ns and min(kern_ns) > 0 else None
    alg = st.get("algorithmic_bytes")
    res = {
        "metric": "queries/sec at top-10 (BM25 OR, block-max pruned, ResultType::Topk)", "value": len(qk) * steps / (ms / 1e3),
        "unit": "queries/s", "ms_per_step": ms / steps, "steps": steps, "dtype": "f32 scores / u16 postings",
        "config": {"workload": f"C3 BM25 OR top-{TOPK}: {a.bm25_docs} docs Zipf(1) V={C3_VOCAB}, {len(qk)} queries/step of 2-4 terms (40/40/20%), ranks log-uniform [20,1e5]",
                   "index_build_s": build_s},
        "e2e": {"value": len(qk) * steps / (ms_e2e / 1e3), "unit": "queries/s", "ms_per_step": ms_e2e / steps,
                "h2d_bytes_per_step": int(keep[0].nbytes + keep[1].nbytes), "d2h_bytes_per_step": len(qk) * (32 * 8 + 8)},
        "gpu_launches": 3 * steps, "variants": variants,
        "roofline": {"bound": "hbm", "achieved": (alg / (kern_ms / 1e3) / 1e9) if (alg and kern_ms) else None, "peak": peak, "unit": "GB/s",
                     "frac": (alg / (kern_ms / 1e3) / 1e9 / peak) if (alg and kern_ms) else None,
                     # dram__bytes_read+write of one ncu --set full capture of this launch shape (profiles/r01_lex_score_v3: 8.39 GB
                     # read + 1.74 GB written): several times the algorithmic bytes — 4- and 8-byte probes cost 32-byte sectors,
                     # and the write side is per-thread stack traffic of the register-array paths (open item)
                     "traffic": 10.13e9 if (a.bm25_docs == C3_DOCS and len(qk) == 4096 and world == 1) else None,
                     "traffic_source": "profiles/r01_lex_score_v3.summary.txt",
                     "peak_kind": f"of {peak_kind}", "kernel": "lex_score", "kernel_ms": kern_ms,
                     "algorithmic_bytes_per_launch": alg, "postings_visited": st.get("postings_visited"), "probes": st.get("probes"),
                     "items_processed": st.get("items_processed"), "items_skipped": st.get("items_skipped")},
    }
    ix.close()
    return res


def bench_hybrid(a, rank, world):
    """C4: SearchMode::Hybrid (BM25 OR top-10 + 768-d cosine top-10, RRF k=0.6) over 5M docs, 1 GPU, through
    ssb_search_hybrid with host buffers (the RRF join runs on the host inside the library, search.rs:1962-2035)."""
    from seekstorm_b200 import Index, QueryType, VectorSimilarity, synth
    from seekstorm_b200._lib import check, lib
    import ctypes as C
    dev = torch.device("cuda", torch.cuda.current_device())
    n_docs = a.hybrid_docs
    ix = Index(dev.index, vector_dims=C2_DIMS, vector_similarity=VectorSimilarity.Cosine, max_batch=1024)
    ix.set_stream(torch.cuda.current_stream().cuda_stream)
    len_sum = 0
    for lv in synth.gen_lexical_corpus(n_docs, C3_VOCAB, 1004, dev):
        ix.add_synth_level(lv)
        len_sum += lv.len_sum_normalized
    ix.commit(n_docs, len_sum)
    for lv in range((n_docs + 65535) // 65536):
        ix.add_vector_level(lv, synth.gen_vectors(min(65536, n_docs - lv * 65536), C2_DIMS, 1005 * 1000 + lv, dev))
    nq = 1000
    qs = synth.gen_queries(nq, 2004, 20, 100000, (2, 3, 4), (0.4, 0.4, 0.2))
    qk = [[int(k) for k in synth.term_keys_np(np.array(q, dtype=np.int64))] for q in qs]
    b, keep = ix.make_lex_batch(qk, QueryType.Union)
    qv = synth.gen_vectors(nq, C2_DIMS, 2005, "cpu").numpy()
    hits, nh = ix.hits_buffer(nq * TOPK), np.zeros(nq, dtype=np.uint32)

    def step():
        check(lib().ssb_search_hybrid(ix._h, C.byref(b), qv.ctypes.data, TOPK, hits.ctypes.data, nh.ctypes.data))
    steps = max(3, a.steps // 4)
    ms = timed_steps(step, steps, 2, world)
    ix.close()
    return {"metric": "queries/sec at top-10 (hybrid: BM25 OR + 768-d cosine, RRF)", "value": nq * steps / (ms / 1e3), "unit": "queries/s",
            "ms_per_step": ms / steps, "steps": steps,
            "config": {"workload": f"C4 hybrid: {n_docs} docs (Zipf lexical index + {n_docs} x {C2_DIMS} f32 vectors), {nq} queries/step, e2e through ssb_search_hybrid (host buffers)"},
            "h2d_bytes_per_step": int(qv.nbytes + keep[0].nbytes + keep[1].nbytes), "d2h_bytes_per_step": nq * 32 * 16}


def cpu_bm25_baseline(a, seconds):
    """Reference-shaped CPU search (oracle pruned path: block-max ordered AND + MAXSCORE sub-queries) on the same index,
    one worker thread per host core (the reference runs one task per shard, default shards = cores)."""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    from seekstorm_b200 import synth
    orc = O.OracleIndex()
    len_sum = 0
    for lv in synth.gen_lexical_corpus(a.bm25_docs, C3_VOCAB, 1003, dev):
        orc.add_level(lv.to_numpy())
        len_sum += lv.len_sum_normalized
    orc.commit(a.bm25_docs, len_sum)
    qk = bm25_queries(a.bm25_batch)
    done = [0] * cores
    stop = time.perf_counter() + seconds
    nxt = [0]
    lock = threading.Lock()

    def work(i):
        while time.perf_counter() < stop:
            with lock:
                j = nxt[0]; nxt[0] += 1
            orc.search(qk[j % len(qk)], O.QUERY_UNION, TOPK, O.RESULT_TOPK, pruned=True)
            done[i] += 1
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
    n = sum(done)
    return {"value": n / dt, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"{n} queries (cycling the {len(qk)} C3 queries) on the full {a.bm25_docs}-doc index, {cores} threads (one query each), {dt:.1f}s"}


# ----------------------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def emit(obj):
    """The ONE JSON line of this run, written straight to the process's original stdout."""
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    global _REAL_STDOUT
    a = parse()
    # stdout carries exactly one JSON line: everything else any library writes to fd 1 (NCCL prints its version banner
    # there whenever NCCL_DEBUG >= VERSION) is sent to stderr instead
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    sections = [s for s in a.sections.split(",") if s]
    if a.impl == "reference":
        rank = int(os.environ.get("RANK", 0))
        if rank != 0:
            return 0
        import __graft_entry__ as g
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
        base = cpu_vector_baseline(a, max(a.cpu_seconds, 2.0) * max(1, min(a.steps, 3)))
        line = {"impl": "reference", "metric": "queries/sec at top-10 (1M x 768 f32 cosine brute-force kNN)", "value": base["value"],
                "unit": "queries/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": None,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"C2 brute-force cosine kNN: {a.rows} x {a.dims} f32, top-{TOPK} (restated reference CPU path, {base['cores']} threads)"},
                "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        if "bm25" in sections:
            try:
                line["bm25"] = {"cpu_baseline": cpu_bm25_baseline(a, a.cpu_seconds)}
                line["bm25"]["value"] = line["bm25"]["cpu_baseline"]["value"]
            except Exception as e:  # pragma: no cover
                line["bm25"] = {"error": repr(e)}
        emit(line)
        return 0

    if not torch.cuda.is_available():
        emit({"error": "no CUDA device: bench.py measures the B200 path only (no CPU fallback)"})
        return 1
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()
    rank, world = dist_setup(a.gpus)
    # one explicit (non-default) stream for everything: library kernels, torch CUDA events and NCCL collectives
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    out = {"n_gpus": world, "steps": a.steps, "warmup": a.warmup, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "impl": "b200"}
    ix, _ = bench_vector(a, rank, world, out)
    ix.close()
    del ix
    torch.cuda.empty_cache()
    if "int8" in sections:
        try:
            out["int8"] = bench_vector_int8(a, rank, world)
        except Exception as e:  # pragma: no cover
            out["int8"] = {"error": repr(e)}
        torch.cuda.empty_cache()
    if "bm25" in sections:
        try:
            out["bm25"] = bench_bm25(a, rank, world)
        except Exception as e:  # pragma: no cover
            out["bm25"] = {"error": repr(e)}
    if "hybrid" in sections and world == 1:
        torch.cuda.empty_cache()
        try:
            out["hybrid"] = bench_hybrid(a, rank, world)
        except Exception as e:  # pragma: no cover
            out["hybrid"] = {"error": repr(e)}
    if rank == 0 and world == 1 and a.cpu_seconds > 0:
        torch.cuda.empty_cache()
        try:
            out["cpu_baseline"] = cpu_vector_baseline(a, a.cpu_seconds)
        except Exception as e:  # pragma: no cover
            out["cpu_baseline"] = {"error": repr(e)}
        if "int8" in sections and isinstance(out.get("int8"), dict) and "error" not in out["int8"]:
            try:
                out["int8"]["cpu_baseline"] = cpu_vector_int8_baseline(a, min(a.cpu_seconds, 8.0))
            except Exception as e:  # pragma: no cover
                out["int8"]["cpu_baseline"] = {"error": repr(e)}
        if "bm25" in sections and isinstance(out.get("bm25"), dict) and "error" not in out["bm25"]:
            try:
                out["bm25"]["cpu_baseline"] = cpu_bm25_baseline(a, a.cpu_seconds)
            except Exception as e:  # pragma: no cover
                out["bm25"]["cpu_baseline"] = {"error": repr(e)}
    if rank == 0:
        emit(out)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
