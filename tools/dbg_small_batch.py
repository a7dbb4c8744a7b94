import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seekstorm_b200 import Index, VectorSimilarity, synth
n, d = 1_000_000, 768
ix = Index(0, vector_dims=d, vector_similarity=VectorSimilarity.Cosine, max_batch=256)
ix.reserve_vectors(n)
for lv in range((n + 65535) // 65536):
    ix.add_vector_level(lv, synth.gen_vectors(min(65536, n - lv * 65536), d, 1002000 + lv, "cuda"))
q = synth.gen_vectors(256, d, 2002, "cpu").numpy()
for kern in (0, 7, 8, 4):
    ix.set_vector_kernel(kern)
    for bs in (1, 8, 64, 128, 129, 256):
        hb, nb = ix.hits_buffer(bs * 10), np.zeros(bs, dtype=np.uint32)
        qq = q[:bs].copy()
        for _ in range(3): ix.search_vector_raw(qq, 10, hb, nb)
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(10): ix.search_vector_raw(qq, 10, hb, nb)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t) * 100
        st = ix.last_stats()
        print(f"kern {kern} bs {bs:4d}: {ms:.3f} ms/call  fallbacks {st['filter_fallbacks']} launches {st['kernel_launches']} scan_ms {st['dominant_kernel_ns']/1e6:.3f} read {st['scan_bytes_read']/1e9:.2f} GB")
