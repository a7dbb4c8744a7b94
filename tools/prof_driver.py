"""ncu driver: build the workload of one kernel family and run it a few times (the caller selects launches with ncu -k / -s / -c)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seekstorm_b200 import Index, QueryType, ResultType, VectorSimilarity, synth  # noqa: E402

what = sys.argv[1]
if what in ("tcb", "tcb256", "filt", "filt256", "filt256p", "ffma", "i8"):
    n, d = 1_000_000, 768
    ix = Index(0, vector_dims=d, vector_similarity=VectorSimilarity.Cosine, max_batch=1024, vector_quantization=1 if what == "i8" else 0)
    ix.reserve_vectors(n)
    for lv in range((n + 65535) // 65536):
        ix.add_vector_level(lv, synth.gen_vectors(min(65536, n - lv * 65536), d, 1002000 + lv, "cuda"))
    nq = {"tcb": 256, "tcb256": 256, "filt": 256, "filt256": 256, "filt256p": 256, "ffma": 16, "i8": 1024}[what]
    q = synth.gen_vectors(nq, d, 2002, "cuda")
    keys = torch.zeros((nq, 32), dtype=torch.int64, device="cuda")
    ix.set_vector_kernel({"tcb": 4, "tcb256": 6, "filt": 7, "filt256": 8, "filt256p": 9, "ffma": 1, "i8": 0}[what])
    for _ in range(3):
        ix.search_vector_keys(q, 10, keys); torch.cuda.synchronize()
else:
    n = 10_000_000
    ix = Index(0, max_batch=4096)
    ls = 0
    for lv in synth.gen_lexical_corpus(n, 1_000_000, 1003, "cuda"):
        ix.add_synth_level(lv); ls += lv.len_sum_normalized
    ix.commit(n, ls)
    qs = synth.gen_queries(4096, 2003, 20, 100000, (2, 3, 4), (0.4, 0.4, 0.2))
    qk = [[int(k) for k in synth.term_keys_np(np.array(q, dtype=np.int64))] for q in qs]
    qt = QueryType.Intersection if what == "lex_and" else QueryType.Union
    rt = ResultType.TopkCount if what == "lex_count" else ResultType.Topk
    b, keep = ix._lex_batch(qk, qt)
    keys = torch.zeros((4096, 32), dtype=torch.int64, device="cuda"); cnt = torch.zeros(4096, dtype=torch.int64, device="cuda")
    for _ in range(3):
        ix.search_lexical_keys(b, 10, rt, keys, cnt); torch.cuda.synchronize()
    print(ix.last_stats())
