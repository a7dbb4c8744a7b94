"""One int8 (ScalarQuantizationI8) scan of the C2 corpus for ncu captures: 1M x 768, 128 queries, top-10."""
import sys
import torch
sys.path.insert(0, ".")
from seekstorm_b200 import Index, VectorSimilarity, synth

n, d = 1_000_000, 768
ix = Index(0, vector_dims=d, vector_similarity=VectorSimilarity.Cosine, vector_quantization=1)
for lv in range((n + 65535) // 65536):
    m = min(65536, n - lv * 65536)
    ix.add_vector_level(lv, synth.gen_vectors(m, d, 1002 + lv, "cuda"))
q = synth.gen_vectors(128, d, 2002, "cuda")
keys = torch.zeros((128, 32), dtype=torch.int64, device="cuda")
for _ in range(3):
    ix.search_vector_keys(q, 10, keys)
torch.cuda.synchronize()
print(ix.last_stats())
