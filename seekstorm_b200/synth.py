"""Seeded synthetic corpora / query sets for the BASELINE.json configs.

Host-side plumbing only (torch tensor ops, on CPU or CUDA).  Produces the *neutral level layout*
consumed by both the C-ABI (`ssb_level_desc`, include/seekstorm_b200.h) and the CPU oracle
(same struct on its side): per 64K-doc level, per term, ascending u16 local doc ids + u16 tf,
plus the byte4 doc-length codes (reference: index.rs:4237-4279, 5397-5405).

Law: vocabulary V, doc length L ~ round(lognormal(ln 80, 0.6)) clamped to [8, 2000],
tokens i.i.d. Zipf(s=1) over V, one indexed field.  Term ids are Zipf ranks (0-based); the 64-bit term
key is splitmix64(term_id) with the low 3 bits cleared (the reference reserves them for the n-gram type,
index.rs:4165-4225; the real key is an ahash of the term string, opaque to this path).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

LEVEL_DOCS = 65536
_M64 = (1 << 64) - 1


def int_to_byte4(i: int) -> int:
    """index.rs:4237-4251 (Lucene SmallFloat.intToByte4)."""
    if i < 24:
        return i
    ii = i - 24
    num_bits = ii.bit_length()
    if num_bits < 4:
        return 24 + ii
    shift = num_bits - 4
    return 24 + (((ii >> shift) & 0x07) | ((shift + 1) << 3))


def byte4_to_int(b: int) -> int:
    """index.rs:4255-4268."""
    if b < 24:
        return b
    i = b - 24
    bits, shift = i & 7, i >> 3
    if shift == 0:
        return 24 + bits
    return 24 + ((bits | 8) << (shift - 1))


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def term_keys_np(term_ids: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64(term_id) & ~7 -> uint64."""
    x = term_ids.astype(np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z & ~np.uint64(7)


def term_keys_torch(term_ids: torch.Tensor) -> torch.Tensor:
    """Same as term_keys_np but on any device; returns int64 holding the u64 bit pattern."""
    def mul(a, c):  # 64-bit wrapping multiply (int64 multiplication wraps in torch)
        return a * c
    def shr(a, s):  # logical shift right on int64
        return (a >> s) & ((1 << (64 - s)) - 1)
    def c64(v):  # python int (u64) -> signed int64 constant
        return v - (1 << 64) if v >= (1 << 63) else v
    x = term_ids.to(torch.int64) + c64(0x9E3779B97F4A7C15)
    z = x
    z = mul(z ^ shr(z, 30), c64(0xBF58476D1CE4E5B9))
    z = mul(z ^ shr(z, 27), c64(0x94D049BB133111EB))
    z = z ^ shr(z, 31)
    return z & ~7


@dataclass
class Level:
    """One 64K-doc level in the neutral layout (tensors live on the generating device)."""
    level_id: int
    n_docs: int
    term_ids: torch.Tensor         # int64 [n_terms] (synthetic only; keys derive from it)
    term_keys: torch.Tensor        # int64 (u64 bit pattern) [n_terms]
    posting_offsets: torch.Tensor  # int32 [n_terms+1]
    doc_ids: torch.Tensor          # int16 (u16 bit pattern) [n_postings]
    tfs: torch.Tensor              # int16 (u16 bit pattern) [n_postings]
    doc_len_bytes: torch.Tensor    # uint8 [n_docs]
    len_sum_normalized: int        # Σ byte4_to_int(doc_len_byte)
    positions: torch.Tensor | None = None   # int16 (u16 bit pattern) [sum of tfs]: token positions of every posting, posting order (phrase queries)

    def to_numpy(self) -> dict:
        extra = {} if self.positions is None else dict(positions=self.positions.cpu().numpy().view(np.uint16).copy())
        return dict(
            **extra,
            level_id=self.level_id, n_docs=self.n_docs,
            term_keys=self.term_keys.cpu().numpy().view(np.uint64).copy(),
            posting_offsets=self.posting_offsets.cpu().numpy().view(np.uint32).copy(),
            doc_ids=self.doc_ids.cpu().numpy().view(np.uint16).copy(),
            tfs=self.tfs.cpu().numpy().view(np.uint16).copy(),
            doc_len_bytes=self.doc_len_bytes.cpu().numpy().copy(),
        )


_B4_LUT = None
_DLC = None


def _luts(device):
    global _B4_LUT, _DLC
    if _B4_LUT is None:
        _B4_LUT = torch.tensor([int_to_byte4(i) for i in range(4096)], dtype=torch.uint8)
        _DLC = torch.tensor([byte4_to_int(b) for b in range(256)], dtype=torch.int64)
    return _B4_LUT.to(device), _DLC.to(device)


def zipf_cdf(vocab: int, device, s: float = 1.0) -> torch.Tensor:
    w = 1.0 / torch.arange(1, vocab + 1, dtype=torch.float64, device=device) ** s
    cdf = torch.cumsum(w, 0)
    return (cdf / cdf[-1]).contiguous()


def gen_level(level_id: int, n_docs: int, vocab: int, seed: int, device="cpu",
              cdf: torch.Tensor | None = None, mean_len: float = 80.0, sigma: float = 0.6,
              min_len: int = 8, max_len: int = 2000, with_positions: bool = False) -> Level:
    """Generate one level; deterministic in (seed, level_id) for a given device type.  with_positions: also return every posting's token
    positions (the documents ARE token sequences: position = index of the token inside its document); the postings are the same either way."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed * 1000003 + level_id)
    if cdf is None:
        cdf = zipf_cdf(vocab, dev)
    b4, dlc = _luts(dev)
    ln = torch.randn(n_docs, generator=g, device=dev, dtype=torch.float32) * sigma + math.log(mean_len)
    lens = torch.exp(ln).round().clamp_(min_len, max_len).to(torch.int64)
    doc_len_bytes = b4[lens]
    len_sum = int(dlc[doc_len_bytes.to(torch.int64)].sum().item())
    n_tok = int(lens.sum().item())
    u = torch.rand(n_tok, generator=g, device=dev, dtype=torch.float64)
    terms = torch.searchsorted(cdf, u).clamp_(max=vocab - 1)
    del u
    docs = torch.repeat_interleave(torch.arange(n_docs, device=dev, dtype=torch.int64), lens)
    key = terms * LEVEL_DOCS + docs
    positions = None
    if with_positions:
        starts = torch.cumsum(lens, 0) - lens
        pos = torch.arange(n_tok, device=dev, dtype=torch.int64) - starts[docs]
        key, order = torch.sort(key, stable=True)       # tokens are in (doc, position) order: a stable sort keeps positions ascending per posting
        positions = pos[order].to(torch.int16)
        del pos, order, starts
    else:
        key, _ = torch.sort(key)
    del terms, docs
    pk, tf = torch.unique_consecutive(key, return_counts=True)
    del key
    p_term = pk // LEVEL_DOCS
    p_doc = pk - p_term * LEVEL_DOCS
    term_ids, counts = torch.unique_consecutive(p_term, return_counts=True)
    offs = torch.zeros(term_ids.numel() + 1, dtype=torch.int64, device=dev)
    offs[1:] = torch.cumsum(counts, 0)
    return Level(
        level_id=level_id, n_docs=n_docs, term_ids=term_ids, term_keys=term_keys_torch(term_ids),
        posting_offsets=offs.to(torch.int32), doc_ids=p_doc.to(torch.int16),
        tfs=tf.clamp(max=65535).to(torch.int16), doc_len_bytes=doc_len_bytes, len_sum_normalized=len_sum, positions=positions)


def gen_lexical_corpus(n_docs: int, vocab: int, seed: int, device="cpu", level_ids=None, with_positions: bool = False):
    """Yield Level objects for the corpus; `level_ids` restricts to a subset (multi-GPU block ranges).
    Returns an iterator; global stats come from `corpus_stats`."""
    dev = torch.device(device)
    cdf = zipf_cdf(vocab, dev)
    n_levels = (n_docs + LEVEL_DOCS - 1) // LEVEL_DOCS
    for lv in (range(n_levels) if level_ids is None else level_ids):
        nd = min(LEVEL_DOCS, n_docs - lv * LEVEL_DOCS)
        yield gen_level(lv, nd, vocab, seed, dev, cdf, with_positions=with_positions)


def gen_queries(n_queries: int, seed: int, rank_lo: int, rank_hi: int, n_terms_choices=(2,),
                n_terms_probs=(1.0,)) -> list[list[int]]:
    """Distinct term ranks per query, log-uniform in [rank_lo, rank_hi]. Returns 0-based
    term ids (= rank-1)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_queries):
        nt = int(rng.choice(n_terms_choices, p=n_terms_probs))
        terms: list[int] = []
        while len(terms) < nt:
            r = int(math.floor(math.exp(rng.uniform(math.log(rank_lo), math.log(rank_hi + 1)))))
            r = min(max(r, rank_lo), rank_hi)
            if r - 1 not in terms:
                terms.append(r - 1)
        out.append(terms)
    return out


def gen_vectors(n: int, dims: int, seed: int, device="cpu", normalize: bool = False,
                chunk: int = 1 << 18) -> torch.Tensor:
    """Row-major f32 [n, dims], i.i.d. N(0,1) (optionally L2-normalised in f32)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = torch.empty((n, dims), dtype=torch.float32, device=dev)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        out[s:e] = torch.randn((e - s, dims), generator=g, device=dev, dtype=torch.float32)
    if normalize:
        out /= out.norm(dim=1, keepdim=True)
    return out
