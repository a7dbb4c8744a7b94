"""Synthetic multi-field lexical levels (BM25F tests): per posting one tf per field (0 = term absent in that field, at least one > 0),
per field one doc-length byte array."""
import numpy as np


def multifield_levels(n_docs_total, vocab, n_fields, seed, docs_per_level=65536):
    """-> (levels as dicts: level_id, n_docs, term_keys u64, posting_offsets u32, doc_ids u16, tfs u16 [np, F], doc_len_bytes u8 [F, n_docs]), len_sum"""
    from seekstorm_b200 import synth
    rng = np.random.default_rng(seed)
    levels, len_sum = [], 0
    for li, base in enumerate(range(0, n_docs_total, docs_per_level)):
        n_docs = min(docs_per_level, n_docs_total - base)
        # Zipf-ish doc frequencies: term r occurs in ~ n_docs * min(0.5, 8 / (r + 8)) docs
        term_ids, offs, ids, tfs = [], [0], [], []
        for r in range(vocab):
            df = int(min(0.45, 6.0 / (r + 6.0)) * n_docs * rng.uniform(0.5, 1.0))
            if df == 0 and rng.uniform() < 0.7:
                continue
            df = max(df, 1)
            d = np.sort(rng.choice(n_docs, size=df, replace=False)).astype(np.uint16)
            t = np.zeros((df, n_fields), dtype=np.uint16)
            present = rng.uniform(size=(df, n_fields)) < (0.6 / (1 + np.arange(n_fields)))[None, :] + 0.15
            vals = rng.geometric(0.45, size=(df, n_fields)).astype(np.uint16)
            t[present] = vals[present]
            none = t.max(axis=1) == 0
            t[none, rng.integers(0, n_fields, size=int(none.sum()))] = 1
            term_ids.append(r); ids.append(d); tfs.append(t); offs.append(offs[-1] + df)
        lens = np.stack([np.clip(rng.lognormal(np.log(8.0 + 40.0 * f), 0.6, size=n_docs), 1, 2000).astype(np.int64) for f in range(n_fields)])
        lb = np.vectorize(lambda x: synth.int_to_byte4(int(x)))(lens).astype(np.uint8)
        len_sum += int(np.vectorize(lambda b: synth.byte4_to_int(int(b)))(lb).sum())
        keys = synth.term_keys_np(np.array(term_ids, dtype=np.int64)).astype(np.uint64)
        levels.append(dict(level_id=li, n_docs=n_docs, term_keys=keys, posting_offsets=np.array(offs, dtype=np.uint32),
                           doc_ids=np.concatenate(ids), tfs=np.ascontiguousarray(np.concatenate(tfs)), doc_len_bytes=np.ascontiguousarray(lb)))
    return levels, len_sum
