"""Shared test helpers: tiny corpora -> neutral level arrays; oracle + GPU index builders."""
import numpy as np

from oracle import oracle as O
from seekstorm_b200 import synth


def key_of(term: str) -> int:
    from seekstorm_b200.index import synthetic_term_key
    return synthetic_term_key(term)


def level_from_postings(level_id, n_docs, postings: dict, len_bytes):
    """postings: term -> sorted [(local_doc, tf)...]"""
    terms = sorted(postings.keys())
    keys = np.array([key_of(t) for t in terms], dtype=np.uint64)
    offs = np.zeros(len(terms) + 1, dtype=np.uint32)
    ids, tfs = [], []
    for i, t in enumerate(terms):
        for d, tf in postings[t]:
            ids.append(d); tfs.append(tf)
        offs[i + 1] = len(ids)
    return dict(level_id=level_id, n_docs=n_docs, term_keys=keys, posting_offsets=offs,
                doc_ids=np.array(ids, dtype=np.uint16), tfs=np.array(tfs, dtype=np.uint16),
                doc_len_bytes=np.array(len_bytes, dtype=np.uint8))


def oracle_index(levels, n_docs, len_sum):
    ix = O.OracleIndex()
    for lv in levels:
        ix.add_level(lv)
    ix.commit(n_docs, len_sum)
    return ix


def gpu_index(levels, n_docs, len_sum, **kw):
    from seekstorm_b200 import Index
    ix = Index(0, **kw)
    for lv in levels:
        ix.add_lexical_level(lv["level_id"], lv["n_docs"], lv["term_keys"], lv["posting_offsets"], lv["doc_ids"],
                             lv["tfs"], lv["doc_len_bytes"])
    ix.commit(n_docs, len_sum)
    return ix


def synth_levels(n_docs, vocab, seed, device="cpu"):
    lvs = list(synth.gen_lexical_corpus(n_docs, vocab, seed, device))
    return lvs, sum(l.len_sum_normalized for l in lvs)


def query_keys(queries):
    return [[int(k) for k in synth.term_keys_np(np.array(q, dtype=np.int64))] for q in queries]
