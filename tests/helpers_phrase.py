"""Token-sequence corpora for the phrase tests: documents are real token sequences, so the ground truth of a phrase query is a
substring search over them (independent of the oracle's and the GPU's position arithmetic)."""
import numpy as np

from seekstorm_b200 import synth


def sequence_corpus(n_docs, vocab, seed, docs_per_level=65536, mean_len=30):
    """-> (docs: list of int arrays (token ids), levels: neutral level dicts incl. 'positions', len_sum)"""
    rng = np.random.default_rng(seed)
    w = 1.0 / (np.arange(vocab) + 3.0)
    w /= w.sum()
    lens = np.clip(rng.geometric(1.0 / mean_len, n_docs), 2, 400)
    docs = [rng.choice(vocab, size=int(n), p=w).astype(np.int64) for n in lens]
    levels, len_sum = [], 0
    for li, base in enumerate(range(0, n_docs, docs_per_level)):
        nd = min(docs_per_level, n_docs - base)
        doc_of = np.concatenate([np.full(len(docs[base + d]), d, dtype=np.int64) for d in range(nd)])
        tok = np.concatenate(docs[base: base + nd])
        pos = np.concatenate([np.arange(len(docs[base + d]), dtype=np.int64) for d in range(nd)])
        order = np.lexsort((pos, doc_of, tok))                         # term-major, doc ascending, position ascending
        tok, doc_of, pos = tok[order], doc_of[order], pos[order]
        # postings = runs of equal (term, doc)
        key = tok * nd + doc_of
        starts = np.flatnonzero(np.concatenate([[True], key[1:] != key[:-1]]))
        p_tok, p_doc = tok[starts], doc_of[starts]
        tfs = np.diff(np.concatenate([starts, [len(key)]]))
        t_starts = np.flatnonzero(np.concatenate([[True], p_tok[1:] != p_tok[:-1]]))
        terms = p_tok[t_starts]
        offs = np.concatenate([t_starts, [len(p_tok)]]).astype(np.uint32)
        lb = np.array([synth.int_to_byte4(len(docs[base + d])) for d in range(nd)], dtype=np.uint8)
        len_sum += int(sum(synth.byte4_to_int(int(b)) for b in lb))
        levels.append(dict(level_id=li, n_docs=nd, term_keys=synth.term_keys_np(terms).astype(np.uint64), posting_offsets=offs,
                           doc_ids=p_doc.astype(np.uint16), tfs=np.minimum(tfs, 65535).astype(np.uint16), doc_len_bytes=lb,
                           positions=pos.astype(np.uint16)))
    return docs, levels, len_sum


def phrase_queries(docs, seed, n, vocab):
    """phrases of 2..6 tokens: most cut out of real documents (they match at least there), some shuffled / random (mostly no match),
    some with a repeated token"""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        d = docs[int(rng.integers(0, len(docs)))]
        m = int(rng.integers(2, 7))
        if len(d) < m:
            continue
        s = int(rng.integers(0, len(d) - m + 1))
        ph = [int(x) for x in d[s: s + m]]
        r = rng.random()
        if r < 0.2:
            rng.shuffle(ph)
        elif r < 0.3:
            ph = [int(x) for x in rng.integers(0, min(vocab, 30), m)]
        elif r < 0.4:
            ph = ph[:2] + ph[:2] + ph[2:3]                      # a repeated bigram ("to be ... to be")
        out.append(ph)
    return out


def contains_phrase(doc, ph):
    m = len(ph)
    return any(all(doc[s + i] == ph[i] for i in range(m)) for s in range(len(doc) - m + 1))
