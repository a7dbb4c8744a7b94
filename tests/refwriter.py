"""TEST INFRASTRUCTURE: a Python restatement of the reference's shard-file WRITER, used to manufacture index.bin / vector.bin
fixtures for the loader tests (no Rust toolchain exists here, so no file written by the reference itself can be produced:
loader parity is "unpinned" by the reference and pinned only against this restatement).

Restated from /root/reference/seekstorm/src/ (single indexed field, no n-gram keys, key_head_size 20):
  * level layout                     commit.rs:203-467 as read back by index.rs:3303-3516
  * key head                         compress_postinglist.rs:339-409
  * container choice                 compress_postinglist.rs:240-332 (RLE when runs < count/2 (count < 4096) or < 2048, else Array / Bitmap)
  * Array / Bitmap / RLE containers  compress_postinglist.rs:694-977
  * pointer width + pivot            index_posting.rs:193-200 (2-byte pointers while the key's position bytes < 32768)
  * embedded positions               index_posting.rs:472-500 (rule) and 590-640 (bit layout), single-field variants
  * position blobs                   VINT positions_count + VINT deltas, compress_postinglist.rs:946-977 (compress_positions)
  * vector levels                    vector.rs:1066-1094, VectorHeader vector.rs:62-73
"""
import struct

import numpy as np

SEGMENT_BITS = 11
KEY_HEAD_SIZE = 20


def vint(v: int) -> bytes:
    if v < 128:
        return bytes([v | 0x80])
    if v < 16384:
        return bytes([(v >> 7) & 0x7F, (v & 0x7F) | 0x80])
    return bytes([(v >> 14) & 0x7F, (v >> 7) & 0x7F, (v & 0x7F) | 0x80])


def _bits(x: int) -> int:
    return int(x).bit_length()


def _embed(deltas, ptr_size):
    """index_posting.rs:472-500: may the delta positions of this posting live inside its rank-position pointer?"""
    n = len(deltas)
    if ptr_size == 2:
        return (n == 1 and _bits(deltas[0]) <= 14) or (n == 2 and _bits(deltas[0]) <= 7 and _bits(deltas[1]) <= 7)
    return ((n == 1 and _bits(deltas[0]) <= 21) or (n == 2 and _bits(deltas[0]) <= 10 and _bits(deltas[1]) <= 11)
            or (n == 3 and all(_bits(d) <= 7 for d in deltas))
            or (n == 4 and all(_bits(d) <= 5 for d in deltas[:3]) and _bits(deltas[3]) <= 6))


def _embed_bytes(deltas, ptr_size):
    """index_posting.rs:590-640, single indexed field."""
    remaining = ptr_size * 8 - (0 if ptr_size == 2 else 1) - 2
    data = 0
    for i, d in enumerate(deltas):
        bits = remaining // (len(deltas) - i)
        remaining -= bits
        data = (data << bits) | int(d)
    if ptr_size == 2:
        return bytes([data & 0xFF, ((data >> 8) & 0xFF) | 0x80 | ((len(deltas) - 1) << 6)])
    return bytes([data & 0xFF, (data >> 8) & 0xFF, ((data >> 16) & 0xFF) | 0x80 | ((len(deltas) - 1) << 5)])


def _key_body(doc_ids, positions):
    """-> (body bytes, rank_position_pointer_range inside the body, pivot, compression type)."""
    count = len(doc_ids)
    blobs, ptrs, cum, pivot = [], [], 0, None
    for p in range(count):
        deltas = positions[p]
        # index_posting.rs:193-200 switches to 3-byte pointers once the key's position bytes reach 32768; a 2-byte pointer carries 15
        # bits, so this writer switches one blob earlier (4 KB of slack) to keep every pointer value representable
        ptr_size = 2 if (cum < 32768 - 4096 and pivot is None) else 3
        if ptr_size == 3 and pivot is None:
            pivot = p
        if _embed(deltas, ptr_size):
            ptrs.append(_embed_bytes(deltas, ptr_size))
            continue
        blob = vint(len(deltas)) + b"".join(vint(int(d)) for d in deltas)
        cum += len(blob)
        blobs.append(blob)
        if ptr_size == 2:
            assert cum < 32768
            ptrs.append(struct.pack("<H", cum & 0x7FFF))
        else:
            ptrs.append(bytes([cum & 0xFF, (cum >> 8) & 0xFF, (cum >> 16) & 0x7F]))
    if pivot is None:
        pivot = count                      # index_posting.rs:196: pointer_pivot_p_docid = posting_count + 1 (count of 2-byte pointers)
    # 2-byte pointers carry 15 bits: the writer switches to 3 bytes before the cumulated size can overflow them
    pos_area = b"".join(reversed(blobs))   # posting 0's blob ends right before the pointer array
    # container choice, compress_postinglist.rs:256-332
    runs = []
    for d in doc_ids:
        if runs and runs[-1][0] + runs[-1][1] + 1 == d:
            runs[-1][1] += 1
        else:
            runs.append([int(d), 0])
    thr = min(count // 2, 65535) if count < 4096 else 2048
    if len(runs) < thr:
        ctype = 3
        cont = struct.pack("<H", len(runs)) + b"".join(struct.pack("<HH", s, l) for s, l in runs)
    elif count < 4096:
        ctype = 1
        cont = np.asarray(doc_ids, dtype="<u2").tobytes()
    else:
        ctype = 2
        bm = np.zeros(8192, dtype=np.uint8)
        ids = np.asarray(doc_ids, dtype=np.int64)
        np.bitwise_or.at(bm, ids >> 3, (1 << (ids & 7)).astype(np.uint8))
        cont = bm.tobytes()
    return pos_area + b"".join(ptrs) + cont, len(pos_area), pivot, ctype


def synth_positions(doc_ids, tfs, rng):
    """delta-encoded positions for each posting: tf strictly positive deltas (first = absolute position)."""
    out = []
    for tf in tfs:
        tf = int(tf)
        kind = rng.integers(0, 4)
        hi = (30, 120, 5000, 60000)[kind]
        out.append([int(x) for x in rng.integers(1, max(2, hi // max(tf, 1)) + 1, size=tf)])
    return out


def deltas_of(positions):
    """absolute ascending positions of one posting -> the reference's coding: first value absolute, then gap - 1 (the phrase matcher
    advances with `pos += next + 1`, add_result.rs:3620-3640)"""
    return [int(positions[0])] + [int(b) - int(a) - 1 for a, b in zip(positions[:-1], positions[1:])]


def write_index_bin(levels, n_docs_total, seed=0):
    """levels: neutral dicts (synth.Level.to_numpy()): level_id ascending from 0, each full (65536 docs) except the last.  A level that
    carries 'positions' (u16 [sum of tfs], posting order) is written with exactly those positions, otherwise with synthetic ones.
    Returns (bytes, positions_sum_normalized)."""
    from seekstorm_b200 import synth
    rng = np.random.default_rng(seed)
    out = [struct.pack("<HH", 6, 1)]
    cum_docs, cum_len = 0, 0
    nseg = 1 << SEGMENT_BITS
    for li, lv in enumerate(levels):
        assert lv["level_id"] == li
        if li == 0:
            out.append(struct.pack("<H", 0))
        dl = np.zeros(65536, dtype=np.uint8)
        dl[:lv["n_docs"]] = lv["doc_len_bytes"]
        out.append(dl.tobytes())
        cum_docs += lv["n_docs"]
        cum_len += int(sum(synth.byte4_to_int(int(b)) for b in lv["doc_len_bytes"]))
        out.append(struct.pack("<QQ", cum_docs, cum_len))
        segs = [[] for _ in range(nseg)]
        offs = lv["posting_offsets"]
        pos_all = lv.get("positions")
        pos_off = None if pos_all is None else np.concatenate([[0], np.cumsum(lv["tfs"].astype(np.int64))])
        for t, key in enumerate(lv["term_keys"]):
            key = int(key)
            segs[(key >> 40) & (nseg - 1)].append((key, t))
        heads, bodies = [], []
        for s in range(nseg):
            segs[s].sort()
            body, hb = b"", b""
            for key, t in segs[s]:
                ids = lv["doc_ids"][offs[t]:offs[t + 1]]
                tfs = lv["tfs"][offs[t]:offs[t + 1]]
                if pos_all is None:
                    plist = synth_positions(ids, tfs, rng)
                else:
                    plist = [deltas_of(pos_all[pos_off[j]:pos_off[j + 1]]) for j in range(int(offs[t]), int(offs[t + 1]))]
                kb, rng_off, pivot, ctype = _key_body(ids, plist)
                ctp = (ctype << 30) | (len(body) + rng_off)
                hb += struct.pack("<QHHHHI", key, len(ids) - 1, int(ids[0]), 0, pivot, ctp)
                body += kb
            heads.append(hb); bodies.append(body)
        out.append(b"".join(struct.pack("<II", len(heads[s]) + len(bodies[s]), len(segs[s])) for s in range(nseg)))
        for s in range(nseg):
            out.append(heads[s]); out.append(bodies[s])
    assert cum_docs == n_docs_total
    return b"".join(out), cum_len


def write_vector_bin(levels):
    """levels: list of (local_ids u16 array, rows f32 [n, dims][, cluster child counts]) per level — `u32 clusters; u32 child_count x clusters;
    records` (vector.rs:1066-1094); without a table: one cluster per level (Clustering::None)."""
    out = []
    for lv in levels:
        ids, rows = lv[0], lv[1]
        n = len(ids)
        counts = [n] if len(lv) < 3 or lv[2] is None else [int(c) for c in lv[2]]
        assert sum(counts) == n
        out.append(struct.pack("<I", len(counts)) + b"".join(struct.pack("<I", c) for c in counts))
        for i in range(n):
            out.append(struct.pack("<HIIffhi", int(ids[i]), 0, 0, 1.0, 1.0, 0, 0))
            out.append(np.asarray(rows[i], dtype="<f4").tobytes())
    return b"".join(out)
