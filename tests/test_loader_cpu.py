"""CPU: the index.bin parser of the library (host code, no GPU) against fixtures manufactured by tests/refwriter.py — every
container type (Array / Bitmap / RLE), embedded and VINT position counts, 2- and 3-byte rank-position pointers."""
import ctypes as C

import numpy as np
import pytest

from seekstorm_b200 import _lib, synth
import refwriter


def _levels_with_special_lists(n_docs=66500, vocab=3000, seed=31):
    # short documents keep the (pure-Python) fixture writer fast; two levels: one full, one partial
    lvs = [synth.gen_level(li, min(65536, n_docs - li * 65536), vocab, seed, "cpu", mean_len=5.0, sigma=0.5, min_len=2, max_len=40).to_numpy()
           for li in range((n_docs + 65535) // 65536)]
    # level 0: add a run-heavy list (RLE), a very long list (Bitmap: >= 4096 postings, many runs) and a list with huge tfs
    lv = lvs[0]
    extra = {
        0x1000: (np.arange(100, 6000, dtype=np.uint16), np.ones(5900, dtype=np.uint16)),                    # one run -> RLE
        0x2000: (np.arange(0, 65536, 3, dtype=np.uint16), (np.arange(21846) % 7 + 1).astype(np.uint16)),    # 21846 postings -> Bitmap
        0x3000: (np.array([5, 9, 70, 40000], dtype=np.uint16), np.array([300, 1, 20000, 2], dtype=np.uint16)),
        0x4000: (np.arange(10, 30000, 2, dtype=np.uint16), np.full(14995, 40, dtype=np.uint16)),             # position bytes > 32 KB -> 3-byte pointers
    }
    keys, offs, ids, tfs = list(lv["term_keys"]), list(lv["posting_offsets"]), [lv["doc_ids"]], [lv["tfs"]]
    for k, (i, t) in extra.items():
        keys.append(np.uint64(k << 3)); ids.append(i); tfs.append(t); offs.append(offs[-1] + len(i))
    lv["term_keys"] = np.array(keys, dtype=np.uint64); lv["posting_offsets"] = np.array(offs, dtype=np.uint32)
    lv["doc_ids"] = np.concatenate(ids); lv["tfs"] = np.concatenate(tfs)
    return lvs, n_docs


def _checksum(levels):
    nseg = 1 << refwriter.SEGMENT_BITS
    h = 1469598103934665603
    M = (1 << 64) - 1
    terms = post = tfsum = 0
    for lv in levels:
        order = sorted(range(len(lv["term_keys"])), key=lambda t: ((int(lv["term_keys"][t]) >> 40) & (nseg - 1), int(lv["term_keys"][t])))
        for t in order:
            h = ((h ^ int(lv["term_keys"][t])) * 1099511628211) & M
            a, b = int(lv["posting_offsets"][t]), int(lv["posting_offsets"][t + 1])
            for d, tf in zip(lv["doc_ids"][a:b], lv["tfs"][a:b]):
                h = ((h ^ ((lv["level_id"] << 32) | (int(d) << 16) | int(tf))) * 1099511628211) & M
                tfsum += int(tf)
            terms += 1; post += b - a
    return terms, post, tfsum, h


def test_index_bin_round_trip_host_parser():
    lvs, n_docs = _levels_with_special_lists()
    data, len_sum = refwriter.write_index_bin(lvs, n_docs, seed=5)
    buf = np.frombuffer(data, dtype=np.uint8)
    prm = _lib.SsbIndexBinParams(1, refwriter.KEY_HEAD_SIZE, refwriter.SEGMENT_BITS, 0)
    out = np.zeros(8, dtype=np.uint64)
    _lib.check(_lib.lib().ssb_index_bin_inspect(buf.ctypes.data, buf.size, C.byref(prm), out.ctypes.data))
    terms, post, tfsum, h = _checksum(lvs)
    assert int(out[0]) == len(lvs) and int(out[1]) == terms and int(out[2]) == post and int(out[3]) == tfsum
    assert int(out[4]) == n_docs and int(out[5]) == len_sum
    assert int(out[6]) == h              # every (key, level, doc id, tf) decoded exactly, in file order


def test_index_bin_rejects_corrupt_files():
    lvs, n_docs = _levels_with_special_lists(n_docs=3000, vocab=200)
    data, _ = refwriter.write_index_bin(lvs, n_docs, seed=6)
    prm = _lib.SsbIndexBinParams(1, 20, 11, 0)
    out = np.zeros(8, dtype=np.uint64)
    L = _lib.lib()
    for cut in (3, 1000, len(data) // 2, len(data) - 1):
        buf = np.frombuffer(data[:cut], dtype=np.uint8).copy()
        assert L.ssb_index_bin_inspect(buf.ctypes.data, buf.size, C.byref(prm), out.ctypes.data) != 0   # status code, no crash
    bad = bytearray(data); bad[0] = 5                      # wrong format version
    buf = np.frombuffer(bytes(bad), dtype=np.uint8)
    assert L.ssb_index_bin_inspect(buf.ctypes.data, buf.size, C.byref(prm), out.ctypes.data) == -5
    for fields, khs in ((2, 20), (1, 21)):
        p2 = _lib.SsbIndexBinParams(fields, khs, 11, 0)
        buf = np.frombuffer(data, dtype=np.uint8)
        assert L.ssb_index_bin_inspect(buf.ctypes.data, buf.size, C.byref(p2), out.ctypes.data) != 0


def test_index_bin_positions_round_trip_host_parser():
    """decode_positions: every term position written by the restated writer (embedded 2- / 3-byte layouts, VINT delta blobs) comes back
    from the library's parser — checksum over all positions in file order — on a corpus of real token sequences (tests/helpers_phrase.py)
    plus postings with long position lists and wide gaps."""
    from helpers_phrase import sequence_corpus
    docs, lvs, _ = sequence_corpus(3000, 120, 13, docs_per_level=65536, mean_len=25)
    lv = lvs[0]
    # extra keys: long position lists (blobs, 3-byte pointers further down the key) and wide gaps (2-byte VINTs)
    extra = {0x5000: ([3, 9], [np.arange(0, 1200, 3), np.array([5, 700, 20000, 65000])]),
             0x6000: (list(range(100, 400)), [np.array([i % 50, 60 + i % 7, 900 + i]) for i in range(300)]),
             # 4200 blobs of 8 bytes push the key past the 2-byte pointer area; the postings behind them use the 3-byte embedded layouts
             # (21 | 10 + 11 | 7 + 7 + 7 | 5 + 5 + 5 + 6 bits) and 3-byte blob pointers
             0x7000: (list(range(0, 8900, 2)), [np.array([200 + i % 90, 4000 + i, 30000 + i]) for i in range(4200)]
                      + [np.array([1000 + i]) for i in range(50)] + [np.array([i % 500, 600 + i % 900]) for i in range(50)]
                      + [np.array([i % 100, 50 + i % 100, 100 + i % 100]) for i in range(50)]
                      + [np.array([i % 20, 25 + i % 20, 50 + i % 20, 80 + i % 40]) for i in range(50)]
                      + [np.array([7, 9000, 9001, 9002, 40000 + i]) for i in range(50)])}
    keys, offs = list(lv["term_keys"]), list(lv["posting_offsets"])
    ids, tfs, pos = [lv["doc_ids"]], [lv["tfs"]], [lv["positions"]]
    for k, (dl, pl) in extra.items():
        keys.append(np.uint64(k << 3)); ids.append(np.array(dl, dtype=np.uint16)); tfs.append(np.array([len(p) for p in pl], dtype=np.uint16))
        pos.append(np.concatenate(pl).astype(np.uint16)); offs.append(offs[-1] + len(dl))
    lv["term_keys"] = np.array(keys, dtype=np.uint64); lv["posting_offsets"] = np.array(offs, dtype=np.uint32)
    lv["doc_ids"] = np.concatenate(ids); lv["tfs"] = np.concatenate(tfs); lv["positions"] = np.concatenate(pos)
    data, len_sum = refwriter.write_index_bin(lvs, 3000, seed=9)
    buf = np.frombuffer(data, dtype=np.uint8)
    out = np.zeros(8, dtype=np.uint64)
    prm = _lib.SsbIndexBinParams(1, refwriter.KEY_HEAD_SIZE, refwriter.SEGMENT_BITS, 1)
    _lib.check(_lib.lib().ssb_index_bin_inspect(buf.ctypes.data, buf.size, C.byref(prm), out.ctypes.data))
    terms, post, tfsum, h = _checksum(lvs)
    assert int(out[1]) == terms and int(out[2]) == post and int(out[3]) == tfsum and int(out[6]) == h
    # expected position checksum: positions in FILE order (segments ascending, keys ascending inside a segment)
    nseg = 1 << refwriter.SEGMENT_BITS
    M = (1 << 64) - 1
    hp = 1469598103934665603
    poff = np.concatenate([[0], np.cumsum(lv["tfs"].astype(np.int64))])
    order = sorted(range(len(lv["term_keys"])), key=lambda t: ((int(lv["term_keys"][t]) >> 40) & (nseg - 1), int(lv["term_keys"][t])))
    for t in order:
        a, b = int(lv["posting_offsets"][t]), int(lv["posting_offsets"][t + 1])
        for p in lv["positions"][poff[a]:poff[b]]:
            hp = ((hp ^ int(p)) * 1099511628211) & M
    assert int(out[7]) == hp
    prm0 = _lib.SsbIndexBinParams(1, refwriter.KEY_HEAD_SIZE, refwriter.SEGMENT_BITS, 0)
    _lib.check(_lib.lib().ssb_index_bin_inspect(buf.ctypes.data, buf.size, C.byref(prm0), out.ctypes.data))
    assert int(out[7]) == 0 and int(out[6]) == h                    # positions off: exactly the previous behaviour
