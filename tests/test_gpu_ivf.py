"""GPU: IVF cluster probe (AnnMode::Nprobe / Similaritythreshold / NprobeSimilaritythreshold, vector.rs:1300-1392) through
ssb_vector_add_level_clustered + ssb_search_vector_ex vs the oracle's restatement of the reference's probe: identical result lists
(ids; scores within 1e-4) and identical observed_vector_count for every mode, kernel and similarity."""
import numpy as np
import pytest

from oracle import oracle as O
from helpers_ivf import clustered_levels
import refwriter

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _same(got, want):
    assert [d for d, _ in got] == [d for d, _ in want], (got, want)
    assert np.allclose([s for _, s in got], [s for _, s in want], rtol=RTOL, atol=1e-6)


MODES = [(1, 3, 0.0), (1, 1, 0.0), (1, 1000, 0.0), (1, 0, 0.0), (2, 0, 0.50001), (3, 2, 0.50001), (3, 5, 0.49999)]


@pytest.mark.parametrize("sim", ["cos", "dot", "euc"])
def test_ivf_probe_matches_oracle(sim):
    from seekstorm_b200 import Index, VectorSimilarity
    simv = {"cos": VectorSimilarity.Cosine, "dot": VectorSimilarity.Dot, "euc": VectorSimilarity.Euclidean}[sim]
    osim = {"cos": O.SIM_COSINE, "dot": O.SIM_DOT, "euc": O.SIM_EUCLIDEAN}[sim]
    dims = 64
    levels = clustered_levels(dims, [(5000, 16), (3000, 7), (120, 1), (70, 3)], seed=31)
    ix = Index(0, vector_dims=dims, vector_similarity=simv)
    olevels = []
    for lid, rows, counts in levels:
        ix.add_vector_level(lid, rows, None, counts)
        orows = np.stack([O.normalize(r) for r in rows]) if sim == "cos" else rows
        olevels.append((lid, orows, None, counts))
    rng = np.random.default_rng(5)
    qs = rng.normal(size=(40, dims)).astype(np.float32)
    allr = np.concatenate([lv[1] for lv in levels])
    qs[:20] = allr[rng.choice(len(allr), size=20, replace=False)] + 0.3 * qs[:20]          # queries near real rows: clusters matter
    modes = MODES if sim != "euc" else [(1, 3, 0.0), (1, 1000, 0.0), (2, 0, 40.0), (3, 2, 60.0)]   # Euclidean threshold: distance^2 (pre-map -t)
    for kern, nq in ((0, 40), (0, 5), (4, 40), (7, 40), (1, 16)):
        if sim == "euc" and kern in (4, 7):
            continue
        ix.set_vector_kernel(kern)
        for mode, n_probe, thr in modes:
            for k in (10, 3):
                got, _, observed = ix.search_vector_ex(qs[:nq], k, ann_mode=mode, n_probe=n_probe, cluster_threshold=thr)
                for i in range(nq):
                    oq = O.normalize(qs[i]) if sim == "cos" else qs[i]
                    want, obs = O.search_vector_ivf(olevels, oq, k, osim, mode, n_probe, thr)
                    assert int(observed[i]) == obs, (kern, mode, n_probe, thr, i, int(observed[i]), obs)
                    _same(got[i], want)
    # AnnMode::All is untouched by the cluster tables
    ix.set_vector_kernel(0)
    allrows = np.concatenate([lv[1] for lv in olevels])
    ids = np.concatenate([np.arange(len(lv[1]), dtype=np.uint32) | np.uint32(lv[0] << 16) for lv in olevels])
    got, _, observed = ix.search_vector_ex(qs, 10)
    for i in (0, 7, 39):
        oq = O.normalize(qs[i]) if sim == "cos" else qs[i]
        _same(got[i], O.search_vector(allrows, oq, 10, osim, doc_ids=ids))
        assert int(observed[i]) == len(allrows)
    ix.close()


def test_ivf_through_vector_bin_and_search_mirror():
    """The cluster table of vector.bin reaches the probe (loader), and Index.search carries AnnMode like the reference's SearchMode::Vector."""
    from seekstorm_b200 import Index, VectorSimilarity
    from seekstorm_b200.index import AnnMode, SearchMode
    dims = 32
    levels = clustered_levels(dims, [(900, 9), (400, 4)], seed=8)
    data = refwriter.write_vector_bin([(np.arange(len(rows), dtype=np.uint16), rows, counts) for _, rows, counts in levels])
    ix = Index(0, vector_dims=dims, vector_similarity=VectorSimilarity.Dot)
    assert ix.load_vector_bin(data) == 1300
    olevels = [(i, rows, None, counts) for i, (_, rows, counts) in enumerate(levels)]      # the file numbers its levels 0, 1, ...
    q = levels[0][1][450] + 0.1
    for am, (mode, n_probe, thr) in ((AnnMode.Nprobe(2), (1, 2, 0.0)), (AnnMode.NprobeSimilaritythreshold(3, 0.5001), (3, 3, 0.5001))):
        ro = ix.search("", q, search_mode=SearchMode.Vector(None, am), length=10)
        want, obs = O.search_vector_ivf(olevels, q, 10, O.SIM_DOT, mode, n_probe, thr)
        assert [r.doc_id for r in ro.results] == [d for d, _ in want]
        assert ro.observed_vector_count == obs
    ix.close()


def test_ivf_rejected_on_int8_index_and_bad_tables():
    from seekstorm_b200 import Index, VectorSimilarity
    rows = np.random.default_rng(1).normal(size=(200, 32)).astype(np.float32)
    ix8 = Index(0, vector_dims=32, vector_similarity=VectorSimilarity.Cosine, vector_quantization=1)
    with pytest.raises(Exception):
        ix8.add_vector_level(0, rows, None, [100, 100])
    ix8.add_vector_level(0, rows)
    with pytest.raises(Exception):
        ix8.search_vector_ex(rows[:2], 5, ann_mode=1, n_probe=1)
    ix8.close()
    ix = Index(0, vector_dims=32, vector_similarity=VectorSimilarity.Cosine)
    with pytest.raises(Exception):
        ix.add_vector_level(0, rows, None, [100, 99])          # table does not cover the rows
    with pytest.raises(Exception):
        ix.add_vector_level(0, rows, None, [200, 0])           # empty cluster
    with pytest.raises(Exception):
        ix.search_vector_ex(rows[:2], 5, ann_mode=9)
    ix.close()
