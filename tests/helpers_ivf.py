"""Synthetic clustered vector levels for the IVF tests: rows in cluster order (cluster = contiguous row range, medoid = first row)."""
import numpy as np


def clustered_levels(dims, spec, seed):
    """spec: [(n_rows, n_clusters)] per level -> [(level_id, rows f32 [n, dims], cluster child counts)].  Cluster members are the medoid
    plus noise, so that the medoid ranking says something about the members (as after the reference's k-medoid clustering)."""
    rng = np.random.default_rng(seed)
    out = []
    for lid, (n, nc) in enumerate(spec):
        cuts = np.sort(rng.choice(np.arange(1, n), size=nc - 1, replace=False)) if nc > 1 else np.array([], dtype=np.int64)
        counts = np.diff(np.concatenate([[0], cuts, [n]])).astype(np.uint32)
        rows = np.empty((n, dims), dtype=np.float32)
        r = 0
        for c in counts:
            centre = rng.normal(size=dims).astype(np.float32) * 2.0
            rows[r] = centre
            rows[r + 1: r + c] = centre + rng.normal(size=(int(c) - 1, dims)).astype(np.float32)
            r += int(c)
        out.append((lid + 2, rows, counts))        # level ids need not start at 0
    return out
