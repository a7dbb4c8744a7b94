"""Facet columns of every FieldType + random FacetFilters for them (tests of the facet-filter path)."""
import numpy as np

from seekstorm_b200 import FacetFilter, _lib


def facet_columns(n, seed):
    """one column per SSB_FACET_* type; floats carry NaN, +-inf and +-0.0, integers their extremes"""
    r = np.random.default_rng(seed)
    f32 = r.normal(0, 10, n).astype(np.float32)
    f64 = r.normal(0, 1e6, n).astype(np.float64)
    for col, special in ((f32, [np.nan, np.inf, -np.inf, 0.0, -0.0]), (f64, [np.nan, np.inf, -np.inf, 0.0, -0.0])):
        idx = r.integers(0, n, 40)
        for j, i in enumerate(idx):
            col[i] = special[j % 5]
    i64 = r.integers(-2**62, 2**62, n, dtype=np.int64)
    i64[r.integers(0, n, 4)] = [np.iinfo(np.int64).min, np.iinfo(np.int64).max, -1, 0]
    u64 = r.integers(0, 2**63, n, dtype=np.uint64) * np.uint64(2) + r.integers(0, 2, n, dtype=np.uint64)
    cols = {
        "u8": r.integers(0, 256, n, dtype=np.uint8), "u16": r.integers(0, 65536, n, dtype=np.uint16),
        "u32": r.integers(0, 2**32, n, dtype=np.uint32), "u64": u64,
        "i8": r.integers(-128, 128, n, dtype=np.int8), "i16": r.integers(-32768, 32768, n, dtype=np.int16),
        "i32": r.integers(-2**31, 2**31, n, dtype=np.int32), "i64": i64,
        "ts": r.integers(1_500_000_000, 1_800_000_000, n, dtype=np.int64),
        "f32": f32, "f64": f64,
        "s16": r.integers(0, 12, n, dtype=np.uint16), "s32": r.integers(0, 300, n, dtype=np.uint32),
    }
    return cols, dict(string_facets=("s16", "s32"), timestamp_facets=("ts",))


def random_filters(cols, seed, n_queries, max_per_query=3):
    """per query 0..max FacetFilters over random facets; ranges are cut at quantiles of the column so that they select 10-90 %"""
    r = np.random.default_rng(seed)
    names = list(cols)
    out = []
    for _ in range(n_queries):
        fl = []
        for name in r.choice(names, int(r.integers(0, max_per_query + 1)), replace=False):
            c = cols[name]
            if name in ("s16", "s32"):
                vals = r.choice(np.unique(c), int(r.integers(1, 6)), replace=False)
                fl.append(FacetFilter(name, values=[int(v) for v in vals] + [10**6]))     # + an id no doc has
            else:
                fin = c[np.isfinite(c)] if c.dtype.kind == "f" else c
                a, b = sorted(r.choice(fin, 2))
                if r.random() < 0.15:
                    a, b = b, a                                              # empty range: nothing passes
                if c.dtype.kind == "f" and r.random() < 0.15:
                    b = np.inf if r.random() < 0.5 else np.nan               # up to +inf / a NaN bound (never contains)
                fl.append(FacetFilter(name, a, b))
        out.append(fl)
    return out


def numpy_pass(cols, fl, doc):
    """the reference's is_facet_filter restated on the typed numpy columns (Range::contains / Vec::contains); True = the doc passes"""
    for f in fl:
        v = cols[f.field][doc]
        if f.values is not None:
            if int(v) not in [int(x) for x in f.values]:
                return False
        else:
            c = cols[f.field]
            if c.dtype.kind == "f":
                x, a, b = float(v), float(f.start), float(f.end)
            else:
                x, a, b = int(v), int(f.start), int(f.end)
            if not (a <= x and x < b):
                return False
    return True


def abi_filters(ix, fl):
    """FacetFilter list -> the C-ABI tuples the oracle wrapper takes, + the set values"""
    offs, arr, sv = ix._encode_filters([fl])
    return [(arr[i].facet, arr[i].kind, arr[i].start, arr[i].end, arr[i].set_first, arr[i].set_count) for i in range(int(offs[1]))], [int(x) for x in sv]
