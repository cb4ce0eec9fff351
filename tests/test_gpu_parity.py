"""GPU parity tests proper: every C-ABI entry point of libsmj.so against the CPU oracle (oracle/, pinned to the
reference's cpu_app.c by tests/test_oracle.py) on the same seeded inputs.  Bit-exact: integer data."""
import hashlib

import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

I32MIN, I32MAX = -2147483648, 2147483647


@pytest.fixture(scope="module")
def smj():
    import smj_b200
    if smj_b200.lib().smj_device_count() < 1:
        pytest.fail("no CUDA device: the gpu-marked tests must run on a B200 (no CPU fallback exists)")
    return smj_b200


def rand_table(rng, n, c, lo=-50, hi=1000):
    return rng.integers(lo, hi, size=(n, c)).astype(np.int32)


def assert_same(got, want, what):
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    if not np.array_equal(got, want):
        bad = np.nonzero((got != want).any(axis=1))[0]
        raise AssertionError(f"{what}: {len(bad)} of {len(want)} rows differ; first at {bad[0]}: got {got[bad[0]]} want {want[bad[0]]}")


# ------------------------------------------------------------------ synthetic generator
@pytest.mark.parametrize("rows,cols,kind,dom,row0,total", [
    (1000, 4, 0, 0, 0, None), (5000, 5, 0, 0, 12345, 100000), (4096, 8, 1, 97, 0, None), (1, 1, 0, 0, 0, None),
    (3000, 4, 0, 2147483646, 7, 2_000_000_000), (200_000, 4, 2, 0, 0, None), (50_000, 3, 2, 5000, 1_000_000, 200_000_000),
])
def test_synth_matches_numpy_twin(smj, rows, cols, kind, dom, row0, total):
    t = smj.synth_device_table(rows, cols, seed=3, kind=kind, key_domain=dom, row0=row0, total_rows=total)
    got = smj.smj.to_numpy(t)
    smj.free(t)
    want = smj.datagen.table(rows, cols, 3, kind=kind, key_domain=dom, row0=row0, total_rows=total)
    assert_same(got, want, "synth")
    if kind == 0 and row0 == 0 and total is None:
        assert len(np.unique(got[:, 0])) == rows
    if kind == 2 and dom == 0:   # Zipf(1.1) over 2^20 keys: key 1 holds ~12 % of the rows
        assert 0.10 < (got[:, 0] == 1).mean() < 0.15 and got[:, 0].min() == 1 and got[:, 0].max() <= 1 << 20


# ------------------------------------------------------------------ select
@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 2047, 2048, 2049, 100_003])
@pytest.mark.parametrize("cols", [1, 4, 5])
def test_select_sizes(smj, port, n, cols):
    rng = np.random.default_rng(n * 7 + cols)
    t = rand_table(rng, n, cols)
    col = int(rng.integers(0, cols))
    assert_same(smj.select(t, col, 300), port.select(t, col, 300), f"select n={n} c={cols}")


@pytest.mark.parametrize("val", [I32MIN - 5, I32MIN, -1, 0, 999, I32MAX - 1, I32MAX, I32MAX + 5])
def test_select_threshold_edges(smj, port, val):
    rng = np.random.default_rng(1)
    t = rand_table(rng, 5000, 3)
    t[::97, 1] = I32MAX
    t[::89, 1] = I32MIN
    assert_same(smj.select(t, 1, val), port.select(t, 1, val), f"select val={val}")


def test_select_selectivity_extremes_and_device_io(smj, port):
    rng = np.random.default_rng(2)
    t = rand_table(rng, 300_000, 8, 0, 1_000_000)
    for val in (-1, 500_000, 900_000, 2_000_000):
        assert_same(smj.select(t, 2, val), port.select(t, 2, val), f"select val={val}")
    d = smj.device_table(t)
    assert_same(smj.select(d, 2, 500_000, on_device=True), port.select(t, 2, 500_000), "select device->device")
    smj.free(d)


# ------------------------------------------------------------------ sort
@pytest.mark.parametrize("n", [0, 1, 2, 33, 4095, 4096, 4097, 50_000, 1_000_003])
def test_sort_sizes(smj, port, n):
    rng = np.random.default_rng(n)
    t = np.column_stack([rng.integers(-10**9, 10**9, n), np.arange(n)]).astype(np.int32)
    assert_same(smj.sort(t, 0), port.sort(t, 0), f"sort n={n}")


@pytest.mark.parametrize("cols,key", [(1, 0), (3, 2), (4, 1), (5, 4), (8, 3)])
def test_sort_is_stable_with_heavy_duplicates(smj, port, cols, key):
    rng = np.random.default_rng(cols * 10 + key)
    t = rand_table(rng, 200_000, cols, -20, 20)
    assert_same(smj.sort(t, key), port.sort(t, key), f"sort dup c={cols} k={key}")


def test_sort_extreme_and_constant_keys(smj, port):
    rng = np.random.default_rng(4)
    t = rand_table(rng, 70_000, 2, I32MIN, I32MAX)
    t[::13, 0] = I32MIN
    t[::17, 0] = I32MAX
    t[::19, 0] = 0
    t[::23, 0] = -1
    assert_same(smj.sort(t, 0), port.sort(t, 0), "sort extremes")
    c = np.column_stack([np.full(30_000, 7), np.arange(30_000)]).astype(np.int32)   # every radix pass is trivial
    assert_same(smj.sort(c, 0), c, "sort constant keys")
    small = np.column_stack([rng.integers(0, 256, 30_000), np.arange(30_000)]).astype(np.int32)   # one live digit
    assert_same(smj.sort(small, 0), port.sort(small, 0), "sort one-digit keys")


def test_sort_device_table_in_place(smj, port):
    rng = np.random.default_rng(5)
    t = rand_table(rng, 123_457, 4, 0, 10**6)
    d = smj.device_table(t)
    smj.sort(d, 0)
    assert_same(smj.smj.to_numpy(d), port.sort(t, 0), "sort device in place")
    smj.free(d)


# ------------------------------------------------------------------ merge
@pytest.mark.parametrize("n1,n2", [(0, 0), (0, 9), (9, 0), (1, 1), (2048, 2048), (2049, 100), (100_000, 37), (150_000, 250_001)])
def test_merge(smj, port, n1, n2):
    rng = np.random.default_rng(n1 + 3 * n2)
    a = port.sort(rand_table(rng, n1, 3, -30, 30 if n1 < 1000 else 5000), 1)
    b = port.sort(rand_table(rng, n2, 3, -30, 30 if n2 < 1000 else 5000), 1)
    assert_same(smj.merge(a, b, 1), port.merge(a, b, 1), f"merge {n1}+{n2}")


def test_merge_disjoint_and_equal_runs(smj, port):
    a = np.column_stack([np.arange(0, 50_000), np.zeros(50_000)]).astype(np.int32)
    b = np.column_stack([np.arange(50_000, 90_000), np.ones(40_000)]).astype(np.int32)
    assert_same(smj.merge(a, b, 0), port.merge(a, b, 0), "merge disjoint a<b")
    assert_same(smj.merge(b, a, 0), port.merge(b, a, 0), "merge disjoint b<a")
    e1 = np.column_stack([np.full(10_000, 5), np.arange(10_000)]).astype(np.int32)
    e2 = np.column_stack([np.full(7_000, 5), -np.arange(7_000)]).astype(np.int32)
    assert_same(smj.merge(e1, e2, 0), port.merge(e1, e2, 0), "merge all-equal (a before b)")


# ------------------------------------------------------------------ join
def sorted_pair(port, rng, n1, n2, c1, c2, k1, k2, lo, hi):
    return port.sort(rand_table(rng, n1, c1, lo, hi), k1), port.sort(rand_table(rng, n2, c2, lo, hi), k2)


@pytest.mark.parametrize("n1,n2,hi", [
    (0, 0, 10), (0, 50, 10), (50, 0, 10), (1, 1, 2), (1000, 1000, 3000), (5000, 3000, 40),
    (100_000, 100_000, 300_000), (100_000, 100_000, 500), (30_000, 200_000, 7), (200_000, 30_000, 7),
])
def test_join_zip(smj, port, n1, n2, hi):
    rng = np.random.default_rng(n1 * 3 + n2 + hi)
    l, r = sorted_pair(port, rng, n1, n2, 3, 4, 1, 2, -5, hi)
    assert_same(smj.join(l, r, 1, 2), port.join(l, r, 1, 2), f"join {n1}x{n2} hi={hi}")
    assert smj.join_count(l, r, 1, 2) == port.join(l, r, 1, 2).shape[0]


def test_join_giant_runs_across_tiles(smj, port):
    """One key repeated far beyond a tile on both sides (zip pairs i-th with i-th), plus neighbours."""
    rng = np.random.default_rng(9)
    def mk(n_big, n_other, c):
        k = np.concatenate([np.full(n_big, 1000), rng.integers(0, 2000, n_other)])
        return port.sort(np.column_stack([k] + [rng.integers(0, 10**6, len(k)) for _ in range(c - 1)]).astype(np.int32), 0)
    for nb1, nb2 in [(20_000, 9_000), (9_000, 20_000), (5_000, 5_000)]:
        l, r = mk(nb1, 7_000, 2), mk(nb2, 11_000, 3)
        assert_same(smj.join(l, r, 0, 0), port.join(l, r, 0, 0), f"join giant {nb1}/{nb2}")


@pytest.mark.parametrize("c1,c2,k1,k2", [(1, 1, 0, 0), (2, 5, 1, 0), (4, 4, 0, 3), (8, 8, 5, 2), (5, 5, 0, 0)])
def test_join_column_layouts(smj, port, c1, c2, k1, k2):
    rng = np.random.default_rng(c1 * 100 + c2 * 10 + k1 + k2)
    l, r = sorted_pair(port, rng, 20_000, 20_000, c1, c2, k1, k2, 0, 30_000)
    assert_same(smj.join(l, r, k1, k2), port.join(l, r, k1, k2), f"join layout {c1},{c2},{k1},{k2}")


def test_join_many_count(smj, port):
    rng = np.random.default_rng(12)
    for n1, n2, hi in [(5000, 4000, 50), (100_000, 80_000, 100_000), (60_000, 60_000, 5)]:
        l, r = sorted_pair(port, rng, n1, n2, 2, 2, 0, 0, 0, hi)
        want = port.lib.oracle_join_many(l, n1, 2, r, n2, 2, 0, 0, None, 0)
        assert smj.join_count(l, r, 0, 0, mode=smj.JOIN_MANY) == want


@pytest.mark.parametrize("n1,n2,hi,c1,c2,k1,k2", [
    (0, 10, 5, 2, 2, 0, 0), (10, 0, 5, 2, 2, 0, 0), (1, 1, 1, 1, 1, 0, 0), (5000, 4000, 50, 2, 3, 0, 1), (100_000, 80_000, 100_000, 4, 4, 0, 0),
    (20_000, 20_000, 5, 3, 2, 2, 0), (3, 70_000, 2, 2, 2, 1, 1), (70_000, 3, 2, 5, 5, 4, 4),
])
def test_join_many_materialised(smj, port, n1, n2, hi, c1, c2, k1, k2):
    """True many-to-many equi-join (extension): every pair of equal keys, order (key, left row, right row)."""
    rng = np.random.default_rng(n1 + 7 * n2 + hi)
    l, r = sorted_pair(port, rng, n1, n2, c1, c2, k1, k2, 0, hi)
    want = port.join(l, r, k1, k2, mode=1)
    got = smj.join(l, r, k1, k2, mode=smj.JOIN_MANY)
    assert_same(got, want, f"many join {n1}x{n2} hi={hi}")
    assert smj.join_count(l, r, k1, k2, mode=smj.JOIN_MANY) == want.shape[0]


def test_join_many_refuses_what_cannot_be_held(smj, port):
    k = np.zeros((3_000_000, 1), np.int32)           # 9e12 pairs
    with pytest.raises(smj.SmjError) as e:
        smj.join(k, k, 0, 0, mode=smj.JOIN_MANY)
    assert e.value.code == -5
    assert smj.join_count(k, k, 0, 0, mode=smj.JOIN_MANY) == 9_000_000_000_000


# ------------------------------------------------------------------ whole pipeline
@pytest.mark.parametrize("case", ["g1", "g2", "kat2", "kat3", "kat4"])
def test_run_golden(smj, port, golden, golden_csv, case, tmp_path):
    g = golden["cases"][case]
    t1 = port.load_csv(golden_csv(f"{case}_data1.csv"))
    t2 = port.load_csv(golden_csv(f"{case}_data2.csv"))
    out, st = smj.run(t1, t2, **{{"sel_col1": "select_col1", "sel_val1": "select_val1", "sel_col2": "select_col2",
                                  "sel_val2": "select_val2", "key1": "join_key1", "key2": "join_key2"}[k]: v
                                 for k, v in g["knobs"].items()})
    assert st["rows_selected"] == g["selected"] and st["rows_joined"] == g["joined"]
    p = str(tmp_path / "out.csv")
    port.save_csv(p, out)
    assert hashlib.sha256(open(p, "rb").read()).hexdigest() == g["sha256"]


@pytest.mark.parametrize("seed", range(6))
def test_run_random_knobs(smj, port, seed):
    rng = np.random.default_rng(100 + seed)
    c1, c2 = int(rng.integers(1, 9)), int(rng.integers(1, 9))
    n1, n2 = int(rng.integers(0, 60_000)), int(rng.integers(0, 60_000))
    hi = [40, 5000, 10**6][seed % 3]
    t1, t2 = rand_table(rng, n1, c1, -hi, hi), rand_table(rng, n2, c2, -hi, hi)
    kn = dict(select_col1=int(rng.integers(0, c1)), select_val1=int(rng.integers(-hi, hi // 2)),
              select_col2=int(rng.integers(0, c2)), select_val2=int(rng.integers(-hi, hi // 2)),
              join_key1=int(rng.integers(0, c1)), join_key2=int(rng.integers(0, c2)))
    want, sel, _ = port.run(t1, t2, kn["select_col1"], kn["select_val1"], kn["select_col2"], kn["select_val2"],
                            kn["join_key1"], kn["join_key2"])
    got, st = smj.run(t1, t2, **kn)
    assert st["rows_selected"] == list(sel)
    assert_same(got, want, f"run seed={seed} knobs={kn}")


@pytest.mark.parametrize("lo1,hi1,lo2,hi2", [
    (7, 8, 7, 8),                          # every key equal: no radix pass runs at all
    (100, 300, 150, 330),                  # one pass (range < 256)
    (-200, 60_000, 10, 40_000),            # two passes, negative keys, different minima
    (1_000_000, 16_000_000, 5, 250),       # three passes left, one pass right (tables take part in different launches)
    (I32MIN, I32MIN + 70_000, I32MAX - 70_000, I32MAX),   # disjoint ranges at the int32 extremes
    (I32MIN, I32MAX, -5, 5),               # the full 32-bit range: four passes
    (0, 1 << 24, 1, (1 << 24) + 2),        # ranges straddling the 3 / 4 pass boundary
])
def test_run_key_range_sort_plans(smj, port, lo1, hi1, lo2, hi2):
    """smj_run sorts digits of (key - smallest surviving key) and runs ceil(bits(range) / 8) passes per table, decided
    on the device (smj_select.cu plan_scan_kernel): every pass count, and the result must not depend on it."""
    rng = np.random.default_rng(abs(lo1) % 1000 + 3)
    n1, n2 = 70_001, 50_003
    t1, t2 = rand_table(rng, n1, 3, 0, 1000), rand_table(rng, n2, 4, 0, 1000)
    t1[:, 1] = rng.integers(lo1, hi1, size=n1, dtype=np.int64, endpoint=(hi1 == I32MAX)).astype(np.int32)
    t2[:, 2] = rng.integers(lo2, hi2, size=n2, dtype=np.int64, endpoint=(hi2 == I32MAX)).astype(np.int32)
    # a shared band of keys so that the join is not empty when the ranges overlap at all
    if max(lo1, lo2) < min(hi1, hi2):
        band = rng.integers(max(lo1, lo2), min(hi1, hi2), size=5000, dtype=np.int64).astype(np.int32)
        t1[:5000, 1] = band
        t2[100:5100, 2] = band[::-1]
    kn = dict(select_col1=0, select_val1=100, select_col2=1, select_val2=200, join_key1=1, join_key2=2)
    want, sel, _ = port.run(t1, t2, 0, 100, 1, 200, 1, 2)
    for _ in range(3):   # eager, graph capture, graph replay
        got, st = smj.run(t1, t2, **kn)
        assert st["rows_selected"] == list(sel)
        assert_same(got, want, f"run key ranges [{lo1},{hi1}) x [{lo2},{hi2})")
    passes = lambda lo, hi: (int(hi - lo).bit_length() + 7) // 8   # an upper bound of the passes the plan may run
    assert 0 <= st["sort_passes"] <= max(passes(lo1, hi1), passes(lo2, hi2))


@pytest.mark.parametrize("n1,n2", [(90_000, 60_000), (50_000, 120_000)])
@pytest.mark.parametrize("overlap", ["all", "none", "half", "dups_left", "dups_both", "fk"])
def test_run_semijoin_filter_cases(smj, port, overlap, n1, n2):
    """The semi-join bitmaps between select and sort (smj_select.cu): rows whose key the other table lacks are dropped
    before the sort, which must not change a single output row.  Covers both table orders (the smaller table is selected
    first and filtered by a separate pass, the larger one probes inside its select kernel), key sets that coincide (the
    device skips the filter pass), disjoint key sets inside the same range (everything is dropped) and duplicate runs."""
    rng = np.random.default_rng(n1 + len(overlap))
    t1, t2 = rand_table(rng, n1, 3, 0, 1000), rand_table(rng, n2, 5, 0, 1000)
    if overlap == "all":
        keys = rng.permutation(max(n1, n2)).astype(np.int32) + 7
        k1, k2 = keys[:n1], rng.permutation(keys)[:n2]
    elif overlap == "none":
        k1 = 2 * rng.permutation(n1).astype(np.int32)            # even keys
        k2 = 2 * rng.permutation(n2).astype(np.int32) + 1        # odd keys, same range
    elif overlap == "half":
        k1 = rng.permutation(2 * n1)[:n1].astype(np.int32)
        k2 = rng.permutation(2 * n1)[:n2].astype(np.int32) if 2 * n1 >= n2 else rng.permutation(n2).astype(np.int32)
    elif overlap == "dups_left":
        k1 = rng.integers(0, n1 // 20, size=n1).astype(np.int32)  # ~20 rows per key on the left
        k2 = rng.permutation(n2).astype(np.int32)                 # unique on the right: min(cL, 1) rows per key
    elif overlap == "dups_both":
        k1 = rng.integers(0, 3000, size=n1).astype(np.int32)
        k2 = rng.integers(1500, 4500, size=n2).astype(np.int32)   # long runs on both sides, half of the keys shared
    else:   # foreign key: every row of the larger table refers to a key of the smaller one, > 3/4 of them survive its
        # select, so the device skips the filter pass over the smaller table
        small = rng.permutation(4 * min(n1, n2))[:min(n1, n2)].astype(np.int32)
        large = rng.choice(small, size=max(n1, n2)).astype(np.int32)
        k1, k2 = (small, large) if n1 <= n2 else (large, small)
    t1[:, 2], t2[:, 0] = k1, k2
    kn = dict(select_col1=0, select_val1=200, select_col2=1, select_val2=100, join_key1=2, join_key2=0)
    if overlap == "fk":   # keep ~95 % of the smaller table so that > 3/4 of the larger one's rows find their key
        kn["select_val1" if n1 <= n2 else "select_val2"] = 50
    want, sel, _ = port.run(t1, t2, kn["select_col1"], kn["select_val1"], kn["select_col2"], kn["select_val2"], 2, 0)
    for _ in range(3):   # eager, graph capture, graph replay
        got, st = smj.run(t1, t2, **kn)
        assert st["rows_selected"] == list(sel)               # the predicate's survivors, not the filter's
        assert_same(got, want, f"run semijoin {overlap} {n1}x{n2}")
    if overlap == "none":
        assert got.shape[0] == 0


def test_run_many_to_many_mode(smj, port):
    """smj_run with join_mode = SMJ_JOIN_MANY (extension): select -> sort -> every pair of equal keys."""
    rng = np.random.default_rng(77)
    t1, t2 = rand_table(rng, 40_000, 3, -30, 400), rand_table(rng, 25_000, 4, -30, 400)
    kn = dict(select_col1=1, select_val1=0, select_col2=0, select_val2=10, join_key1=2, join_key2=3)
    want, sel, _ = port.run(t1, t2, 1, 0, 0, 10, 2, 3, mode=1)
    got, st = smj.run(t1, t2, join_mode=smj.JOIN_MANY, **kn)
    assert st["rows_selected"] == list(sel) and st["rows_joined"] == want.shape[0]
    assert_same(got, want, "run many-to-many")


def test_run_zipf_heavy_duplicates(smj, port):
    """BASELINE config 3 at test size: Zipf(1.1) keys on both sides, zip semantics."""
    t1 = smj.datagen.zipf_table(300_000, 4, 7)
    t2 = smj.datagen.zipf_table(200_000, 4, 8)
    want, sel, _ = port.run(t1, t2)
    got, st = smj.run(t1, t2)
    assert st["rows_selected"] == list(sel)
    assert_same(got, want, "run zipf")


def test_run_device_resident_and_repeatable(smj, port):
    t1, t2 = smj.datagen.table(400_000, 4, 1), smj.datagen.table(400_000, 4, 2)
    want, sel, _ = port.run(t1, t2, 0, 600_000, 0, 600_000, 0, 0)
    d1, d2 = smj.device_table(t1), smj.device_table(t2)
    for _ in range(3):   # workspace reuse must not leak state between runs
        got, st = smj.run(d1, d2, on_device=True, select_val1=600_000, select_val2=600_000)
        assert_same(got, want, "run device resident")
        assert st["kernel_launches"] > 0 and st["total_device_ms"] > 0
    smj.free(d1)
    smj.free(d2)


def test_run_graph_replays_on_new_tables_of_the_same_shape(smj, port):
    """The pipeline graph captured for one pair of device tables is replayed for OTHER tables of the same shape and knobs
    (their addresses reach the kernels through device cells): every call must still give its own tables' result."""
    n = 300_000
    tabs = [(smj.datagen.table(n, 4, s1), smj.datagen.table(n, 4, s2)) for s1, s2 in ((1, 2), (3, 4), (5, 6))]
    wants = [port.run(a, b, 0, 450_000, 0, 450_000, 0, 0)[0] for a, b in tabs]
    devs = [(smj.device_table(a), smj.device_table(b)) for a, b in tabs]
    replayed = []
    for i in (0, 0, 1, 2, 1, 0, 2):      # first call eager, second captures, the others replay whatever the pair
        got, st = smj.run(devs[i][0], devs[i][1], on_device=True, select_val1=450_000, select_val2=450_000)
        assert_same(got, wants[i], f"graph replay, pair {i}")
        replayed.append(st["graph_replayed"])
    if not os.environ.get("SMJ_NO_GRAPH"):
        assert replayed[1:] == [1] * 6, replayed
    # a different shape or knob must not replay the old graph
    got, st = smj.run(devs[0][0], devs[0][1], on_device=True, select_val1=500_000, select_val2=450_000)
    assert st["graph_replayed"] == 0
    assert_same(got, port.run(tabs[0][0], tabs[0][1], 0, 500_000, 0, 450_000, 0, 0)[0], "graph replay, other knob")
    for a, b in devs:
        smj.free(a); smj.free(b)


def test_run_config2_full_size(smj, port):
    """BASELINE config 2 at full size: 10M x 10M rows, 4 int32 cols, unique keys, 50% selectivity."""
    n = 10_000_000
    d1, d2 = smj.synth_device_table(n, 4, 1), smj.synth_device_table(n, 4, 2)
    out, st = smj.run(d1, d2, on_device=True, select_val1=3 * n // 2, select_val2=3 * n // 2)
    smj.free(d1)
    smj.free(d2)
    # size-independent properties
    assert out.shape[1] == 7 and out.shape[0] == st["rows_joined"]
    k = out[:, 0].astype(np.int64)
    assert (np.diff(k) > 0).all(), "unique keys: result strictly ascending by key"
    assert k.min() > 3 * n // 2
    # and the full oracle (the C port finishes this in seconds)
    t1, t2 = smj.datagen.table(n, 4, 1), smj.datagen.table(n, 4, 2)
    want, sel, _ = port.run(t1, t2, 0, 3 * n // 2, 0, 3 * n // 2, 0, 0)
    assert st["rows_selected"] == list(sel)
    assert_same(out, want, "config 2 full size")


def test_join_on_unsorted_input_is_an_error_code_not_a_fault(smj, port):
    """smj_join / smj_join_count / smj_merge take tables SORTED by key (as join.c / merge_dpu.c do).  Unsorted input must not
    take the process down: inconsistent co-ranks are skipped and reported (SMJ_EINVAL), and the library keeps working."""
    rng = np.random.default_rng(3)
    a = rng.integers(0, 1 << 30, size=(300_000, 4)).astype(np.int32)      # unsorted
    b = rng.integers(0, 1 << 30, size=(200_000, 4)).astype(np.int32)
    for call in (lambda: smj.join(a, b, 0, 0), lambda: smj.join_count(a, b, 0, 0, mode=smj.JOIN_MANY), lambda: smj.merge(a, b, 0)):
        try:
            call()                      # garbage in, garbage out is allowed; a device fault is not
        except smj.SmjError as e:
            assert e.code == -1, e
    t1, t2 = smj.datagen.table(50_000, 4, 1), smj.datagen.table(50_000, 4, 2)
    want, sel, _ = port.run(t1, t2, 0, 75_000, 0, 75_000, 0, 0)
    got, st = smj.run(t1, t2, select_val1=75_000, select_val2=75_000)
    assert_same(got, want, "run after unsorted joins")


def test_errors_are_codes_not_exits(smj):
    t = np.zeros((10, 3), np.int32)
    with pytest.raises(smj.SmjError) as e:
        smj.select(t, 3, 0)
    assert e.value.code == -1
    with pytest.raises(smj.SmjError):
        smj.join(t, t, 0, 5)
    with pytest.raises(smj.SmjError):
        smj.run(t, t, select_col1=9)


# ------------------------------------------------------------------ the reference's T = int64_t cells at the boundary
@pytest.mark.parametrize("rows,cols", [(0, 3), (1, 1), (3, 1), (5, 3), (1001, 5), (100_003, 4)])
@pytest.mark.parametrize("to_device", [True, False])
def test_i64_tables_round_trip(smj, rows, cols, to_device):
    """smj_table_from_i64 / smj_table_to_i64: the reference's T[rows*cols] (T = int64_t, common.h:1-9) in and out."""
    rng = np.random.default_rng(rows + cols)
    a = rng.integers(I32MIN, I32MAX, size=(rows, cols), endpoint=True).astype(np.int64)
    if rows:
        a[0, 0], a[-1, -1] = I32MIN, I32MAX
    t = smj.from_i64(a, on_device=to_device)
    assert (t.rows, t.cols, t.on_device) == (rows, cols, int(to_device))
    assert_same(smj.to_numpy(t), a.astype(np.int32), "from_i64")
    back = smj.to_i64(t)
    assert back.dtype == np.int64 and np.array_equal(back, a)
    smj.lib().smj_table_free(smj.smj.C.byref(t))
    assert np.array_equal(smj.to_i64(a.astype(np.int32)), a)          # host int32 table in


def test_i64_tables_device_source_and_chunked_host_source(smj):
    """Host cells travel in 64 MB chunks through two staging buffers (2.5 chunks here); device cells are narrowed in place."""
    C = smj.smj.C
    rows, cols = 5_000_000, 4
    a = (np.arange(rows * cols, dtype=np.int64).reshape(rows, cols) * 2654435761 % (1 << 32)) - (1 << 31)
    t = smj.from_i64(a)
    assert_same(smj.to_numpy(t), a.astype(np.int32), "from_i64, chunked")
    smj.lib().smj_table_free(C.byref(t))
    small = a[:70_001]
    p = C.c_void_p()
    smj.smj.check(smj.lib().smj_device_alloc(C.byref(p), small.nbytes))
    smj.smj.check(smj.lib().smj_memcpy_h2d(p, small.ctypes.data, small.nbytes))
    out = smj.Table(None, 0, 0, 1)
    smj.smj.check(smj.lib().smj_table_from_i64(p, small.shape[0], cols, 1, C.byref(out)))
    assert_same(smj.to_numpy(out), small.astype(np.int32), "from_i64, device source")
    back = C.c_void_p()
    smj.smj.check(smj.lib().smj_device_alloc(C.byref(back), small.nbytes))
    smj.smj.check(smj.lib().smj_table_to_i64(C.byref(out), back, 1))
    got = np.empty_like(small)
    smj.smj.check(smj.lib().smj_memcpy_d2h(got.ctypes.data, back, small.nbytes))
    assert np.array_equal(got, small)
    smj.lib().smj_table_free(C.byref(out))
    smj.lib().smj_device_free(p)
    smj.lib().smj_device_free(back)


@pytest.mark.parametrize("bad", [1 << 31, -(1 << 31) - 1, 1 << 40, -(1 << 62)])
def test_i64_cell_outside_int32_is_refused_not_truncated(smj, bad):
    a = np.arange(40_000 * 3, dtype=np.int64).reshape(40_000, 3)
    a[31_337, 2] = bad
    with pytest.raises(smj.SmjError) as e:
        smj.from_i64(a)
    assert e.value.code == -9 and "[31337][2]" in str(e.value)
    t = smj.from_i64(np.clip(a, I32MIN, I32MAX))                          # the library goes on working
    smj.lib().smj_table_free(smj.smj.C.byref(t))


def test_pipeline_on_the_references_int64_arrays(smj, ref):
    """The reference's own cpu_app.c functions on its own T = int64_t arrays against from_i64 -> smj_run -> to_i64."""
    rng = np.random.default_rng(64)
    t1 = rng.integers(-200, 9000, size=(3000, 4)).astype(np.int64)
    t2 = rng.integers(-200, 9000, size=(2500, 5)).astype(np.int64)
    want = ref.join(ref.sort(ref.select(t1, 0, 5000), 0), ref.sort(ref.select(t2, 0, 5000), 0), 0, 0).astype(np.int64)
    d1, d2 = smj.from_i64(t1), smj.from_i64(t2)
    out, st = smj.run(d1, d2, select_val1=5000, select_val2=5000, on_device=True, keep_output=True)
    got = smj.to_i64(out)
    for t in (d1, d2, out):
        smj.lib().smj_table_free(smj.smj.C.byref(t))
    assert got.dtype == np.int64 and got.shape == want.shape and np.array_equal(got, want)
