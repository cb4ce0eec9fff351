"""The DPU-path emulator (oracle/dpu_path.py: a restatement of app.c's partition / select / sort / merge tournament /
join range split) against the golden vectors cpu_app.c produced and against the port:
for unique join keys the reference's two paths are designed to write the same result.csv (app.c:739-753)."""
import hashlib
import math

import numpy as np
import pytest

from oracle import dpu_path


def unique_table(rng, n, c, key, lo=1, hi=None):
    hi = hi or 3 * n + lo
    t = rng.integers(1, 3 * n + 1, size=(n, c)).astype(np.int32)
    t[:, key] = rng.choice(np.arange(lo, hi, dtype=np.int64), size=n, replace=False).astype(np.int32)
    return t


@pytest.mark.parametrize("case", ["g1", "g2"])
def test_dpu_path_writes_the_cpu_path_result_on_the_bundled_data(port, golden, golden_csv, case, tmp_path):
    """G-1 is the reference's own data1.csv x data2.csv with the default user.h (NR_DPUS 64)."""
    g = golden["cases"][case]
    t1 = port.load_csv(golden_csv(f"{case}_data1.csv"))
    t2 = port.load_csv(golden_csv(f"{case}_data2.csv"))
    r = dpu_path.run(t1, t2, nr_dpus=64)
    assert [r["selected"][0].shape[0], r["selected"][1].shape[0]] == g["selected"]
    assert r["result"].shape[0] == g["joined"]
    p = str(tmp_path / "dpu.csv")
    port.save_csv(p, r["result"])
    assert hashlib.sha256(open(p, "rb").read()).hexdigest() == g["sha256"]
    # the shape of the run the reports describe: 64 DPUs, table 0 on the first half (100 000 rows = 32 x 3 125; the
    # 10 000-row files leave 16 rows for a 33rd block), log2 merge rounds
    assert len(r["blocks"]) == 64 and r["pivot_id"] == (32 if case == "g1" else 33)
    assert len(r["merge_rounds"]) == (5 if case == "g1" else 6)


@pytest.mark.parametrize("nr_dpus", [2, 3, 4, 7, 16, 64])
@pytest.mark.parametrize("seed", [0, 1])
def test_stage_intermediates_match_the_single_pass_path(port, nr_dpus, seed):
    rng = np.random.default_rng(10 * nr_dpus + seed)
    n1, n2 = sorted((int(rng.integers(3000, 9000)), int(rng.integers(3000, 9000))))   # table 0 must end before the last DPU
    c1, c2 = int(rng.integers(2, 6)), int(rng.integers(2, 6))
    k1, k2 = int(rng.integers(0, c1)), int(rng.integers(0, c2))
    s1, s2 = int(rng.integers(0, c1)), int(rng.integers(0, c2))
    t1, t2 = unique_table(rng, n1, c1, k1), unique_table(rng, n2, c2, k2)
    v1, v2 = int(rng.integers(0, n1)), int(rng.integers(0, n2))
    r = dpu_path.run(t1, t2, s1, v1, s2, v2, k1, k2, nr_dpus=nr_dpus)
    # partition: contiguous row blocks that tile each table exactly, table 0 first
    for tn, n in ((0, n1), (1, n2)):
        blk = [(row0, rows) for t, row0, rows in r["blocks"] if t == tn]
        assert blk[0][0] == 0 and sum(rows for _, rows in blk) == n
        assert all(blk[i][0] + blk[i][1] == blk[i + 1][0] for i in range(len(blk) - 1) if blk[i + 1][1])
    # select: the per-DPU survivors, concatenated in DPU order, are the single-pass select
    want_sel = [port.select(t1, s1, v1), port.select(t2, s2, v2)]
    for tn in (0, 1):
        assert np.array_equal(r["selected"][tn], want_sel[tn])
    # sort + merge tournament: every chunk sorted, the final runs equal the single-pass sort
    for tn, k in ((0, k1), (1, k2)):
        for ch in r["sorted_chunks"][tn]:
            assert np.all(np.diff(ch[:, k].astype(np.int64)) >= 0)
        assert np.array_equal(r["merged"][tn], port.sort(want_sel[tn], k))
    nchunks = max(len(r["sorted_chunks"][0]), len(r["sorted_chunks"][1]))
    assert len(r["merge_rounds"]) == max(1, math.ceil(math.log2(nchunks)))
    # join split: the table-2 slices are disjoint, ascending and cover the merged table; every chunk's matches lie in
    # its own slice (unique keys), so the per-DPU joins concatenate to the single-pass join
    assert sum(s.shape[0] for s in r["join_slices"]) == r["merged"][1].shape[0]
    want, _, _ = port.run(t1, t2, s1, v1, s2, v2, k1, k2)
    assert np.array_equal(r["result"], want)
    for i, j in enumerate(r["join_per_dpu"]):
        assert np.array_equal(j, port.join(r["join_chunks"][i], r["join_slices"][i], k1, k2))


def test_fewer_rows_than_dpus_uses_two_dpus(port):
    """app.c:161-165: row_size == 0 -> two DPUs, one per table."""
    rng = np.random.default_rng(5)
    t1, t2 = unique_table(rng, 20, 3, 0, lo=6000), unique_table(rng, 30, 2, 1, lo=6000)
    r = dpu_path.run(t1, t2, key1=0, key2=1, nr_dpus=64)
    assert r["blocks"] == [(0, 0, 20), (1, 0, 30)] and r["pivot_id"] == 1
    want, _, _ = port.run(t1, t2, 0, 5000, 0, 5000, 0, 1)
    assert np.array_equal(r["result"], want)


def test_shapes_the_reference_leaves_undefined_are_refused():
    """Table 1 filling every DPU but the last leaves pivot_id = -1 (app.c:186-199)."""
    with pytest.raises(ValueError):
        dpu_path.partition(1000, 10, 4)      # row_size 252: table 0 needs 4 blocks, only 3 are dealt before the last DPU


def test_negative_select_knob_is_unsigned_on_the_dpu(port):
    """select.c:73-74 holds SELECT_VAL in an unsigned int: -5 becomes 4294967291 and nothing passes; cpu_app.c:88
    compares signed.  The two paths of the reference disagree here; the B200 engine follows cpu_app.c."""
    rng = np.random.default_rng(9)
    t1, t2 = unique_table(rng, 500, 2, 0), unique_table(rng, 500, 2, 0)
    r = dpu_path.run(t1, t2, 0, -5, 0, -5, 0, 0, nr_dpus=4)
    assert r["selected"][0].shape[0] == 0 and r["result"].shape[0] == 0
    r = dpu_path.run(t1, t2, 0, -5, 0, -5, 0, 0, nr_dpus=4, unsigned_select_val=False)
    want, _, _ = port.run(t1, t2, 0, -5, 0, -5, 0, 0)
    assert np.array_equal(r["result"], want) and want.shape[0] > 0
