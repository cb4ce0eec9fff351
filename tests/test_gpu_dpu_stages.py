"""The B200 stage entry points driven the way app.c drives the DPU programs: one smj_select per DPU row block, one
smj_sort per re-split chunk, the merge tournament through smj_merge, one smj_join per (table-1 chunk, table-2 slice),
with the stage intermediates of the DPU-path emulator (oracle/dpu_path.py, a restatement of app.c:155-688) as the
expected values.  This is the cross-check of the DPU path BASELINE.json asks for, without the UPMEM simulator."""
import numpy as np
import pytest

from oracle import dpu_path

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def smj():
    import smj_b200
    if smj_b200.lib().smj_device_count() < 1:
        pytest.fail("no CUDA device: the gpu-marked tests must run on a B200 (no CPU fallback exists)")
    return smj_b200


def unique_table(rng, n, c, key):
    t = rng.integers(1, 3 * n + 1, size=(n, c)).astype(np.int32)
    t[:, key] = rng.choice(np.arange(1, 3 * n + 1, dtype=np.int64), size=n, replace=False).astype(np.int32)
    return t


@pytest.mark.parametrize("nr_dpus,n1,n2", [(8, 6000, 7001), (5, 2500, 9000), (64, 20_000, 20_000)])
def test_stage_entry_points_reproduce_the_dpu_path(smj, nr_dpus, n1, n2):
    rng = np.random.default_rng(nr_dpus)
    c1, c2, k1, k2, s1, s2 = 4, 3, 1, 0, 0, 2
    t1, t2 = unique_table(rng, n1, c1, k1), unique_table(rng, n2, c2, k2)
    v1, v2 = n1 // 2, n2
    r = dpu_path.run(t1, t2, s1, v1, s2, v2, k1, k2, nr_dpus=nr_dpus)
    tabs, sel_col, sel_val, key = (t1, t2), (s1, s2), (v1, v2), (k1, k2)
    # select.c on every DPU's row block (app.c:221-270)
    for d, (tn, row0, rows) in enumerate(r["blocks"]):
        got = smj.select(tabs[tn][row0:row0 + rows], sel_col[tn], sel_val[tn])
        assert np.array_equal(got, r["select_per_dpu"][d]), f"select on DPU {d}"
    # sort_dpu.c on every re-split chunk (app.c:319-373)
    for tn in (0, 1):
        ndpu = len(r["sorted_chunks"][tn])
        rs = r["selected"][tn].shape[0] // ndpu
        for d in range(ndpu):
            lo, hi = d * rs, ((d + 1) * rs if d < ndpu - 1 else r["selected"][tn].shape[0])
            got = smj.sort(r["selected"][tn][lo:hi], key[tn])
            assert np.array_equal(got, r["sorted_chunks"][tn][d]), f"sort of table {tn} chunk {d}"
    # merge_dpu.c through the tournament (app.c:413-547)
    for tn in (0, 1):
        runs = list(r["sorted_chunks"][tn])
        while len(runs) > 1:
            nxt = [smj.merge(runs[p], runs[p + 1], key[tn]) for p in range(0, len(runs) - 1, 2)]
            if len(runs) % 2:
                nxt.append(runs[-1])
            runs = nxt
        assert np.array_equal(runs[0], r["merged"][tn]), f"merge tournament of table {tn}"
    # join.c on every (chunk, slice) pair the host's range split produces (app.c:585-688), results in DPU order
    parts = []
    for i, (l, s) in enumerate(zip(r["join_chunks"], r["join_slices"])):
        got = smj.join(l, s, k1, k2)
        assert np.array_equal(got, r["join_per_dpu"][i]), f"join on DPU {i}"
        parts.append(got)
    whole, st = smj.run(t1, t2, select_col1=s1, select_val1=v1, select_col2=s2, select_val2=v2, join_key1=k1, join_key2=k2)
    assert np.array_equal(np.concatenate(parts, axis=0), whole)
    assert np.array_equal(whole, r["result"])
    assert st["rows_selected"] == [r["selected"][0].shape[0], r["selected"][1].shape[0]]
