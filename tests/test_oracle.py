"""Pins the oracle: port (C restatement) == numpy restatement == the reference's own
cpu_app.c (oracle/_ref) == committed golden vectors."""
import hashlib
import os

import numpy as np
import pytest

from oracle import oracle

CASES = ["g1", "g2", "kat2", "kat3", "kat4"]


def sha_text(s):
    return hashlib.sha256(s.encode()).hexdigest()


def sha_file(p):
    return hashlib.sha256(open(p, "rb").read()).hexdigest()


def rand_tables(rng, n1, n2, c1, c2, dup=False):
    hi = 40 if dup else 50000
    t1 = rng.integers(-20, hi, size=(n1, c1)).astype(np.int32)
    t2 = rng.integers(-20, hi, size=(n2, c2)).astype(np.int32)
    return t1, t2


@pytest.mark.parametrize("case", CASES)
def test_port_matches_golden_csv(case, golden, golden_csv, port, tmp_path):
    g = golden["cases"][case]
    t1 = port.load_csv(golden_csv(f"{case}_data1.csv"))
    t2 = port.load_csv(golden_csv(f"{case}_data2.csv"))
    out, sel, _ = port.run(t1, t2, **g["knobs"])
    assert list(sel) == g["selected"]
    assert out.shape[0] == g["joined"]
    p = str(tmp_path / "out.csv")
    port.save_csv(p, out)
    assert sha_file(p) == g["sha256"]
    # numpy restatement and its text writer agree too
    assert sha_text(oracle.csv_text(oracle.np_run(t1, t2, **g["knobs"]), t1.shape[1] + t2.shape[1] - 1)) == g["sha256"]


def test_golden_inputs_are_the_reference_files(golden, golden_csv):
    for case in ("g1", "g2"):
        for i in (1, 2):
            assert sha_file(golden_csv(f"{case}_data{i}.csv")) == golden["cases"][case]["input_sha256"][i - 1]


def test_g2_stage_dumps(golden, golden_csv, port, tmp_path):
    g = golden["cases"]["g2"]
    for i, (sc, sv, k) in enumerate([(0, 5000, 0), (0, 5000, 0)], start=1):
        t = port.load_csv(golden_csv(f"g2_data{i}.csv"))
        s = port.select(t, sc, sv)
        p = str(tmp_path / "s.csv")
        port.save_csv(p, s)
        assert sha_file(p) == g["stages"][f"select{i}"]
        port.save_csv(p, port.sort(s, k))
        assert sha_file(p) == g["stages"][f"sort{i}"]


@pytest.mark.parametrize("case", ["g2", "kat2", "kat3", "kat4"])
def test_reference_itself_reproduces_golden(case, golden, golden_csv, ref, tmp_path):
    g = golden["cases"][case]
    p = str(tmp_path / "ref.csv")
    j, sel, _ = ref.pipeline_csv(golden_csv(f"{case}_data1.csv"), golden_csv(f"{case}_data2.csv"), p, **g["knobs"])
    assert j == g["joined"] and list(sel) == g["selected"] and sha_file(p) == g["sha256"]


@pytest.mark.parametrize("seed,dup", [(0, False), (1, True), (2, True), (3, False)])
def test_port_equals_reference_stagewise_random(seed, dup, port, ref):
    rng = np.random.default_rng(seed)
    c1, c2 = int(rng.integers(1, 7)), int(rng.integers(1, 7))
    t1, t2 = rand_tables(rng, int(rng.integers(0, 700)), int(rng.integers(0, 700)), c1, c2, dup)
    sc1, sc2, k1, k2 = (int(rng.integers(0, c)) for c in (c1, c2, c1, c2))
    sv1, sv2 = int(rng.integers(-25, 10)), int(rng.integers(-25, 10))
    a_p, a_r = port.select(t1, sc1, sv1), ref.select(t1, sc1, sv1)
    b_p, b_r = port.select(t2, sc2, sv2), ref.select(t2, sc2, sv2)
    assert np.array_equal(a_p, a_r) and np.array_equal(b_p, b_r)
    assert np.array_equal(a_p, oracle.np_select(t1, sc1, sv1))
    sa_p, sa_r = port.sort(a_p, k1), ref.sort(a_r, k1)
    sb_p, sb_r = port.sort(b_p, k2), ref.sort(b_r, k2)
    assert np.array_equal(sa_p, sa_r) and np.array_equal(sb_p, sb_r)
    assert np.array_equal(sa_p, oracle.np_sort(a_p, k1))
    j_p, j_r = port.join(sa_p, sb_p, k1, k2), ref.join(sa_r, sb_r, k1, k2)
    assert np.array_equal(j_p, j_r)
    assert np.array_equal(j_p, oracle.np_join(sa_p, sb_p, k1, k2))
    out, _, _ = port.run(t1, t2, sc1, sv1, sc2, sv2, k1, k2)
    assert np.array_equal(out, j_r)


def test_prop5_heavy_duplicates(port, ref):
    """SURVEY.md 8c PROP-5: Zipf keys, zip semantics (min(cL,cR) rows per key)."""
    rng = np.random.default_rng(7)

    def tab(n, c):
        k = np.minimum(rng.zipf(1.1, n), 2_000_000_000) + 5000
        over = rng.random(n) < 0.1
        k[over] = rng.integers(-50, 6000, over.sum())
        return np.column_stack([k] + [rng.integers(-1000, 300000, n) for _ in range(c - 1)]).astype(np.int32)

    t1, t2 = tab(6000, 4), tab(4000, 5)
    out, sel, _ = port.run(t1, t2)
    a, b = ref.sort(ref.select(t1, 0, 5000), 0), ref.sort(ref.select(t2, 0, 5000), 0)
    assert np.array_equal(out, ref.join(a, b, 0, 0))
    many = port.join(port.sort(port.select(t1, 0, 5000), 0), port.sort(port.select(t2, 0, 5000), 0), 0, 0, mode=1)
    assert many.shape[0] > out.shape[0]
    assert np.array_equal(many, oracle.np_join(a, b, 0, 0, mode=1))


def test_merge_is_stable_sort_of_concat(port):
    rng = np.random.default_rng(11)
    for n1, n2 in [(0, 0), (0, 5), (7, 0), (100, 33), (257, 1000)]:
        a = port.sort(rng.integers(0, 30, (n1, 3)).astype(np.int32), 1)
        b = port.sort(rng.integers(0, 30, (n2, 3)).astype(np.int32), 1)
        assert np.array_equal(port.merge(a, b, 1), oracle.np_merge(a, b, 1))


def test_csv_quirks_match_reference(port, ref, tmp_path):
    """atoi/strtok behaviour of cpu_app.c:15-79: CRLF, +/- signs, spaces, junk suffix, empty fields collapse."""
    p = str(tmp_path / "q.csv")
    with open(p, "w", newline="") as f:
        f.write("a,b,c\r\n 12,+7,-3\r\n4x,,5\n2147483647,-2147483648,0009\n3000000000,1,2\n")
    t, r = port.load_csv(p), ref.load_csv(p)
    r[1, 2] = t[1, 2]  # "4x,,5" has 2 tokens: the reference leaves cell [1][2] as uninitialised malloc memory
    assert np.array_equal(t, r)
    assert t[1].tolist()[:2] == [4, 5]
    q1, q2 = str(tmp_path / "o1.csv"), str(tmp_path / "o2.csv")
    port.save_csv(q1, t)
    ref.save_csv(q2, t)
    assert open(q1, "rb").read() == open(q2, "rb").read()


def test_sort_large_is_stable(port):
    rng = np.random.default_rng(5)
    t = np.column_stack([rng.integers(-1000, 1000, 200000), np.arange(200000)]).astype(np.int32)
    s = port.sort(t, 0)
    assert np.array_equal(s, oracle.np_sort(t, 0))
