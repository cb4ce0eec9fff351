"""Small-shape walk over every C-ABI entry point for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
Checks results against the oracle too, so a silent corruption also fails."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smj_b200
from oracle import oracle

port = oracle.Port()
rng = np.random.default_rng(1)
for n, c in [(0, 3), (1, 1), (33, 2), (2049, 4), (5000, 5), (9001, 8), (20000, 33)]:
    t = rng.integers(-50, 300, size=(n, c)).astype(np.int32)
    col = int(rng.integers(0, c))
    assert np.array_equal(smj_b200.select(t, col, 100), port.select(t, col, 100)), ("select", n, c)
    assert np.array_equal(smj_b200.sort(t, col), port.sort(t, col)), ("sort", n, c)
for n1, n2 in [(0, 5), (100, 3000), (8193, 8191), (20000, 100)]:
    a = port.sort(rng.integers(-5, 60, size=(n1, 3)).astype(np.int32), 1)
    b = port.sort(rng.integers(-5, 60, size=(n2, 3)).astype(np.int32), 1)
    assert np.array_equal(smj_b200.merge(a, b, 1), port.merge(a, b, 1)), ("merge", n1, n2)
    assert np.array_equal(smj_b200.join(a, b, 1, 1), port.join(a, b, 1, 1)), ("join", n1, n2)
    if n1 * n2 < 10_000_000:
        assert np.array_equal(smj_b200.join(a, b, 1, 1, mode=smj_b200.JOIN_MANY), port.join(a, b, 1, 1, mode=1)), ("many", n1, n2)
for n1, n2, c1, c2 in [(0, 0, 2, 2), (5000, 7000, 4, 4), (30000, 20000, 5, 3), (70000, 60000, 8, 8)]:
    t1 = rng.integers(-100, 5000, size=(n1, c1)).astype(np.int32)
    t2 = rng.integers(-100, 5000, size=(n2, c2)).astype(np.int32)
    kn = dict(select_col1=c1 - 1, select_val1=0, select_col2=0, select_val2=50, join_key1=0, join_key2=c2 - 1)
    want, sel, _ = port.run(t1, t2, c1 - 1, 0, 0, 50, 0, c2 - 1)
    for _ in range(3):          # third call replays the CUDA graph
        got, st = smj_b200.run(t1, t2, **kn)
        assert np.array_equal(got, want) and st["rows_selected"] == list(sel), ("run", n1, n2)
text = smj_b200.csv_format(t1)
assert np.array_equal(smj_b200.csv_parse(text), t1)
smj_b200.lib().smj_shutdown()
print("sanitize walk ok")
