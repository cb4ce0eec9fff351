"""The many-CTA scan of the tile counts (scan_large_* in csrc/smj_dev.cuh; plan_blocksum / plan_apply in smj_select.cu,
join_blocksum / join_apply in smj_join.cu, tile_blocksum / tile_apply for the stage entry points) replaces the one-CTA scan above 8192 tiles per table, i.e. above ~16 M rows:
sizes only tests/test_gpu_full_size.py reaches.  SMJ_SCAN_CHUNK shrinks the chunk so that small, oracle-checked shapes
take the same kernels.  libsmj.so reads the knob once per process, hence the child process."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys
import numpy as np
sys.path.insert(0, ROOT)
import smj_b200
from oracle import oracle
dry = os.environ.get("SMJ_TEST_DRY") == "1"          # CPU-only rehearsal of this script: oracle half only
if not dry and smj_b200.lib().smj_device_count() < 1:
    print("NO_DEVICE")
    sys.exit(3)
oracle.build(ref=False)
port = oracle.Port()
rng = np.random.default_rng(11)
# (rows1, rows2, cols1, cols2, key spread): > 32 tiles per table and per join, both table orders, ragged sizes, heavy
# duplicates (runs that straddle tiles: the galloping run-start search of join_partition_kernel)
cases = [(200_000, 150_000, 4, 4, 300_000), (70_001, 260_003, 5, 3, 2000), (131_072, 65_536, 8, 2, 50),
         (90_000, 90_000, 4, 4, 7)]
for n1, n2, c1, c2, hi in cases:
    t1 = rng.integers(-hi, hi, size=(n1, c1)).astype(np.int32)
    t2 = rng.integers(-hi, hi, size=(n2, c2)).astype(np.int32)
    v1, v2 = -hi // 2, -hi // 3
    want, sel, _ = port.run(t1, t2, 0, v1, 0, v2, 0, 0)
    if dry:
        continue
    for rep in range(3):                                # eager, graph capture, graph replay
        got, st = smj_b200.run(t1, t2, select_col1=0, select_val1=v1, select_col2=0, select_val2=v2, join_key1=0, join_key2=0)
        assert st["rows_selected"] == list(sel), (n1, n2, rep, st["rows_selected"], sel)
        assert got.shape == want.shape and np.array_equal(got, want), (n1, n2, c1, c2, hi, rep, got.shape, want.shape)
# the stage entry points take the same many-CTA scan (tile_blocksum / tile_apply in smj_select.cu)
t = rng.integers(-500, 500, size=(150_001, 4)).astype(np.int32)
u = rng.integers(-500, 500, size=(90_000, 4)).astype(np.int32)
ws, wt, wu = port.select(t, 1, 100), port.sort(t, 0), port.sort(u, 0)
wm, wj = port.merge(wt, wu, 0), port.join(wt, wu, 0, 0)
if not dry:
    assert np.array_equal(smj_b200.select(t, 1, 100), ws)
    assert np.array_equal(smj_b200.sort(t, 0), wt)
    assert np.array_equal(smj_b200.merge(wt, wu, 0), wm)
    assert np.array_equal(smj_b200.join(wt, wu, 0, 0), wj)
print("SCAN_CHUNK_OK")
'''


def _child(extra_env):
    env = dict(os.environ, SMJ_SCAN_CHUNK="32", **extra_env)
    return subprocess.run([sys.executable, "-c", CHILD.replace("ROOT", repr(ROOT))], env=env, capture_output=True, text=True,
                          timeout=600)


@pytest.mark.gpu
def test_many_cta_scan_at_small_sizes():
    r = _child({})
    assert r.returncode == 0 and "SCAN_CHUNK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_child_script_rehearsal_on_cpu():
    """The oracle half of the child script runs here, so a typo in it cannot be what fails on the GPU box."""
    r = _child({"SMJ_TEST_DRY": "1"})
    assert r.returncode == 0 and "SCAN_CHUNK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
