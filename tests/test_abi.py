"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every symbol
include/smj.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def smj():
    import smj_b200
    if not os.path.exists(smj_b200.lib_path()):
        smj_b200.build()
    return smj_b200


def test_header_symbols_all_exported(smj):
    hdr = open(os.path.join(ROOT, "include", "smj.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(smj_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(smj.smj.ABI_SYMBOLS), declared ^ set(smj.smj.ABI_SYMBOLS)
    L = ctypes.CDLL(smj.lib_path())
    for s in declared:
        assert hasattr(L, s), f"libsmj.so does not export {s}"


def test_struct_layouts_match_header(smj):
    assert ctypes.sizeof(smj.Table) == 24
    assert ctypes.sizeof(smj.Config) == 48
    assert ctypes.sizeof(smj.Stats) == 8 * 8 + 5 * 8 + 2 * 8 + 8 + 8 + 8 + 8 + 8   # + sort_pass_bytes_avg, bytes_planned


def test_defaults_are_the_reference_user_h(smj):
    cfg = smj.smj.default_config()
    assert (cfg.select_col1, cfg.select_val1, cfg.select_col2, cfg.select_val2, cfg.join_key1, cfg.join_key2) == \
        (0, 5000, 0, 5000, 0, 0)          # reference user.h:6-13
    assert cfg.nr_gpus == 1 and cfg.join_mode == smj.JOIN_ZIP


def test_no_cpu_fallback(smj):
    if smj.lib().smj_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(smj.SmjError) as e:
        smj.run(np.ones((4, 2), np.int32), np.ones((4, 2), np.int32))
    assert e.value.code == -2   # SMJ_ENODEVICE


def test_sass_is_sm100a(smj):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", smj.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_datagen_unique_keys_and_ranges(smj):
    t = smj.datagen.table(50_000, 4, 1)
    assert len(np.unique(t[:, 0])) == 50_000 and t[:, 0].min() >= 1 and t[:, 0].max() <= 150_000
    assert t[:, 1:].min() >= 1 and t[:, 1:].max() < 150_000
    a = smj.datagen.table(1000, 4, 1, row0=500, total_rows=50_000)
    assert np.array_equal(a, t[500:1500])


def test_zipf_threshold_table(smj):
    """smj_synth_zipf_cdf (host-only): monotone 64-bit thresholds, exact end, Zipf(1.1) head mass; the numpy twin draws from it."""
    cdf = smj.datagen.zipf_cdf(1 << 20)
    assert cdf.dtype == np.uint64 and cdf.shape == (1 << 20,)
    assert bool((cdf[1:] >= cdf[:-1]).all()) and int(cdf[-1]) == 2**64 - 1
    p1 = float(cdf[0]) / 2.0**64
    h = (np.arange(1, (1 << 20) + 1, dtype=np.float64) ** -1.1).sum()
    assert abs(p1 - 1.0 / h) < 1e-9 and 0.12 < p1 < 0.13
    t = smj.datagen.table(100_000, 4, 9, kind=2)
    assert 0.11 < (t[:, 0] == 1).mean() < 0.14 and t[:, 0].min() >= 1 and t[:, 0].max() <= 1 << 20
    u = smj.datagen.table(1000, 4, 9, kind=2, row0=500, total_rows=100_000)
    assert np.array_equal(u[:, 0], t[500:1500, 0])
