"""BASELINE.json's larger single-GPU shapes at FULL size, through the C-ABI, checked on the GPU itself.

The CPU oracle cannot reach these sizes in test time, so the checker here is an independent restatement of the
reference semantics in plain torch ops (stable sort + rank-within-run zipper; the same closed
form as oracle.np_join, which tests/test_oracle.py pins to cpu_app.c).  torch is TEST INFRASTRUCTURE here: the product
never imports it.  Inputs are generated in HBM by smj_synth_table straight into torch-owned buffers."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    import smj_b200
    if smj_b200.lib().smj_device_count() < 1 or not torch.cuda.is_available():
        pytest.fail("no CUDA device: the gpu-marked tests must run on a B200")
    return smj_b200, torch


def synth(env, rows, cols, seed, kind=0, key_domain=0, total_rows=None):
    smj, torch = env
    t = torch.empty((rows, cols), dtype=torch.int32, device="cuda:0")
    smj.smj.check(smj.lib().smj_synth_table(t.data_ptr(), 0, rows, total_rows or rows, cols, 0, seed, kind, key_domain))
    return t


def torch_reference(torch, t1, t2, v1, v2):
    """select (cell > val on column 0) -> stable sort by column 0 -> zipper join on column 0; returns the joined rows."""
    a = t1[t1[:, 0] > v1]
    b = t2[t2[:, 0] > v2]
    ka, ia = torch.sort(a[:, 0], stable=True)
    kb, ib = torch.sort(b[:, 0], stable=True)
    rank = torch.arange(ka.numel(), device=ka.device) - torch.searchsorted(ka, ka, right=False)
    lb = torch.searchsorted(kb, ka, right=False)
    ub = torch.searchsorted(kb, ka, right=True)
    hit = rank < (ub - lb)
    li = ia[hit]
    ri = ib[(lb + rank)[hit]]
    return torch.cat([a[li], b[ri][:, 1:]], dim=1), a.shape[0], b.shape[0]


def run_on_device(env, t1, t2, v1, v2):
    smj, torch = env
    d1 = smj.Table(t1.data_ptr(), t1.shape[0], t1.shape[1], 1)
    d2 = smj.Table(t2.data_ptr(), t2.shape[0], t2.shape[1], 1)
    out, st = smj.run(d1, d2, on_device=True, keep_output=True, select_val1=v1, select_val2=v2)
    got = torch.empty((out.rows, out.cols), dtype=torch.int32, device="cuda:0")
    if out.rows:
        rt = C.CDLL("libcudart.so.12")
        rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        assert rt.cudaMemcpy(got.data_ptr(), out.data, out.rows * out.cols * 4, 3) == 0   # device to device
    smj.lib().smj_table_free(C.byref(out))
    return got, st


def test_config4_shape_500M_x_100M_8cols_10pct(env):
    """500M x 100M rows, 8 int32 cols, unique keys, 10 % select selectivity (payload-heavy gather), one GPU."""
    smj, torch = env
    n1, n2, cols = 500_000_000, 100_000_000, 8
    t1, t2 = synth(env, n1, cols, 1), synth(env, n2, cols, 2, total_rows=n1)   # both key sets unique in [1, 3 * n1]
    v1 = v2 = int(3 * n1 * 0.9)
    got, st = run_on_device(env, t1, t2, v1, v2)
    want, m1, m2 = torch_reference(torch, t1, t2, v1, v2)
    assert st["rows_selected"] == [m1, m2]
    assert abs(m1 / n1 - 0.1) < 0.01 and abs(m2 / n2 - 0.1) < 0.01
    assert got.shape == want.shape and bool((got == want).all())
    assert got.shape[0] > 3_000_000          # ~ m2 * m1 / (3 n1) matches
    k = got[:, 0]
    assert got.shape[1] == 15 and (k.numel() < 2 or bool((k[1:] > k[:-1]).all()))   # unique keys: strictly ascending


def test_heavy_duplicates_200M_x_200M(env):
    """200M x 200M rows, 4 cols, every key repeated ~10 times on both sides: zip pairing of long equal-key runs at the
    size of BASELINE config 3 (whose Zipf keys are covered at test size in test_gpu_parity.py)."""
    smj, torch = env
    n, cols = 200_000_000, 4
    t1, t2 = synth(env, n, cols, 1, kind=1, key_domain=20_000_000), synth(env, n, cols, 2, kind=1, key_domain=20_000_000)
    got, st = run_on_device(env, t1, t2, 5000, 5000)
    want, m1, m2 = torch_reference(torch, t1, t2, 5000, 5000)
    assert st["rows_selected"] == [m1, m2] and st["rows_joined"] == want.shape[0]
    assert got.shape == want.shape and bool((got == want).all())
    k = got[:, 0]
    assert bool((k[1:] >= k[:-1]).all())


def test_zipf_3M(env, port):
    """Zipf(1.1) keys with a 10 % heavy hitter, against the CPU port (minutes on CPU beyond this size)."""
    smj, torch = env
    t1, t2 = smj.datagen.zipf_table(3_000_000, 4, 21), smj.datagen.zipf_table(2_000_000, 5, 22)
    want, sel, _ = port.run(t1, t2)
    got, st = smj.run(t1, t2)
    assert st["rows_selected"] == list(sel)
    assert got.shape == want.shape and np.array_equal(got, want)
