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


def test_config4_shape_500M_x_100M_8cols_10pct(env, port):
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
    # pinned to the oracle: the C port (tests/test_oracle.py ties it to cpu_app.c) on the rows that pass the predicate
    # (~50 M + ~10 M; select keeps order, so select(survivors) == select(table) and everything downstream is the same)
    a = t1[t1[:, 0] > v1].cpu().numpy()
    b = t2[t2[:, 0] > v2].cpu().numpy()
    del want
    want_o, sel_o, _ = port.run(a, b, 0, v1, 0, v2, 0, 0)
    assert list(sel_o) == [m1, m2]
    g = got.cpu().numpy()
    assert g.shape == want_o.shape and np.array_equal(g, want_o)


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


def test_config3_zipf_200M_x_200M(env, port):
    """BASELINE config 3 at full size: 200M x 200M rows, 4 cols, Zipf(1.1) keys over 2^20 values (key 1 holds ~12 % = a
    ~25 M-row run on each side), every row selected.  Whole result against the torch closed form; and PINNED TO THE ORACLE
    per key: zip pairing is independent per key (cpu_app.c:213-218), so the rows of a key subset -- the heavy hitter, a
    few mid keys, a block of tail keys -- fed to the C port must give exactly the result rows of those keys."""
    smj, torch = env
    n, cols = 200_000_000, 4
    t1, t2 = synth(env, n, cols, 1, kind=2), synth(env, n, cols, 2, kind=2)
    got, st = run_on_device(env, t1, t2, 0, 0)
    assert st["rows_selected"] == [n, n]
    k = got[:, 0]
    assert bool((k[1:] >= k[:-1]).all())
    c1 = torch.bincount(t1[:, 0].long(), minlength=(1 << 20) + 1)
    c2 = torch.bincount(t2[:, 0].long(), minlength=(1 << 20) + 1)
    assert got.shape[0] == int(torch.minimum(c1, c2).sum()) == st["rows_joined"]
    assert int(c1[1]) > 20_000_000 and int(c2[1]) > 20_000_000
    want, _, _ = torch_reference(torch, t1, t2, 0, 0)
    assert got.shape == want.shape and bool((got == want).all())
    del want

    def subset(x):
        kk = x[:, 0]
        return (kk == 1) | ((kk >= 50) & (kk <= 60)) | ((kk >= 100_000) & (kk <= 120_000))
    a, b = t1[subset(t1)].cpu().numpy(), t2[subset(t2)].cpu().numpy()
    want_o, _, _ = port.run(a, b, 0, 0, 0, 0, 0, 0)
    g = got[subset(got)].cpu().numpy()
    assert g.shape == want_o.shape and np.array_equal(g, want_o)
    # the many-to-many COUNT of the same inputs (smj_join_count on key-sorted tables), against the per-key products
    many = int((c1.double() * c2.double()).sum().item())
    s1, s2 = t1[torch.sort(t1[:, 0], stable=True)[1]].contiguous(), t2[torch.sort(t2[:, 0], stable=True)[1]].contiguous()
    torch.cuda.synchronize()   # torch's stream and the library's are unrelated: the sorted copies must be complete before libsmj reads them
    cnt = smj.join_count(smj.Table(s1.data_ptr(), n, cols, 1), smj.Table(s2.data_ptr(), n, cols, 1), 0, 0, mode=smj.JOIN_MANY)
    assert abs(cnt - many) <= many * 1e-12 and cnt == int((c1.long() * c2.long()).sum().item())


def test_zipf_3M(env, port):
    """Zipf(1.1) keys with a 10 % heavy hitter, against the CPU port (minutes on CPU beyond this size)."""
    smj, torch = env
    t1, t2 = smj.datagen.zipf_table(3_000_000, 4, 21), smj.datagen.zipf_table(2_000_000, 5, 22)
    want, sel, _ = port.run(t1, t2)
    got, st = smj.run(t1, t2)
    assert st["rows_selected"] == list(sel)
    assert got.shape == want.shape and np.array_equal(got, want)
