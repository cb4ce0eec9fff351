"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Python face of the two CPU checkers:

* ``port``  -- ctypes binding of oracle/_build/libsmj_oracle.so, our plain-C
  restatement of sort-merge-join/cpu_app.c (int32 cells, 64-bit sizes, stable
  O(n log n) sort).
* ``ref``   -- ctypes binding of oracle/_ref/libref_oracle.so, the reference's own
  cpu_app.c compiled by oracle/Makefile through oracle/ref_wrap.c (int64 cells,
  O(n^2) insertion sort; only usable up to ~1e5 rows).
* ``np_*``  -- a numpy restatement (argsort(kind='stable') + rank-within-run zipper,
  SURVEY.md appendix A) used to cross-check both at small sizes.

PARITY PINNED by tests/test_oracle.py (port == ref == numpy == tests/golden/).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libsmj_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_oracle.so")
REF_BIN = os.path.join(HERE, "_ref", "ref_oracle")

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(ref=True):
    """Compile the checkers (building the checker is not using it)."""
    targets = ["port"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


def _as_table(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    assert a.ndim == 2
    return a


class Port:
    """Plain-C restatement (oracle/smj_oracle.c)."""

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build(ref=False)
        L = self.lib = C.CDLL(PORT_SO)
        L.oracle_select.restype = C.c_int64
        L.oracle_select.argtypes = [_i32p, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_void_p]
        L.oracle_sort.restype = None
        L.oracle_sort.argtypes = [_i32p, C.c_int64, C.c_int, C.c_int]
        L.oracle_merge.restype = None
        L.oracle_merge.argtypes = [_i32p, C.c_int64, _i32p, C.c_int64, C.c_int, C.c_int, _i32p]
        for f in (L.oracle_join_zip,):
            f.restype = C.c_int64
            f.argtypes = [_i32p, C.c_int64, C.c_int, _i32p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.oracle_join_many.restype = C.c_int64
        L.oracle_join_many.argtypes = [_i32p, C.c_int64, C.c_int, _i32p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_int64]
        L.oracle_csv_size.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int64)]
        L.oracle_load_csv.argtypes = [C.c_char_p, C.c_int, C.c_int64, _i32p]
        L.oracle_save_csv.argtypes = [C.c_char_p, C.c_int, C.c_int64, _i32p]
        L.oracle_run.restype = C.c_int64
        L.oracle_run.argtypes = [_i32p, C.c_int64, C.c_int, _i32p, C.c_int64, C.c_int,
                                 C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int,
                                 C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_double)]
        L.oracle_free.argtypes = [C.c_void_p]

    def select(self, t, col, val):
        t = _as_table(t)
        out = np.empty_like(t)
        m = self.lib.oracle_select(t, t.shape[0], t.shape[1], col, int(val), out.ctypes.data)
        return out[:m].copy()

    def sort(self, t, key):
        t = _as_table(t).copy()
        self.lib.oracle_sort(t, t.shape[0], t.shape[1], key)
        return t

    def merge(self, a, b, key):
        a, b = _as_table(a), _as_table(b)
        out = np.empty((a.shape[0] + b.shape[0], a.shape[1]), np.int32)
        self.lib.oracle_merge(a, a.shape[0], b, b.shape[0], a.shape[1], key, out)
        return out

    def join(self, l, r, key1, key2, mode=0):
        l, r = _as_table(l), _as_table(r)
        args = (l, l.shape[0], l.shape[1], r, r.shape[0], r.shape[1], key1, key2)
        tc = l.shape[1] + r.shape[1] - 1
        if mode == 0:
            j = self.lib.oracle_join_zip(*args, None)
            out = np.empty((j, tc), np.int32)
            self.lib.oracle_join_zip(*args, out.ctypes.data)
        else:
            j = self.lib.oracle_join_many(*args, None, 0)
            out = np.empty((j, tc), np.int32)
            self.lib.oracle_join_many(*args, out.ctypes.data, j)
        return out

    def run(self, t1, t2, sel_col1=0, sel_val1=5000, sel_col2=0, sel_val2=5000, key1=0, key2=0, mode=0):
        """Returns (result, (m1, m2), (select_ms, sort_ms, join_ms))."""
        t1, t2 = _as_table(t1), _as_table(t2)
        outp = C.c_void_p()
        sel = (C.c_int64 * 2)()
        ms = (C.c_double * 3)()
        j = self.lib.oracle_run(t1, t1.shape[0], t1.shape[1], t2, t2.shape[0], t2.shape[1],
                                sel_col1, int(sel_val1), sel_col2, int(sel_val2), key1, key2, mode,
                                C.byref(outp), sel, ms)
        tc = t1.shape[1] + t2.shape[1] - 1
        if j:
            buf = (C.c_int32 * (j * tc)).from_address(outp.value)
            out = np.frombuffer(buf, dtype=np.int32).reshape(j, tc).copy()
        else:
            out = np.empty((0, tc), np.int32)
        self.lib.oracle_free(outp)
        return out, (sel[0], sel[1]), tuple(ms)

    def load_csv(self, path):
        cols, rows = C.c_int(), C.c_int64()
        if self.lib.oracle_csv_size(path.encode(), C.byref(cols), C.byref(rows)):
            raise FileNotFoundError(path)
        out = np.zeros((max(rows.value, 0), cols.value), np.int32)
        self.lib.oracle_load_csv(path.encode(), cols.value, rows.value, out)
        return out

    def save_csv(self, path, t):
        t = _as_table(t)
        if self.lib.oracle_save_csv(path.encode(), t.shape[1], t.shape[0], t):
            raise OSError(path)


class Ref:
    """The reference's own cpu_app.c (int64 cells, `int` sizes, O(n^2) sort)."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(f"{REF_SO} missing: run `make -C oracle ref` where /root/reference exists")
        L = self.lib = C.CDLL(REF_SO)
        L.ref_select.restype = C.c_int
        L.ref_select.argtypes = [_i64p, C.c_int, C.c_int, C.c_int, C.c_int64, _i64p]
        L.ref_sort.restype = None
        L.ref_sort.argtypes = [_i64p, C.c_int, C.c_int, C.c_int]
        L.ref_join.restype = C.c_int
        L.ref_join.argtypes = [_i64p, C.c_int, C.c_int, _i64p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_join_cols.restype = C.c_int
        L.ref_join_copy.argtypes = [_i64p]
        L.ref_load_csv.restype = C.c_void_p
        L.ref_load_csv.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_save_csv.argtypes = [C.c_char_p, C.c_int, C.c_int, _i64p]
        L.ref_pipeline_csv.restype = C.c_int
        L.ref_pipeline_csv.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_long, C.c_int, C.c_long,
                                       C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_char_p]

    @staticmethod
    def _t64(t):
        t = np.ascontiguousarray(t, dtype=np.int64)
        assert t.ndim == 2
        return t

    def select(self, t, col, val):
        t = self._t64(t)
        out = np.empty_like(t)
        m = self.lib.ref_select(t, t.shape[0], t.shape[1], col, int(val), out)
        return out[:m].astype(np.int32)

    def sort(self, t, key):
        t = self._t64(t).copy()
        self.lib.ref_sort(t, t.shape[0], t.shape[1], key)
        return t.astype(np.int32)

    def join(self, l, r, key1, key2):
        l, r = self._t64(l), self._t64(r)
        j = self.lib.ref_join(l, l.shape[0], l.shape[1], r, r.shape[0], r.shape[1], key1, key2)
        out = np.empty((j, self.lib.ref_join_cols()), np.int64)
        if j:
            self.lib.ref_join_copy(out)
        return out.astype(np.int32)

    def load_csv(self, path):
        cols, rows = C.c_int(), C.c_int()
        p = self.lib.ref_load_csv(path.encode(), C.byref(cols), C.byref(rows))
        n = max(rows.value, 0) * cols.value
        out = np.frombuffer((C.c_int64 * n).from_address(p), dtype=np.int64).reshape(-1, cols.value).copy() \
            if n else np.empty((0, cols.value), np.int64)
        self.lib.ref_free(p)
        return out.astype(np.int32)

    def save_csv(self, path, t):
        t = self._t64(t)
        self.lib.ref_save_csv(path.encode(), t.shape[1], t.shape[0], t)

    def pipeline_csv(self, f1, f2, out, sel_col1=0, sel_val1=5000, sel_col2=0, sel_val2=5000, key1=0, key2=0,
                     dump_prefix=None):
        ms = (C.c_double * 5)()
        sel = (C.c_int * 2)()
        j = self.lib.ref_pipeline_csv(f1.encode(), f2.encode(), out.encode() if out else None,
                                      sel_col1, int(sel_val1), sel_col2, int(sel_val2), key1, key2, ms, sel,
                                      dump_prefix.encode() if dump_prefix else None)
        return j, (sel[0], sel[1]), dict(zip(("load", "select", "sort", "join", "save"), ms))


def have_ref():
    return os.path.exists(REF_SO)


# ---------------------------------------------------------------- numpy restatement
def np_select(t, col, val):
    """cpu_app.c:81-112 -- strict >, order preserved."""
    t = np.asarray(t)
    return t[t[:, col].astype(np.int64) > int(val)]


def np_sort(t, key):
    """Ordering contract of cpu_app.c:172-202 -- stable ascending."""
    t = np.asarray(t)
    return t[np.argsort(t[:, key], kind="stable")]


def np_merge(a, b, key):
    """Stable merge, a before b on ties == stable sort of the concatenation."""
    return np_sort(np.concatenate([np.asarray(a), np.asarray(b)]), key)


def np_join(l, r, key1, key2, mode=0):
    """cpu_app.c:204-266 in closed form.  l, r sorted by key."""
    l, r = np.asarray(l), np.asarray(r)
    kl, kr = l[:, key1], r[:, key2]
    rcols = [c for c in range(r.shape[1]) if c != key2]
    lb_r = np.searchsorted(kr, kl, "left")
    ub_r = np.searchsorted(kr, kl, "right")
    if mode == 0:
        rank = np.arange(len(kl)) - np.searchsorted(kl, kl, "left")
        hit = rank < (ub_r - lb_r)
        li = np.nonzero(hit)[0]
        ri = (lb_r + rank)[hit]
    else:
        cnt = ub_r - lb_r
        li = np.repeat(np.arange(len(kl)), cnt)
        start = np.repeat(lb_r, cnt)
        off = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        ri = start + off
    return np.concatenate([l[li], r[ri][:, rcols]], axis=1).astype(np.int32)


def np_run(t1, t2, sel_col1=0, sel_val1=5000, sel_col2=0, sel_val2=5000, key1=0, key2=0, mode=0):
    a = np_sort(np_select(t1, sel_col1, sel_val1), key1)
    b = np_sort(np_select(t2, sel_col2, sel_val2), key2)
    return np_join(a, b, key1, key2, mode)


def csv_text(t, cols=None):
    """save_to_csv (cpu_app.c:268-301): header col1..colN, %ld cells, LF."""
    t = np.asarray(t)
    cols = t.shape[1] if cols is None else cols
    head = ",".join(f"col{i}" for i in range(1, cols + 1)) + "\n"
    if t.shape[0] == 0:
        return head
    body = "\n".join(",".join(map(str, row)) for row in t.tolist()) + "\n"
    return head + body
