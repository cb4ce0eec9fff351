"""ctypes binding of libsmj.so -- one Python function per C entry point of include/smj.h.

No compute happens in Python and nothing here falls back to the CPU: if libsmj.so is missing the import of
``lib()`` raises, and without a CUDA device every call raises SmjError(SMJ_ENODEVICE)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
JOIN_ZIP, JOIN_MANY = 0, 1


class SmjError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsmj error {code}: {msg}")
        self.code = code


class Table(C.Structure):
    """smj_table_t"""
    _fields_ = [("data", C.c_void_p), ("rows", C.c_int64), ("cols", C.c_int32), ("on_device", C.c_int32)]


class Config(C.Structure):
    """smj_config_t -- the user.h knobs (reference user.h:1-13)"""
    _fields_ = [("nr_gpus", C.c_int), ("select_col1", C.c_int), ("select_col2", C.c_int),
                ("select_val1", C.c_int64), ("select_val2", C.c_int64),
                ("join_key1", C.c_int), ("join_key2", C.c_int), ("join_mode", C.c_int), ("debug", C.c_int)]


class Stats(C.Structure):
    """smj_stats_t"""
    _fields_ = [("h2d_ms", C.c_double), ("select_ms", C.c_double), ("sort_ms", C.c_double),
                ("exchange_ms", C.c_double), ("merge_ms", C.c_double), ("join_ms", C.c_double),
                ("d2h_ms", C.c_double), ("total_device_ms", C.c_double),
                ("rows_in", C.c_int64 * 2), ("rows_selected", C.c_int64 * 2), ("rows_joined", C.c_int64),
                ("bytes_model", C.c_double), ("bytes_nvlink", C.c_double), ("kernel_launches", C.c_int64),
                ("sort_pass_ms_avg", C.c_double), ("sort_passes", C.c_int32), ("graph_replayed", C.c_int32),
                ("sort_pass_bytes_avg", C.c_double), ("bytes_planned", C.c_double)]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else v
        return d


# every symbol include/smj.h declares (tests/test_abi.py checks the library exports each one)
ABI_SYMBOLS = [
    "smj_config_default", "smj_init", "smj_shutdown", "smj_dist_unique_id", "smj_init_dist", "smj_select", "smj_sort",
    "smj_merge", "smj_join", "smj_run", "smj_join_count", "smj_table_free", "smj_strerror", "smj_last_error",
    "smj_host_alloc", "smj_host_free", "smj_device_alloc", "smj_device_free", "smj_memcpy_h2d", "smj_memcpy_d2h",
    "smj_device_sync", "smj_synth_table", "smj_kernel_launches", "smj_device_count", "smj_version",
    "smj_plan_splitters", "smj_plan_exchange", "smj_csv_parse", "smj_csv_format", "smj_synth_zipf_cdf",
    "smj_plan_fabric", "smj_table_from_i64", "smj_table_to_i64",
]

_lib = None


def lib_path():
    # SMJ_LIB: an alternative build of the same library (kernel-variant experiments; tools/bin/)
    return os.environ.get("SMJ_LIB") or os.path.join(HERE, "libsmj.so")


def build(verbose=False):
    """Compile libsmj.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-s", "-j8", "-C", HERE] if not verbose else ["make", "-j8", "-C", HERE], check=True)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise ImportError(f"{p} is missing: run `make -C {HERE}` (or __graft_entry__.build()); "
                          "there is no Python/CPU fallback for the CUDA engine")
    L = C.CDLL(p)
    TP, CP, SP = C.POINTER(Table), C.POINTER(Config), C.POINTER(Stats)
    L.smj_config_default.argtypes = [CP]
    L.smj_config_default.restype = None
    L.smj_init.argtypes = [CP]
    L.smj_shutdown.restype = None
    L.smj_dist_unique_id.argtypes = [C.c_void_p]
    L.smj_init_dist.argtypes = [CP, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.smj_select.argtypes = [TP, C.c_int, C.c_int64, TP]
    L.smj_sort.argtypes = [TP, C.c_int]
    L.smj_merge.argtypes = [TP, TP, C.c_int, TP]
    L.smj_join.argtypes = [TP, TP, C.c_int, C.c_int, C.c_int, TP]
    L.smj_join_count.argtypes = [TP, TP, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]
    L.smj_run.argtypes = [CP, TP, TP, TP, SP]
    L.smj_table_free.argtypes = [TP]
    L.smj_table_free.restype = None
    L.smj_strerror.argtypes = [C.c_int]
    L.smj_strerror.restype = C.c_char_p
    L.smj_last_error.restype = C.c_char_p
    L.smj_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    L.smj_host_free.argtypes = [C.c_void_p]
    L.smj_host_free.restype = None
    L.smj_device_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    L.smj_device_free.argtypes = [C.c_void_p]
    L.smj_device_free.restype = None
    L.smj_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.smj_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.smj_synth_table.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_uint64, C.c_int,
                                  C.c_int64]
    L.smj_plan_fabric.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int),
                                  C.POINTER(C.c_int64)]
    L.smj_synth_zipf_cdf.argtypes = [C.c_int64, C.c_double, C.c_void_p]
    L.smj_kernel_launches.restype = C.c_int64
    L.smj_csv_parse.argtypes = [C.c_char_p, C.c_size_t, TP]
    L.smj_csv_format.argtypes = [TP, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.smj_plan_splitters.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    L.smj_plan_exchange.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int64)]
    L.smj_table_from_i64.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int, TP]
    L.smj_table_to_i64.argtypes = [TP, C.c_void_p, C.c_int]
    _lib = L
    return L


def check(code):
    if code != 0:
        L = lib()
        raise SmjError(code, f"{L.smj_strerror(code).decode()}: {L.smj_last_error().decode()}")


def default_config(**kw):
    cfg = Config()
    lib().smj_config_default(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


# ------------------------------------------------------------------ tables
def host_table(a):
    """numpy int32 [rows, cols] -> (smj_table_t view, keep-alive array)"""
    a = np.ascontiguousarray(a, dtype=np.int32)
    assert a.ndim == 2
    return Table(a.ctypes.data if a.size else None, a.shape[0], a.shape[1], 0), a


def device_table(a):
    """Copies a numpy table into device memory owned by the caller (free with free())."""
    a = np.ascontiguousarray(a, dtype=np.int32)
    p = C.c_void_p()
    check(lib().smj_device_alloc(C.byref(p), a.nbytes))
    if a.nbytes:
        check(lib().smj_memcpy_h2d(p, a.ctypes.data, a.nbytes))
    return Table(p.value, a.shape[0], a.shape[1], 1)


def synth_device_table(rows, cols, seed, key_col=0, kind=0, key_domain=0, row0=0, total_rows=None):
    p = C.c_void_p()
    check(lib().smj_device_alloc(C.byref(p), rows * cols * 4))
    check(lib().smj_synth_table(p, row0, rows, total_rows or rows, cols, key_col, seed, kind, key_domain))
    return Table(p.value, rows, cols, 1)


def free(t):
    """Releases a caller-owned device table made by device_table()/synth_device_table()."""
    if t.data:
        lib().smj_device_free(t.data)
        t.data = None


def to_numpy(t):
    """Copies a library table (host or device) into a fresh numpy array."""
    out = np.empty((t.rows, t.cols), np.int32)
    if out.nbytes:
        if t.on_device:
            check(lib().smj_memcpy_d2h(out.ctypes.data, t.data, out.nbytes))
        else:
            C.memmove(out.ctypes.data, t.data, out.nbytes)
    return out


def _take(out):
    a = to_numpy(out)
    lib().smj_table_free(C.byref(out))
    return a


def _in(t):
    if isinstance(t, Table):
        return t, None
    return host_table(t)


# ------------------------------------------------------------------ stage entry points
def select(t, col, val, on_device=False):
    """== select.c / cpu_app.c:81-112"""
    tin, keep = _in(t)
    out = Table(None, 0, 0, int(on_device))
    check(lib().smj_select(C.byref(tin), col, int(val), C.byref(out)))
    return _take(out)


def sort(t, key):
    """== sort_dpu.c / cpu_app.c:172-202 (stable); returns a sorted copy for numpy input, sorts Tables in place"""
    if isinstance(t, Table):
        check(lib().smj_sort(C.byref(t), key))
        return t
    a = np.array(t, dtype=np.int32, order="C", copy=True)
    tin, _ = host_table(a)
    check(lib().smj_sort(C.byref(tin), key))
    return a


def merge(a, b, key, on_device=False):
    """== merge_dpu.c + app.c:413-547"""
    ta, ka = _in(a)
    tb, kb = _in(b)
    out = Table(None, 0, 0, int(on_device))
    check(lib().smj_merge(C.byref(ta), C.byref(tb), key, C.byref(out)))
    return _take(out)


def join(l, r, key1, key2, mode=JOIN_ZIP, on_device=False):
    """== join.c / cpu_app.c:204-266"""
    tl, kl = _in(l)
    tr, kr = _in(r)
    out = Table(None, 0, 0, int(on_device))
    check(lib().smj_join(C.byref(tl), C.byref(tr), key1, key2, mode, C.byref(out)))
    return _take(out)


def join_count(l, r, key1, key2, mode=JOIN_ZIP):
    tl, kl = _in(l)
    tr, kr = _in(r)
    n = C.c_int64()
    check(lib().smj_join_count(C.byref(tl), C.byref(tr), key1, key2, mode, C.byref(n)))
    return n.value


def run(t1, t2, cfg=None, on_device=False, keep_output=False, **knobs):
    """== app.c:main select..join / cpu_app.c:336-344.  Returns (result ndarray | Table, stats dict)."""
    cfg = cfg or default_config(**knobs)
    a, ka = _in(t1)
    b, kb = _in(t2)
    out = Table(None, 0, 0, int(on_device))
    st = Stats()
    check(lib().smj_run(C.byref(cfg), C.byref(a), C.byref(b), C.byref(out), C.byref(st)))
    if keep_output:
        return out, st.as_dict()
    return _take(out), st.as_dict()


# ------------------------------------------------------------------ the reference's T = int64_t cells at the boundary
def from_i64(a, on_device=True):
    """numpy int64 [rows, cols] (the reference's T[rows*cols], common.h:1-9) -> library-owned int32 Table via
    smj_table_from_i64 (narrowed on the GPU); raises SmjError(-9) when a cell is not an int32 value."""
    a = np.ascontiguousarray(a, dtype=np.int64)
    assert a.ndim == 2
    out = Table(None, 0, 0, int(on_device))
    check(lib().smj_table_from_i64(a.ctypes.data if a.size else None, a.shape[0], a.shape[1], 0, C.byref(out)))
    return out


def to_i64(t):
    """int32 table (numpy or Table) -> numpy int64 [rows, cols] via smj_table_to_i64."""
    tin, keep = _in(t)
    out = np.empty((tin.rows, tin.cols), np.int64)
    check(lib().smj_table_to_i64(C.byref(tin), out.ctypes.data if out.size else None, 0))
    return out


# ------------------------------------------------------------------ CSV on the GPU
def csv_parse(data):
    """bytes of a whole CSV file -> numpy table via smj_csv_parse; raises SmjError(-8) when the text is irregular."""
    out = Table(None, 0, 0, 1)
    check(lib().smj_csv_parse(data, len(data), C.byref(out)))
    if out.rows <= 0:
        r, c = out.rows, out.cols
        lib().smj_table_free(C.byref(out))
        return np.empty((max(r, 0), c), np.int32)
    return _take(out)


def csv_format(t):
    """table (numpy or Table) -> CSV bytes via smj_csv_format."""
    tin, keep = _in(t)
    p, n = C.c_void_p(), C.c_size_t()
    check(lib().smj_csv_format(C.byref(tin), C.byref(p), C.byref(n)))
    data = C.string_at(p.value, n.value)
    lib().smj_host_free(p)
    return data
