"""numpy twin of csrc/smj_synth.cu: the same closed-form cells, so CPU checkers can rebuild any row range of
a synthetic table that was generated in HBM (replaces the unseeded data/generate_data.py:4-26)."""
import numpy as np

U = np.uint64
_M1, _M2 = U(0xbf58476d1ce4e5b9), U(0x94d049bb133111eb)
_G = U(0x9E3779B97F4A7C15)
INT32_MAX = 2147483647


def mix64(x):
    x = np.asarray(x, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> U(30); x *= _M1
        x ^= x >> U(27); x *= _M2
        x ^= x >> U(31)
    return x


def _perm_bits(x, bits, k0, k1, k2):
    mask = U((1 << bits) - 1)
    s = U(bits // 2 if bits > 1 else 1)
    with np.errstate(over="ignore"):
        x = (x + k0) & mask
        x = (x * (k1 | U(1))) & mask
        x ^= x >> s
        x = (x * (k2 | U(1))) & mask
        x ^= x >> s
        x = (x + k1) & mask
        x = (x * (k0 | U(1))) & mask
        x ^= x >> s
    return x


ZIPF_S, ZIPF_DEFAULT_DOMAIN = 1.1, 1 << 20
_zipf_cdf = {}


def zipf_cdf(dom, s=ZIPF_S):
    """The 64-bit cumulative thresholds of Zipf(s) over [1, dom], from the product's own host function
    (smj_synth_zipf_cdf): the device generator and this twin search the same integers."""
    if (dom, s) not in _zipf_cdf:
        from . import smj as S
        out = np.zeros(dom, np.uint64)
        S.check(S.lib().smj_synth_zipf_cdf(dom, s, out.ctypes.data))
        _zipf_cdf[(dom, s)] = out
    return _zipf_cdf[(dom, s)]


def domains(total_rows, key_domain=0, kind=0):
    dom = key_domain if key_domain > 0 else (ZIPF_DEFAULT_DOMAIN if kind == 2 else 3 * total_rows)
    dom = min(dom, INT32_MAX - 1)
    vdom = max(min(3 * total_rows - 1, INT32_MAX - 1), 1)
    return dom, vdom


def table(rows, cols, seed, key_col=0, kind=0, key_domain=0, row0=0, total_rows=None):
    total_rows = total_rows or rows
    dom, vdom = domains(total_rows, key_domain, kind)
    seed_u = U(seed)
    row = np.arange(row0, row0 + rows, dtype=np.uint64)
    out = np.empty((rows, cols), np.int32)
    with np.errstate(over="ignore"):
        for col in range(cols):
            if col == key_col:
                if kind == 0:
                    bits = 1
                    while (1 << bits) < dom:
                        bits += 1
                    k0, k1, k2 = (mix64(np.array([seed + i], dtype=np.uint64))[0] for i in (1, 2, 3))
                    x = _perm_bits(row.copy(), bits, k0, k1, k2)
                    bad = x >= U(dom)
                    while bad.any():
                        x[bad] = _perm_bits(x[bad], bits, k0, k1, k2)
                        bad = x >= U(dom)
                    v = U(1) + x
                elif kind == 2:   # Zipf(1.1): first rank whose threshold is >= the draw (csrc/smj_synth.cu, kind 2)
                    u = mix64(seed_u * _G + U(0x51ed270b7f4a7c15) + row)
                    v = U(1) + np.searchsorted(zipf_cdf(dom), u, side="left").astype(np.uint64)
                else:
                    v = U(1) + mix64(seed_u * _G + U(0x51ed270b7f4a7c15) + row) % U(dom)
            else:
                v = U(1) + mix64((seed_u * _G) ^ (row * U(cols) + U(col + 1))) % U(vdom)
            out[:, col] = v.astype(np.int64).astype(np.int32)
    return out


def zipf_table(rows, cols, seed, s=1.1, key_col=0, key_offset=5000, clamp=2_000_000_000):
    """Heavy-duplicate keys for the Zipf config (numpy only; used at test sizes)."""
    rng = np.random.default_rng(seed)
    out = rng.integers(1, max(3 * rows, 2), size=(rows, cols)).astype(np.int32)
    out[:, key_col] = (np.minimum(rng.zipf(s, rows), clamp) + key_offset).astype(np.int32)
    return out
