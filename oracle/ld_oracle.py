"""TEST INFRASTRUCTURE ONLY -- ctypes/numpy front end of the CPU oracle (oracle/ld_oracle.c).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import this.
Parity status: PINNED against the reference's own calc_ld via tests/golden/ (see
tests/test_oracle.py).  Reference citations are relative to /root/reference.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libldoracle.so")

# mirrors `ldo_result` in ld_oracle.c
RESULT_DTYPE = np.dtype([
    ("n_hap", "<i8"), ("n_11", "<i8"), ("n_a1", "<i8"), ("n_a0", "<i8"), ("n_b1", "<i8"),
    ("n_b0", "<i8"),
    ("d", "<f8"), ("dprime", "<f8"), ("r2", "<f8"), ("p_a", "<f8"), ("p_b", "<f8"),
    ("r2_rounded", "<f8"), ("dprime_rounded", "<f8"), ("p_a_rounded", "<f8"),
    ("p_b_rounded", "<f8"),
    ("dprime_is_int0", "<i4"), ("r2_is_int0", "<i4"),
])

_lib = None


def build(force=False):
    """Compile the C restatement (gcc, seconds).  Building the checker is not using it."""
    src = os.path.join(_HERE, "ld_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libldoracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        assert L.ldo_sizeof_result() == RESULT_DTYPE.itemsize
        L.ldo_round4.restype = C.c_double
        L.ldo_round4.argtypes = [C.c_double]
        L.ldo_popc_and3.restype = C.c_int64
        L.ldo_window.restype = C.c_int64
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _i64(x):
    return C.c_int64(int(x))


# ---------------------------------------------------------------- bit planes (numpy only)

def stride_words(n_hap):
    """Store row pitch in 64-bit words: ceil(n_hap/64) rounded up to 16 words (128 B)."""
    words = (n_hap + 63) // 64
    return (words + 15) // 16 * 16


def pack_bits(h01):
    """[V, n_hap] array of 0/1 -> [V, stride] uint64 planes; haplotype h = bit h%64 of word h//64."""
    h01 = np.ascontiguousarray(h01, dtype=np.uint8)
    n_var, n_hap = h01.shape
    stride = stride_words(n_hap)
    padded = np.zeros((n_var, stride * 64), dtype=np.uint8)
    padded[:, :n_hap] = h01
    return np.packbits(padded, axis=1, bitorder="little").view("<u8").reshape(n_var, stride)


def unpack_bits(planes, n_hap):
    planes = np.ascontiguousarray(planes, dtype="<u8")
    bits = np.unpackbits(planes.view(np.uint8).reshape(planes.shape[0], -1), axis=1,
                         bitorder="little")
    return bits[:, :n_hap]


def mask_from_haplotypes(hap_idx, n_hap):
    """Population mask plane from selected haplotype indices (get_sample_names -> columns)."""
    m = np.zeros((1, n_hap), dtype=np.uint8)
    m[0, np.asarray(hap_idx, dtype=np.int64)] = 1
    return pack_bits(m)[0]


def n11_matrix(planes, mask, n_hap):
    """All-pairs (1,1) counts by dense integer matmul -- an arithmetic route independent of popcount."""
    bits = unpack_bits(planes & mask[None, :], n_hap).astype(np.int32)
    return bits @ bits.T


# ---------------------------------------------------------------- C oracle calls

def round4(x):
    return float(lib().ldo_round4(float(x)))


def finalise(n_hap, n_11, n_a1, n_a0, n_b1, n_b0):
    out = np.zeros(1, dtype=RESULT_DTYPE)
    rc = lib().ldo_finalise(_i64(n_hap), _i64(n_11), _i64(n_a1), _i64(n_a0), _i64(n_b1),
                            _i64(n_b0), _p(out))
    if rc:
        raise ZeroDivisionError("division by zero")   # calc_ld.py:33 with an empty pairing
    return out[0]


def finalise_many(n_hap, n_11, n_a1, n_b1):
    """Vector form for biallelic complete data (n_x0 = n_hap - n_x1)."""
    n_11 = np.asarray(n_11, dtype=np.int64)
    out = np.zeros(n_11.shape[0], dtype=RESULT_DTYPE)
    L = lib()
    a1 = np.broadcast_to(np.asarray(n_a1, dtype=np.int64), n_11.shape)
    b1 = np.broadcast_to(np.asarray(n_b1, dtype=np.int64), n_11.shape)
    base = out.ctypes.data
    for k in range(n_11.shape[0]):
        L.ldo_finalise(_i64(n_hap), _i64(n_11[k]), _i64(a1[k]), _i64(n_hap - a1[k]), _i64(b1[k]),
                       _i64(n_hap - b1[k]), C.c_void_p(base + k * RESULT_DTYPE.itemsize))
    return out


def packed_words(n_hap, n_11, n_a1, n_b1):
    """Engine-format packed words (include/ldx.h) for arrays of counts, via the C oracle."""
    n_11 = np.ascontiguousarray(n_11, dtype=np.int32)
    a1 = np.ascontiguousarray(np.broadcast_to(np.asarray(n_a1, dtype=np.int32), n_11.shape))
    b1 = np.ascontiguousarray(np.broadcast_to(np.asarray(n_b1, dtype=np.int32), n_11.shape))
    out = np.zeros(n_11.shape[0], dtype=np.uint32)
    if lib().ldo_finalise_packed_many(_i64(n_hap), _p(n_11), _p(a1), _p(b1), _i64(n_11.shape[0]), _p(out)):
        raise ZeroDivisionError("division by zero")
    return out


def packed_of(res):
    """Packed words of an array of ldo_result records."""
    res = np.atleast_1d(res)
    # the 14-bit fields of the engine's word saturate at 1.6383 (reachable only with entries that are neither 0 nor 1)
    w = np.where(res["r2_is_int0"] != 0, 0x8000, np.minimum(np.rint(res["r2_rounded"] * 10000.0), 16383).astype(np.int64))
    w = w | np.where(res["dprime_is_int0"] != 0, 0x80000000,
                     np.minimum(np.rint(res["dprime_rounded"] * 10000.0), 16383).astype(np.int64) << 16)
    return w.astype(np.uint32)


def encode_genotypes(g):
    """Python genotype sequence -> byte codes 0 / 1 / 255 (anything that is neither == 0 nor == 1)."""
    arr = np.asarray(list(g), dtype=object)
    out = np.full(arr.shape[0], 255, dtype=np.uint8)
    if arr.shape[0]:
        out[np.array([x == 1 for x in arr], dtype=bool)] = 1
        out[np.array([x == 0 for x in arr], dtype=bool)] = 0
    return out


def calc_ld_bytes(g_a, g_b):
    g_a = np.ascontiguousarray(g_a, dtype=np.uint8)
    g_b = np.ascontiguousarray(g_b, dtype=np.uint8)
    out = np.zeros(1, dtype=RESULT_DTYPE)
    rc = lib().ldo_calc_ld_bytes(_p(g_a), _i64(g_a.shape[0]), _p(g_b), _i64(g_b.shape[0]), _p(out))
    if rc:
        raise ZeroDivisionError("division by zero")
    return out[0]


def as_reference_dict(res):
    """ldo_result -> the dict the reference returns, with its int-0 / float types (calc_ld.py:94-99)."""
    return {"r_square": 0 if res["r2_is_int0"] else float(res["r2_rounded"]),
            "d_prime": 0 if res["dprime_is_int0"] else float(res["dprime_rounded"]),
            "var_1_alt_freq": float(res["p_a_rounded"]),
            "var_2_alt_freq": float(res["p_b_rounded"])}


def variant_counts(planes, mask, n_hap):
    planes = np.ascontiguousarray(planes, dtype="<u8")
    mask = np.ascontiguousarray(mask, dtype="<u8")
    n1 = np.zeros(planes.shape[0], dtype=np.int32)
    lib().ldo_variant_counts(_p(planes), _i64(planes.shape[1]), _i64(planes.shape[0]), _p(mask),
                             _i64((n_hap + 63) // 64), _p(n1))
    return n1


def pairs(planes, mask, n_hap, ia, ib):
    planes = np.ascontiguousarray(planes, dtype="<u8")
    mask = np.ascontiguousarray(mask, dtype="<u8")
    ia = np.ascontiguousarray(ia, dtype=np.int64)
    ib = np.ascontiguousarray(ib, dtype=np.int64)
    out = np.zeros(ia.shape[0], dtype=RESULT_DTYPE)
    rc = lib().ldo_pairs(_p(planes), _i64(planes.shape[1]), _p(mask), _i64((n_hap + 63) // 64),
                         _p(ia), _p(ib), _i64(ia.shape[0]), _p(out))
    if rc:
        raise ZeroDivisionError("division by zero")
    return out


def triangle(planes, mask, n_hap, rows):
    """Lower triangle (row > col), packed by rows: index = r*(r-1)//2 + c (ld_triangle.py:133-193)."""
    planes = np.ascontiguousarray(planes, dtype="<u8")
    mask = np.ascontiguousarray(mask, dtype="<u8")
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    v = rows.shape[0]
    out = np.zeros(v * (v - 1) // 2, dtype=RESULT_DTYPE)
    rc = lib().ldo_triangle(_p(planes), _i64(planes.shape[1]), _p(mask), _i64((n_hap + 63) // 64),
                            _p(rows), _i64(v), _p(out))
    if rc:
        raise ZeroDivisionError("division by zero")
    return out


def window(planes, mask, n_hap, pos0, end0, idnum, eligible, q_row, win_start, win_end,
           measure, thres, lo=None, hi=None):
    """One ld_area query by FULL scan of the store (no index shortcuts): returns (rows, results)."""
    planes = np.ascontiguousarray(planes, dtype="<u8")
    mask = np.ascontiguousarray(mask, dtype="<u8")
    pos0 = np.ascontiguousarray(pos0, dtype=np.int32)
    end0 = np.ascontiguousarray(end0, dtype=np.int32)
    idnum = np.ascontiguousarray(idnum, dtype=np.int64)
    eligible = np.ascontiguousarray(eligible, dtype=np.uint8)
    lo = 0 if lo is None else lo
    hi = planes.shape[0] if hi is None else hi
    cap = max(hi - lo, 1)
    rows = np.zeros(cap, dtype=np.int64)
    res = np.zeros(cap, dtype=RESULT_DTYPE)
    n = lib().ldo_window(_p(planes), _i64(planes.shape[1]), _p(mask), _i64((n_hap + 63) // 64),
                         _p(pos0), _p(end0), _p(idnum), _p(eligible), _i64(q_row), _i64(lo),
                         _i64(hi), C.c_int32(int(win_start)), C.c_int32(int(win_end)),
                         C.c_int(int(measure)), C.c_double(float(thres)), _p(rows), _p(res),
                         _i64(cap))
    if n < 0:
        raise ZeroDivisionError("division by zero")
    return rows[:n].copy(), res[:n].copy()


def pack_gt(text, row_off, n_samples):
    """GT text ("a|b" + 1 separator byte per sample) -> planes, per-row status."""
    text = np.ascontiguousarray(text, dtype=np.uint8)
    row_off = np.ascontiguousarray(row_off, dtype=np.int64)
    n_var = row_off.shape[0]
    stride = stride_words(2 * n_samples)
    planes = np.zeros((n_var, stride), dtype="<u8")
    status = np.zeros(n_var, dtype=np.uint8)
    lib().ldo_pack_gt(_p(text), _p(row_off), _i64(n_var), C.c_int32(n_samples), _p(planes),
                      _i64(stride), _p(status))
    return planes, status
