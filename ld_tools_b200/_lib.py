"""ctypes binding of libldx.so (include/ldx.h).  There is no fallback: if the library or a CUDA
device is missing, importing succeeds (so CPU-only tooling can inspect symbols) but the first
call that needs the GPU raises LdxError."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# LDX_LIB: another build of the same ABI (A/B timing of kernel variants); the default is the in-tree library
LIB_PATH = os.environ.get("LDX_LIB") or os.path.join(_HERE, "lib", "libldx.so")

OK, ERR_ARG, ERR_CUDA, ERR_NOMEM, ERR_EMPTY, ERR_CAPACITY, ERR_STATE, ERR_DATA = 0, -1, -2, -3, -4, -5, -6, -7
MEASURE_R2, MEASURE_DPRIME = 0, 1
ENGINE_AUTO, ENGINE_POPC, ENGINE_MMA = 0, 1, 2
TUNE_MMA_TILE_N, TUNE_MMA_MIN_V, TUNE_MMA_PAIR, TUNE_DEFER_CAP, TUNE_WINDOW_MQ, TUNE_MMA_DIRECT = 1, 2, 3, 4, 5, 6
TEXT_PACKED_ON_DEVICE, TEXT_OUT_ON_DEVICE = 1, 2
AREA_TSV, AREA_JSON, AREA_RSIDS = 0, 1, 2
R2_MASK, R2_INT0, DP_SHIFT, DP_MASK, BELOW_THRES, DP_INT0 = 0x3FFF, 0x8000, 16, 0x3FFF0000, 0x40000000, 0x80000000

HIT_DTYPE = np.dtype([("query", "<i4"), ("row", "<i4"), ("n11", "<i4"), ("packed", "<u4")])
PAIR_HIT_DTYPE = np.dtype([("row", "<i4"), ("col", "<i4"), ("packed", "<u4")])        # ldx_pair_hit
V16_VALUE, V16_BELOW, V16_INT0 = 0x3fff, 0x4000, 0x8000                              # ldx_triangle_values
VCF_ROW_DTYPE = np.dtype([("line_off", "<i8"), ("idnum", "<i8"), ("gt_off", "<i4"), ("id_off", "<i4"), ("ref_off", "<i4"),
                          ("alt_off", "<i4"), ("info_off", "<i4"), ("fmt_off", "<i4"), ("pos", "<i4"), ("ref_len", "<i4"),
                          ("status", "u1"), ("eligible", "u1"), ("multi", "u1"), ("pad", "u1", (5,))])
TRIANGLE_SET_DTYPE = np.dtype([("store", "<u8"), ("rows", "<u8"), ("v", "<i8"), ("dev_packed", "<u8"), ("dev_n11", "<u8")])   # ldx_triangle_set
LD_RESULT_DTYPE = np.dtype([
    ("n_hap", "<i8"), ("n_11", "<i8"), ("n_a1", "<i8"), ("n_a0", "<i8"), ("n_b1", "<i8"), ("n_b0", "<i8"),
    ("d", "<f8"), ("dprime", "<f8"), ("r2", "<f8"), ("p_a", "<f8"), ("p_b", "<f8"),
    ("r2_e4", "<f8"), ("dprime_e4", "<f8"), ("p_a_e4", "<f8"), ("p_b_e4", "<f8"),
    ("dprime_is_int0", "<i4"), ("r2_is_int0", "<i4"),
])


class LdxError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libldx error {code}: {message}")
        self.code = code


_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64
_P = C.POINTER

# name -> argtypes; every function returns int32 except ldx_last_error
SIGNATURES = {
    "ldx_abi_version": [],
    "ldx_device_count": [_P(_i32)],
    "ldx_init": [_i32, _P(_vp)],
    "ldx_destroy": [_vp],
    "ldx_set_stream": [_vp, _vp],
    "ldx_use_own_stream": [_vp],
    "ldx_synchronize": [_vp],
    "ldx_debug_trace": [_vp, _i32, _vp],
    "ldx_set_tuning": [_vp, _i32, _i32],
    "ldx_kernel_timing": [_vp, _i32, _P(C.c_double), _P(_i64)],
    "ldx_sm_count": [_vp, _P(_i32)],
    "ldx_dev_alloc": [_vp, _i64, _P(_vp)],
    "ldx_dev_free": [_vp, _vp],
    "ldx_launch_count": [_vp, _P(_i64)],
    "ldx_calc_ld_lists": [_vp, _vp, _i64, _vp, _i64, _vp],
    "ldx_finalise_counts": [_vp, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp],
    "ldx_store_create": [_vp, _i64, _i32, _P(_vp)],
    "ldx_store_destroy": [_vp],
    "ldx_store_shape": [_vp, _P(_i64), _P(_i32), _P(_i32)],
    "ldx_store_planes_ptr": [_vp, _P(_vp)],
    "ldx_store_pack_gt": [_vp, _i64, _i64, _vp, _i64, _vp, _i64, _i32, _vp],
    "ldx_store_ingest_vcf": [_vp, _vp, _i64, _i32, _P(_vp), _vp, _i64, _P(_i64)],
    "ldx_store_ingest_vcf_file": [_vp, C.c_char_p, _i32, _i64, _i32, _P(_vp), _P(_vp), _P(_i64), _P(_vp), _P(_vp), _P(_i64)],
    "ldx_vcf_copy_prefixes": [_vp, _i64, _vp, _i64, _vp, _i64, _vp],
    "ldx_inflate_gz_file": [C.c_char_p, _i32, _P(_vp), _P(_i64), _P(_i32)],
    "ldx_free_host": [_vp],
    "ldx_store_upload": [_vp, _i64, _i64, _vp],
    "ldx_store_upload_async": [_vp, _i64, _i64, _vp],
    "ldx_store_download": [_vp, _i64, _i64, _vp],
    "ldx_store_save": [_vp, C.c_char_p],
    "ldx_store_load": [_vp, C.c_char_p, _P(_vp)],
    "ldx_store_set_mask": [_vp, _vp],
    "ldx_store_counts": [_vp, _vp, _vp, _P(_i32)],
    "ldx_store_subset": [_vp, _vp, _i32, _P(_vp)],
    "ldx_store_row_counts": [_vp, _vp, _vp, _vp, _P(_i64)],
    "ldx_store_set_annotations": [_vp, _vp, _vp, _vp, _vp],
    "ldx_pairs": [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp],
    "ldx_window": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _i64, _P(_i64), _P(_i64)],
    "ldx_triangle": [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp],
    "ldx_triangle_dev": [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp],
    "ldx_triangle_values": [_vp, _vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _vp],
    "ldx_triangle_hits": [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _i64, _P(_i64)],
    "ldx_triangle_rows": [_vp, _vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _vp],
    "ldx_triangle_rows_dev": [_vp, _vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _vp],
    "ldx_window_dev": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _i64, _vp],
    "ldx_resolve": [_vp, _P(_i64)],
    "ldx_triangle_batch_dev": [_vp, _vp, _i32, _i32, _i32, _i32, _i32],
    "ldx_triangle_text": [_vp, _vp, _i64, _i64, _i64, _i32, _vp, _vp, _i32, _vp, _i64, _P(_i64)],
    "ldx_triangle_table": [_vp, _vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _i64, _P(_i64)],
    "ldx_format_e4": [_i32, _vp],
    "ldx_area_format": [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _i32, _i32, _P(_vp), _P(_i64), _vp],
}

_lib = None


def load():
    """Load libldx.so and declare every prototype of include/ldx.h.  Raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LdxError(ERR_STATE, f"{LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; "
                                      "g.build()'` or `make -C ld_tools_b200/csrc`); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = _i32
        lib.ldx_last_error.argtypes = []
        lib.ldx_last_error.restype = C.c_char_p
        _lib = lib
    return _lib


def check(rc):
    if rc != OK:
        raise LdxError(rc, load().ldx_last_error().decode("utf-8", "replace"))


def ptr(arr):
    """Host pointer of a C-contiguous numpy array (None -> NULL)."""
    return None if arr is None else arr.ctypes.data_as(_vp)
