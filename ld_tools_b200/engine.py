"""Thin object layer over the C ABI (include/ldx.h): Context and Store.

Everything numeric happens in libldx.so on the GPU; this module only marshals numpy arrays,
turns status codes into exceptions and decodes the packed result words into the Python objects
the reference returns (int `0` vs float, calc_ld.py:68-69/:89-90; round(x, 4), calc_ld.py:94-97).
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import (BELOW_THRES, DP_INT0, DP_MASK, DP_SHIFT, ENGINE_AUTO, ENGINE_MMA, ENGINE_POPC,  # noqa: F401
                   HIT_DTYPE, PAIR_HIT_DTYPE, LD_RESULT_DTYPE, MEASURE_DPRIME, MEASURE_R2, R2_INT0, R2_MASK, TRIANGLE_SET_DTYPE, VCF_ROW_DTYPE,
                   LdxError, check, ptr)

MEASURES = {"r_square": MEASURE_R2, "d_prime": MEASURE_DPRIME}   # the CLI's -l choices


def stride_words(n_hap):
    """Row pitch of a store in 64-bit words: ceil(n_hap/64) rounded up to 16 (128-byte rows)."""
    return ((n_hap + 63) // 64 + 15) // 16 * 16


def threshold_e4(thres):
    """Smallest integer n with n / 10000.0 >= thres: the rounded-value test of ld_area.py:248
    (`trg_vals[measure] < thres -> skip`) restated on round(x,4)*10^4 integers, exactly."""
    n = int(np.floor(float(thres) * 10000.0))
    while n / 10000.0 >= thres:
        n -= 1
    while n / 10000.0 < thres:
        n += 1
    return max(n, 0)


def r2_e4(packed):
    return (np.asarray(packed) & R2_MASK).astype(np.int32)


def dprime_e4(packed):
    return ((np.asarray(packed) & DP_MASK) >> DP_SHIFT).astype(np.int32)


def r2_value(word):
    """Packed word -> the object calc_ld returns under 'r_square' (int 0 or a rounded float)."""
    word = int(word)
    return 0 if word & R2_INT0 else (word & R2_MASK) / 10000.0


def dprime_value(word):
    word = int(word)
    return 0 if word & DP_INT0 else ((word & DP_MASK) >> DP_SHIFT) / 10000.0


def measure_value(word, measure):
    return r2_value(word) if measure in ("r_square", MEASURE_R2) else dprime_value(word)


FORMAT_CELL_MAX = 7       # bytes of the longest matrix cell with its separator ("1.6383" + tab)


def format_e4(value_e4):
    """str(value_e4 / 10000.0) as libldx prints it (ldx_format_e4: host code, no device needed)."""
    buf = C.create_string_buffer(8)
    check(_lib.load().ldx_format_e4(int(value_e4), buf))
    return buf.value.decode("ascii")


def _measure_code(measure):
    return MEASURES[measure] if isinstance(measure, str) else int(measure)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def area_format(lib, hits, q_row, blob, blob_off, rows, p_e4, fmt, overrides=None, threads=0):
    """The ld_area writers' body rows for all queries of a window scan (ldx_area_format: host code in libldx, all cores).
    hits: HIT_DTYPE array sorted by (query, row); rows / blob / blob_off: the records' field offsets and fixed columns
    (Store.ingest_vcf, Store.vcf_fixed_columns); p_e4: round(alt freq, 4) * 10^4 per store row.
    -> (uint8 array of the text, int64 query_off[nq + 1]): query k's rows are text[query_off[k]:query_off[k + 1]]."""
    hits = np.ascontiguousarray(hits, dtype=HIT_DTYPE)
    q_row = _i64(q_row)
    rows = np.ascontiguousarray(rows, dtype=VCF_ROW_DTYPE)
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    blob_off = _i64(blob_off)
    p_e4 = np.ascontiguousarray(p_e4, dtype=np.int32)
    alt = None if overrides is None else np.ascontiguousarray(overrides, dtype=np.int32).reshape(-1, 3)      # {alt_freq, r2, D'} * 10^4, < 0: default
    assert alt is None or alt.shape[0] == hits.shape[0]
    qoff = np.zeros(q_row.shape[0] + 1, dtype=np.int64)
    n, p = C.c_int64(), C.c_void_p()
    check(lib.ldx_area_format(ptr(hits), hits.shape[0], ptr(q_row), q_row.shape[0], ptr(blob), ptr(blob_off), ptr(rows), rows.shape[0], ptr(p_e4),
                              ptr(alt), int(fmt), int(threads), C.byref(p), C.byref(n), ptr(qoff)))
    return _lib_text(lib, p, n.value), qoff


class _Allocation:
    """Host memory malloc'ed by the library, released with ldx_free_host when the last array viewing it goes away."""

    def __init__(self, lib, p):
        self.lib, self.p = lib, p

    def __del__(self):
        if self.p is not None and self.p.value:
            self.lib.ldx_free_host(self.p)
        self.p = None


class _LibText(np.ndarray):
    """uint8 view of library-owned text; slices of it keep the allocation alive."""
    _owner = None

    def __array_finalize__(self, obj):
        self._owner = getattr(obj, "_owner", None)


def _lib_text(lib, p, nbytes):
    owner = _Allocation(lib, p)
    if not nbytes:
        return np.zeros(0, dtype=np.uint8)
    out = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,)).view(_LibText)
    out._owner = owner
    return out


def device_count():
    """CUDA devices visible to the library (ldx_device_count): no context is created."""
    n = C.c_int32()
    check(_lib.load().ldx_device_count(C.byref(n)))
    return n.value


class HostText:
    """A .gz file inflated by the library (ldx_inflate_gz_file: BGZF blocks in parallel on all host cores).
    `.array` is a uint8 view of the C-owned text; it is released when this object goes away."""

    def __init__(self, path, threads=0):
        self._lib = _lib.load()
        p, n, b = C.c_void_p(), C.c_int64(), C.c_int32()
        check(self._lib.ldx_inflate_gz_file(os.fsencode(path), int(threads), C.byref(p), C.byref(n), C.byref(b)))
        self._p, self.nbytes, self.was_bgzf = p, n.value, bool(b.value)
        self.array = (np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n.value,)) if n.value
                      else np.zeros(0, dtype=np.uint8))

    def close(self):
        if getattr(self, "_p", None) is not None:
            self.array = None
            self._lib.ldx_free_host(self._p)
            self._p = None

    __del__ = close


class Context:
    """One CUDA context/stream of libldx.  Create it AFTER fork (see include/ldx.h)."""

    def __init__(self, device=-1):
        self._lib = _lib.load()
        h = C.c_void_p()
        check(self._lib.ldx_init(int(device), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ldx_destroy(self._h)
            self._h = None

    __del__ = close

    def set_stream(self, cuda_stream):
        """Launch on the given cudaStream_t handle (0 = CUDA's legacy default stream)."""
        check(self._lib.ldx_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def use_own_stream(self):
        check(self._lib.ldx_use_own_stream(self._h))

    def synchronize(self):
        check(self._lib.ldx_synchronize(self._h))

    def set_tuning(self, key, value):
        check(self._lib.ldx_set_tuning(self._h, int(key), int(value)))

    def kernel_timing(self, enable):
        """Read-and-clear the dominant-kernel timer (ms, launches), then switch it on/off."""
        ms, n = C.c_double(), C.c_int64()
        check(self._lib.ldx_kernel_timing(self._h, int(bool(enable)), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    @property
    def sm_count(self):
        n = C.c_int32()
        check(self._lib.ldx_sm_count(self._h, C.byref(n)))
        return n.value

    @property
    def launch_count(self):
        n = C.c_int64()
        check(self._lib.ldx_launch_count(self._h, C.byref(n)))
        return n.value

    def dev_alloc(self, nbytes):
        """Raw device memory (an address) for the outputs of the *_dev calls; release it with dev_free()."""
        p = C.c_void_p()
        check(self._lib.ldx_dev_alloc(self._h, int(nbytes), C.byref(p)))
        return p.value

    def dev_free(self, addr):
        if addr and getattr(self, "_h", None):
            check(self._lib.ldx_dev_free(self._h, C.c_void_p(addr)))

    def resolve(self):
        n = C.c_int64()
        check(self._lib.ldx_resolve(self._h, C.byref(n)))
        return n.value

    def triangle_batch_dev(self, sets, measure="r_square", thres_e4_=None, engine=ENGINE_AUTO):
        """Several variant sets in ONE launch of the all-pairs engine (ldx_triangle_batch_dev).  `sets`: iterable of
        (store, rows, dev_packed[, dev_n11]) with raw device addresses; enqueue only -- call resolve() after."""
        rec = np.zeros(len(sets), dtype=TRIANGLE_SET_DTYPE)
        keep = []                                              # the row arrays must outlive the call
        for k, t in enumerate(sets):
            rows = _i64(t[1])
            keep.append(rows)
            rec[k] = (t[0]._h.value, rows.ctypes.data, rows.shape[0], int(t[2]), int(t[3]) if len(t) > 3 and t[3] else 0)
        check(self._lib.ldx_triangle_batch_dev(self._h, ptr(rec), len(sets), _measure_code(measure), int(thres_e4_ is not None),
                                               int(thres_e4_ or 0), int(engine)))

    def triangle_text(self, packed, v, measure, prefixes, row_begin=0, row_end=None, out=None, dev_text=0):
        """The body lines of ld_triangle's table (ld_triangle.py:356-360), formatted on the GPU.
        `packed`: the words of matrix rows row_begin..row_end-1 -- a uint32 numpy array or a raw device address
        (int).  `prefixes`: one bytes object per matrix row (all v of them; the reference's is rsID + tab +
        position + tab).  -> uint8 array of the text (a view of `out` when given); with `dev_text` = (device
        address, capacity) the text stays in HBM and the byte count is returned."""
        row_end = v if row_end is None else row_end
        assert len(prefixes) == v
        blob = np.frombuffer(b"".join(prefixes) + b"\0", dtype=np.uint8)
        off = np.zeros(v + 1, dtype=np.int64)
        np.cumsum([len(p) for p in prefixes], out=off[1:])
        flags, pk = 0, None
        if isinstance(packed, (int, np.integer)):
            flags, p_packed = _lib.TEXT_PACKED_ON_DEVICE, C.c_void_p(int(packed))
        else:
            pk = np.ascontiguousarray(packed, dtype=np.uint32)
            assert pk.shape[0] == tri_index(row_end, 0) - tri_index(row_begin, 0)
            p_packed = ptr(pk)
        n = C.c_int64()
        args = (self._h, p_packed, int(v), int(row_begin), int(row_end), _measure_code(measure), ptr(blob), ptr(off))
        if dev_text:
            check(self._lib.ldx_triangle_text(*args, flags | _lib.TEXT_OUT_ON_DEVICE, C.c_void_p(dev_text[0]),
                                              int(dev_text[1]), C.byref(n)))
            return n.value
        if out is None:
            out = np.empty(FORMAT_CELL_MAX * v * (row_end - row_begin) + int(off[row_end] - off[row_begin]), dtype=np.uint8)
        check(self._lib.ldx_triangle_text(*args, flags, ptr(out), out.shape[0], C.byref(n)))
        return out[:n.value]

    def finalise_counts(self, n_hap, n11, n1a, n1b):
        """calc_ld.py:33-97 for arrays of counts.  -> dict(d, dprime, r2, packed)."""
        n11 = np.ascontiguousarray(n11, dtype=np.int32)
        n1a = np.ascontiguousarray(np.broadcast_to(np.asarray(n1a, dtype=np.int32), n11.shape))
        n1b = np.ascontiguousarray(np.broadcast_to(np.asarray(n1b, dtype=np.int32), n11.shape))
        n = n11.shape[0]
        out = {"d": np.zeros(n), "dprime": np.zeros(n), "r2": np.zeros(n), "packed": np.zeros(n, np.uint32)}
        rc = self._lib.ldx_finalise_counts(self._h, int(n_hap), ptr(n11), ptr(n1a), ptr(n1b), n, ptr(out["d"]),
                                           ptr(out["dprime"]), ptr(out["r2"]), ptr(out["packed"]))
        if rc == _lib.ERR_EMPTY:
            raise ZeroDivisionError("division by zero")
        check(rc)
        return out

    def calc_ld_lists(self, codes_a, codes_b):
        """Genotype byte codes (0 ref, 1 alt, other = neither) -> LD_RESULT_DTYPE record."""
        a = np.ascontiguousarray(codes_a, dtype=np.uint8)
        b = np.ascontiguousarray(codes_b, dtype=np.uint8)
        out = np.zeros(1, dtype=LD_RESULT_DTYPE)
        rc = self._lib.ldx_calc_ld_lists(self._h, ptr(a), a.shape[0], ptr(b), b.shape[0], ptr(out))
        if rc == _lib.ERR_EMPTY:
            raise ZeroDivisionError("division by zero")   # what calc_ld.py:33 raises on empty input
        check(rc)
        return out[0]


class Store:
    """Bit-plane store of one chromosome in HBM (see include/ldx.h "Data model")."""

    def __init__(self, ctx, n_variants, n_hap, _handle=None):
        self.ctx = ctx
        self._lib = ctx._lib
        if _handle is None:
            _handle = C.c_void_p()
            check(self._lib.ldx_store_create(ctx._h, int(n_variants), int(n_hap), C.byref(_handle)))
        self._h = _handle
        nv, nh, sw = C.c_int64(), C.c_int32(), C.c_int32()
        check(self._lib.ldx_store_shape(self._h, C.byref(nv), C.byref(nh), C.byref(sw)))
        self.n_variants, self.n_hap, self.stride_words = nv.value, nh.value, sw.value

    @classmethod
    def from_planes(cls, ctx, planes, n_hap):
        planes = np.ascontiguousarray(planes, dtype="<u8")
        s = cls(ctx, planes.shape[0], n_hap)
        assert planes.shape[1] == s.stride_words, (planes.shape, s.stride_words)
        s.upload(0, planes)
        return s

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self._lib.ldx_store_destroy(self._h)
        self._h = None

    __del__ = close

    @property
    def planes_ptr(self):
        p = C.c_void_p()
        check(self._lib.ldx_store_planes_ptr(self._h, C.byref(p)))
        return p.value

    # ---- loading
    def pack_gt(self, first_row, text, n_samples, row_off=None, row_pitch=0):
        """GPU bit-packing of VCF GT text (K1).  Returns the per-row status bytes."""
        text = np.ascontiguousarray(text, dtype=np.uint8)
        n_rows = len(row_off) if row_off is not None else (text.shape[0] + 1) // row_pitch if row_pitch else 0
        off = _i64(row_off) if row_off is not None else None
        status = np.zeros(n_rows, dtype=np.uint8)
        check(self._lib.ldx_store_pack_gt(self._h, int(first_row), n_rows, ptr(text), text.shape[0], ptr(off),
                                          int(row_pitch), int(n_samples), ptr(status)))
        return status

    def upload(self, first_row, planes, wait=True):
        """wait=False: only enqueue the copy (ldx_store_upload_async); a pinned `planes` must stay unchanged until the next blocking call."""
        planes = np.ascontiguousarray(planes, dtype="<u8")
        fn = self._lib.ldx_store_upload if wait else self._lib.ldx_store_upload_async
        check(fn(self._h, int(first_row), planes.shape[0], ptr(planes)))

    @classmethod
    def ingest_vcf(cls, ctx, text, n_samples, rows_cap=None):
        """A whole decompressed VCF (bytes-like) -> (store with planes + window annotations, one VCF_ROW_DTYPE
        record per variant), everything parsed and packed on the GPU (ldx_store_ingest_vcf)."""
        buf = np.frombuffer(text, dtype=np.uint8)
        if rows_cap is None:       # a record line holds at least 4 * n_samples - 1 genotype bytes; the rest are header lines
            rows_cap = buf.shape[0] // max(4 * int(n_samples) - 1, 1) + 4096
        while True:
            rows = np.zeros(max(rows_cap, 1), dtype=VCF_ROW_DTYPE)
            h, n = C.c_void_p(), C.c_int64()
            rc = ctx._lib.ldx_store_ingest_vcf(ctx._h, ptr(buf), buf.shape[0], int(n_samples), C.byref(h), ptr(rows), int(rows_cap),
                                               C.byref(n))
            if rc == _lib.ERR_CAPACITY and n.value > rows_cap:       # many short (malformed) lines: now the count is known
                rows_cap = n.value
                continue
            check(rc)
            return cls(ctx, 0, 0, _handle=h), rows[:n.value]

    @classmethod
    def ingest_vcf_file(cls, ctx, path, n_samples, slab_bytes=0, threads=0):
        """<chrom>.vcf.gz -> (store, rows, blob, off, text_bytes) in bounded memory (ldx_store_ingest_vcf_file): the file is inflated,
        parsed and packed `slab_bytes` of text at a time (0: 256 MiB), so neither the host nor the GPU ever holds the whole text.
        rows: one VCF_ROW_DTYPE record per variant; blob/off: the records' fixed columns (as vcf_fixed_columns)."""
        h, n, tb = C.c_void_p(), C.c_int64(), C.c_int64()
        p_rows, p_blob, p_off = C.c_void_p(), C.c_void_p(), C.c_void_p()
        lib = ctx._lib
        check(lib.ldx_store_ingest_vcf_file(ctx._h, os.fsencode(path), int(n_samples), int(slab_bytes), int(threads), C.byref(h),
                                            C.byref(p_rows), C.byref(n), C.byref(p_blob), C.byref(p_off), C.byref(tb)))
        nr = n.value
        off = np.array(_lib_text(lib, p_off, (nr + 1) * 8).view(np.int64))
        rows = np.array(_lib_text(lib, p_rows, nr * VCF_ROW_DTYPE.itemsize).view(VCF_ROW_DTYPE))      # an empty table is released at once
        blob = _lib_text(lib, p_blob, int(off[-1]))
        return cls(ctx, 0, 0, _handle=h), rows, blob, off, tb.value

    @staticmethod
    def vcf_fixed_columns(lib, text, rows):
        """(blob, off): the nine fixed columns of every record back to back, record r = blob[off[r]:off[r+1]]."""
        buf = np.frombuffer(text, dtype=np.uint8)
        rows = np.ascontiguousarray(rows, dtype=VCF_ROW_DTYPE)
        off = np.zeros(rows.shape[0] + 1, dtype=np.int64)
        check(lib.ldx_vcf_copy_prefixes(ptr(buf), buf.shape[0], ptr(rows), rows.shape[0], None, 0, ptr(off)))
        blob = np.zeros(max(int(off[-1]), 1), dtype=np.uint8)
        check(lib.ldx_vcf_copy_prefixes(ptr(buf), buf.shape[0], ptr(rows), rows.shape[0], ptr(blob), blob.shape[0], ptr(off)))
        return blob[:int(off[-1])], off

    def save(self, path):
        """Planes + annotations to one file (ldx_store_save): built once per chromosome, like the reference's cache."""
        check(self._lib.ldx_store_save(self._h, os.fsencode(path)))

    @classmethod
    def load(cls, ctx, path):
        h = C.c_void_p()
        check(ctx._lib.ldx_store_load(ctx._h, os.fsencode(path), C.byref(h)))
        return cls(ctx, 0, 0, _handle=h)

    def download(self, first_row=0, n_rows=None):
        n_rows = self.n_variants - first_row if n_rows is None else n_rows
        out = np.zeros((n_rows, self.stride_words), dtype="<u8")
        check(self._lib.ldx_store_download(self._h, int(first_row), int(n_rows), ptr(out)))
        return out

    # ---- sample selection
    def set_mask(self, mask_words):
        m = np.zeros(self.stride_words, dtype="<u8")
        mask_words = np.asarray(mask_words, dtype="<u8")
        m[:mask_words.shape[0]] = mask_words
        rc = self._lib.ldx_store_set_mask(self._h, ptr(m))
        if rc == _lib.ERR_EMPTY:
            raise ZeroDivisionError("division by zero")
        check(rc)

    def select_haplotypes(self, hap_idx):
        """Mask from haplotype column indices (2*sample_column + allele slot)."""
        bits = np.zeros(self.stride_words * 64, dtype=np.uint8)
        bits[np.asarray(hap_idx, dtype=np.int64)] = 1
        self.set_mask(np.packbits(bits, bitorder="little").view("<u8"))

    def select_all(self):
        self.select_haplotypes(np.arange(self.n_hap))

    def counts(self):
        n1 = np.zeros(self.n_variants, dtype=np.int32)
        p_e4 = np.zeros(self.n_variants, dtype=np.int32)
        n = C.c_int32()
        check(self._lib.ldx_store_counts(self._h, ptr(n1), ptr(p_e4), C.byref(n)))
        return n1, p_e4, n.value

    def row_counts(self):
        """-> (n1, own list length, kind, number of general-route rows) per variant under the current selection."""
        n1 = np.zeros(self.n_variants, dtype=np.int32)
        ln = np.zeros(self.n_variants, dtype=np.int32)
        kind = np.zeros(self.n_variants, dtype=np.int32)
        ng = C.c_int64()
        check(self._lib.ldx_store_row_counts(self._h, ptr(n1), ptr(ln), ptr(kind), C.byref(ng)))
        return n1, ln, kind, ng.value

    def subset(self, hap_idx):
        sel = np.ascontiguousarray(hap_idx, dtype=np.int32)
        h = C.c_void_p()
        rc = self._lib.ldx_store_subset(self._h, ptr(sel), sel.shape[0], C.byref(h))
        check(rc)
        return Store(self.ctx, self.n_variants, sel.shape[0], _handle=h)

    def set_annotations(self, pos0, end0, idnum, eligible):
        pos0 = np.ascontiguousarray(pos0, dtype=np.int32)
        end0 = np.ascontiguousarray(end0, dtype=np.int32)
        idnum = _i64(idnum)
        eligible = np.ascontiguousarray(eligible, dtype=np.uint8)
        assert pos0.shape[0] == end0.shape[0] == idnum.shape[0] == eligible.shape[0] == self.n_variants
        check(self._lib.ldx_store_set_annotations(self._h, ptr(pos0), ptr(end0), ptr(idnum), ptr(eligible)))

    # ---- compute
    def pairs(self, ia, ib, raw=True):
        """LD of explicit row pairs (var_1 = ia[k], var_2 = ib[k]).  -> dict of arrays."""
        ia, ib = _i64(ia), _i64(ib)
        n = ia.shape[0]
        out = {"n11": np.zeros(n, np.int32), "packed": np.zeros(n, np.uint32)}
        if raw:
            out.update(d=np.zeros(n), dprime=np.zeros(n), r2=np.zeros(n))
        check(self._lib.ldx_pairs(self._h, ptr(ia), ptr(ib), n, ptr(out["n11"]), ptr(out.get("d")),
                                  ptr(out.get("dprime")), ptr(out.get("r2")), ptr(out["packed"])))
        return out

    def window(self, q_row, lo, hi, win_start, win_end, measure, thres_e4_, cap=None):
        """Fused ld_area scan.  -> (hits sorted by (query, row), pairs scanned)."""
        q_row, lo, hi = _i64(q_row), _i64(lo), _i64(hi)
        ws = np.ascontiguousarray(win_start, dtype=np.int32)
        we = np.ascontiguousarray(win_end, dtype=np.int32)
        nq = q_row.shape[0]
        if cap is None:
            cap = int(min(max(int((hi - lo).sum()), 1), 1 << 22))
        while True:
            hits = np.zeros(cap, dtype=HIT_DTYPE)
            n_hits, n_scanned = C.c_int64(), C.c_int64()
            rc = self._lib.ldx_window(self._h, ptr(q_row), ptr(lo), ptr(hi), ptr(ws), ptr(we), nq,
                                      _measure_code(measure), int(thres_e4_), ptr(hits), cap,
                                      C.byref(n_hits), C.byref(n_scanned))
            if rc == _lib.ERR_CAPACITY:
                cap = int(n_hits.value) + 1024
                continue
            check(rc)
            return hits[:n_hits.value], n_scanned.value

    def triangle(self, rows, measure="r_square", thres_e4_=None, engine=ENGINE_AUTO, want_n11=False, out=None):
        """All pairs of `rows` (matrix order).  -> (packed lower triangle by rows, n11 or None).
        `out` may be a caller-owned uint32 array (e.g. pinned memory) of v*(v-1)/2 entries."""
        rows = _i64(rows)
        v = rows.shape[0]
        n_pairs = v * (v - 1) // 2
        packed = np.zeros(n_pairs, dtype=np.uint32) if out is None else out
        assert packed.dtype == np.uint32 and packed.shape[0] == n_pairs and packed.flags.c_contiguous
        n11 = np.zeros(n_pairs, dtype=np.int32) if want_n11 else None
        check(self._lib.ldx_triangle(self._h, ptr(rows), v, _measure_code(measure), int(thres_e4_ is not None),
                                     int(thres_e4_ or 0), int(engine), ptr(packed), ptr(n11)))
        return packed, n11

    def triangle_rows(self, rows, row_begin, row_end, measure="r_square", thres_e4_=None, engine=ENGINE_AUTO,
                      want_n11=False, out=None):
        """Matrix rows row_begin..row_end-1 of the triangle (row_begin % 128 == 0): the contiguous slice
        [tri(row_begin), tri(row_end)) of the packed lower triangle.  The shard unit of the multi-GPU path."""
        rows = _i64(rows)
        n_pairs = tri_index(row_end, 0) - tri_index(row_begin, 0)
        packed = np.zeros(n_pairs, dtype=np.uint32) if out is None else out
        assert packed.dtype == np.uint32 and packed.shape[0] == n_pairs and packed.flags.c_contiguous
        n11 = np.zeros(n_pairs, dtype=np.int32) if want_n11 else None
        check(self._lib.ldx_triangle_rows(self._h, ptr(rows), rows.shape[0], int(row_begin), int(row_end),
                                          _measure_code(measure), int(thres_e4_ is not None), int(thres_e4_ or 0),
                                          int(engine), ptr(packed), ptr(n11)))
        return packed, n11

    def triangle_values(self, rows, measure="r_square", thres_e4_=None, engine=ENGINE_AUTO, row_begin=0, row_end=None, out=None):
        """The same triangle as 2 bytes per pair for ONE measure (ldx_triangle_values): value * 10^4 in bits 0..13, V16_BELOW, V16_INT0."""
        rows = _i64(rows)
        row_end = rows.shape[0] if row_end is None else row_end
        n_pairs = tri_index(row_end, 0) - tri_index(row_begin, 0)
        out = np.zeros(n_pairs, dtype=np.uint16) if out is None else out
        assert out.dtype == np.uint16 and out.shape[0] == n_pairs and out.flags.c_contiguous
        check(self._lib.ldx_triangle_values(self._h, ptr(rows), rows.shape[0], int(row_begin), int(row_end), _measure_code(measure),
                                            int(thres_e4_ is not None), int(thres_e4_ or 0), int(engine), ptr(out)))
        return out

    def triangle_hits(self, rows, measure, thres_e4_, engine=ENGINE_AUTO, cap=None, out=None):
        """Only the pairs whose rounded measure passes the threshold (-z), in matrix order (ldx_triangle_hits) -> PAIR_HIT_DTYPE array."""
        rows = _i64(rows)
        cap = (1 << 16) if cap is None and out is None else (out.shape[0] if out is not None else cap)
        while True:
            hits = np.zeros(max(cap, 1), dtype=PAIR_HIT_DTYPE) if out is None else out
            n = C.c_int64()
            rc = self._lib.ldx_triangle_hits(self._h, ptr(rows), rows.shape[0], _measure_code(measure), int(thres_e4_), int(engine),
                                             ptr(hits), int(cap), C.byref(n))
            if rc == _lib.ERR_CAPACITY and out is None:
                cap = int(n.value) + 1024
                continue
            check(rc)
            return hits[:n.value]

    def triangle_table(self, rows, prefixes, measure="r_square", thres_e4_=None, engine=ENGINE_AUTO, row_begin=0,
                       row_end=None, out=None):
        """Matrix rows row_begin..row_end-1 (row_begin % 128 == 0) of ld_triangle's table as text: all-pairs kernel,
        settlement and the writer (ld_triangle.py:133-230, :356-360) in one library call; the words never leave HBM.
        `prefixes`: one bytes object per matrix row.  -> uint8 array (a view of `out` when given)."""
        rows = _i64(rows)
        v = rows.shape[0]
        row_end = v if row_end is None else row_end
        assert len(prefixes) == v
        blob = np.frombuffer(b"".join(prefixes) + b"\0", dtype=np.uint8)
        off = np.zeros(v + 1, dtype=np.int64)
        np.cumsum([len(p) for p in prefixes], out=off[1:])
        if out is None:
            out = np.empty(FORMAT_CELL_MAX * v * (row_end - row_begin) + int(off[row_end] - off[row_begin]), dtype=np.uint8)
        n = C.c_int64()
        check(self._lib.ldx_triangle_table(self._h, ptr(rows), v, int(row_begin), int(row_end), _measure_code(measure),
                                           int(thres_e4_ is not None), int(thres_e4_ or 0), int(engine), ptr(blob), ptr(off),
                                           0, ptr(out), out.shape[0], C.byref(n)))
        return out[:n.value]

    def triangle_rows_dev(self, rows, row_begin, row_end, dev_packed, dev_n11=0, measure="r_square", thres_e4_=None,
                          engine=ENGINE_AUTO):
        rows = _i64(rows)
        check(self._lib.ldx_triangle_rows_dev(self._h, ptr(rows), rows.shape[0], int(row_begin), int(row_end),
                                              _measure_code(measure), int(thres_e4_ is not None), int(thres_e4_ or 0),
                                              int(engine), C.c_void_p(dev_packed), C.c_void_p(dev_n11 or 0)))

    def triangle_dev(self, rows, dev_packed, dev_n11=0, measure="r_square", thres_e4_=None, engine=ENGINE_AUTO):
        """Device-resident output (raw device addresses); enqueue only.  Call ctx.resolve() after."""
        rows = _i64(rows)
        check(self._lib.ldx_triangle_dev(self._h, ptr(rows), rows.shape[0], _measure_code(measure),
                                         int(thres_e4_ is not None), int(thres_e4_ or 0), int(engine),
                                         C.c_void_p(dev_packed), C.c_void_p(dev_n11 or 0)))

    def window_dev(self, q_row, lo, hi, win_start, win_end, measure, thres_e4_, dev_hits, cap, dev_counters):
        q_row, lo, hi = _i64(q_row), _i64(lo), _i64(hi)
        ws = np.ascontiguousarray(win_start, dtype=np.int32)
        we = np.ascontiguousarray(win_end, dtype=np.int32)
        check(self._lib.ldx_window_dev(self._h, ptr(q_row), ptr(lo), ptr(hi), ptr(ws), ptr(we), q_row.shape[0],
                                       _measure_code(measure), int(thres_e4_), C.c_void_p(dev_hits), int(cap),
                                       C.c_void_p(dev_counters)))


def tri_index(row, col):
    """Position of pair (row > col) in the packed lower triangle."""
    return row * (row - 1) // 2 + col if row > 0 else col
