"""Drop-in replacement for ld-tools' `backend/calc_ld.py::calc_ld` (calc_ld.py:3-99).

Same signature, same dict (keys r_square, d_prime, var_1_alt_freq, var_2_alt_freq -- they are
API, indexed by the CLI's -l choice at ld_area.py:248 / ld_triangle.py:224), same value types
(int 0 vs float), same ZeroDivisionError on empty input -- but the counting and the D/D'/r2
arithmetic run in libldx.so on the GPU (ldx_calc_ld_lists).  No CPU fallback.

This scalar entry point exists for API parity; the batched entry points (Store.pairs / window /
triangle) are what the re-pointed drivers use.
"""
import os

import numpy as np

from .engine import Context

__all__ = ["calc_ld"]

_ctx_by_pid = {}
_CODE_TABLE = bytes([0, 1] + [255] * 254)          # byte value -> genotype code: 0 ref, 1 alt, anything else neither


def _context():
    """One lazily created context per process: never carried across fork()
    (the reference fans out with multiprocessing.Pool, ld_area.py:336)."""
    pid = os.getpid()
    ctx = _ctx_by_pid.get(pid)
    if ctx is None:
        ctx = _ctx_by_pid[pid] = Context()
    return ctx


def encode_genotypes(genotypes):
    """Sequence of numeric genotypes -> byte codes: 1 where x == 1, 0 where x == 0, 255 otherwise.

    Mirrors list.count(1) / list.count(0) (calc_ld.py:37-40): equality, not identity, so 1.0 and
    True count as 1; None, 2, '.' count as neither."""
    if isinstance(genotypes, (list, tuple)):
        # fast path for the drivers' lists of small ints (and bools): bytes() refuses anything else -- floats, None, '.',
        # negative or > 255 -- and those take the general path below
        try:
            return np.frombuffer(bytes(genotypes).translate(_CODE_TABLE), dtype=np.uint8)
        except (TypeError, ValueError):
            pass
    arr = np.asarray(genotypes)
    if arr.dtype == object or arr.dtype.kind not in "biuf":
        arr = np.asarray(genotypes, dtype=object)
        is1 = np.fromiter((x == 1 for x in arr), dtype=bool, count=arr.shape[0])
        is0 = np.fromiter((x == 0 for x in arr), dtype=bool, count=arr.shape[0])
    else:
        is1, is0 = arr == 1, arr == 0
    codes = np.full(arr.shape[0], 255, dtype=np.uint8)
    codes[is1] = 1
    codes[is0] = 0
    return codes


def calc_ld(var_1_genotypes, var_2_genotypes):
    res = _context().calc_ld_lists(encode_genotypes(var_1_genotypes), encode_genotypes(var_2_genotypes))
    return {'r_square': 0 if res["r2_is_int0"] else float(res["r2_e4"]) / 10000.0,
            'd_prime': 0 if res["dprime_is_int0"] else float(res["dprime_e4"]) / 10000.0,
            'var_1_alt_freq': float(res["p_a_e4"]) / 10000.0,
            'var_2_alt_freq': float(res["p_b_e4"]) / 10000.0}
