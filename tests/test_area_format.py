"""ldx_area_format (host code in libldx: no GPU needed) against the reference's own per-hit writers restated with Python's
str() / json.dump (ld_area.py:252-283): TSV lines, JSON list elements, rsID lines -- int 0 vs float, negative distances,
multi-valued ALT / VT, text that needs JSON escapes, queries without hits."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_table(rng, n):
    from ld_tools_b200._lib import VCF_ROW_DTYPE
    rows = np.zeros(n, dtype=VCF_ROW_DTYPE)
    recs, blob, off = [], bytearray(), [0]
    pos = np.cumsum(rng.integers(1, 2000, size=n)) + 16_000_000
    for k in range(n):
        rid = f"rs{int(rng.integers(1, 10**9))}" if k % 7 else ["esv12", ".", 'weird"id\\x', "rsé中"][k % 4]
        ref = "ACGT"[k % 4] * int(rng.integers(1, 6))
        alt = ",".join(rng.choice(["A", "C", "G", "T", "<CN0>", "AT"], size=int(rng.integers(1, 3))))
        vt = ["SNP", "INDEL", "SNP,INDEL", "SV"][k % 4]
        info = [f"AC={k}", f"VT={vt}", "AN=5008"]
        if k % 5 == 0:
            info = [f"VT={vt}", "MULTI_ALLELIC"]
        if k % 11 == 0:
            info = ["AC=1", "NOVT=1"]                     # no VT key: an empty type column here (the reference would raise)
            vt = ""
        fixed = f"22\t{pos[k]}\t{rid}\t{ref}\t{alt}\t100\tPASS\t{';'.join(info)}\tGT\t".encode()
        f = fixed.split(b"\t")
        o = np.cumsum([0] + [len(x) + 1 for x in f[:-1]])
        rows[k]["pos"], rows[k]["ref_len"] = pos[k], len(ref)
        rows[k]["id_off"], rows[k]["ref_off"], rows[k]["alt_off"], rows[k]["info_off"], rows[k]["fmt_off"], rows[k]["gt_off"] = o[2], o[3], o[4], o[7], o[8], o[9]
        blob += fixed
        off.append(len(blob))
        recs.append({"pos": int(pos[k]), "id": rid, "ref": ref, "alt": alt, "vt": vt})
    return rows, np.frombuffer(bytes(blob), dtype=np.uint8), np.array(off, dtype=np.int64), recs


def value(word, r2):
    from ld_tools_b200.engine import dprime_value, r2_value
    return r2_value(word) if r2 else dprime_value(word)


@pytest.mark.parametrize("threads", [1, 0])
def test_area_format_equals_python_writers(threads):
    from ld_tools_b200 import _lib
    from ld_tools_b200._lib import AREA_JSON, AREA_RSIDS, AREA_TSV, DP_INT0, HIT_DTYPE, R2_INT0
    from ld_tools_b200.engine import area_format
    lib = _lib.load()
    rng = np.random.default_rng(3)
    n, nq = 400, 60
    rows, blob, off, recs = make_table(rng, n)
    p_e4 = rng.integers(0, 10001, size=n).astype(np.int32)
    q_row = rng.integers(0, n, size=nq)
    hits = []
    for k in range(nq):
        if k % 9 == 4:
            continue                                      # a query without hits: an empty slice
        for r in np.sort(rng.choice(n, size=int(rng.integers(1, 200 if k == 7 else 12)), replace=False)):
            w = int(rng.integers(0, 10001)) | (int(rng.integers(0, 10001)) << 16)
            u = rng.random()
            if u < 0.1:
                w = R2_INT0 | DP_INT0
            elif u < 0.2:
                w = (w & 0xFFFF0000) | R2_INT0            # D' = 0.0 exactly, r2 the int 0 (calc_ld.py:89-90)
            elif u < 0.3:
                w = [0, 10000, 10000 << 16, 5000 | (1 << 16)][int(rng.integers(4))]
            hits.append((k, int(r), int(rng.integers(0, 5000)), w))
    hits = np.array(hits, dtype=HIT_DTYPE)
    header = ["hg38_pos", "rsID", "ref", "alt", "type", "alt_freq", "r2", "D'", "dist"]
    for fmt in (AREA_TSV, AREA_JSON, AREA_RSIDS):
        text, qoff = area_format(lib, hits, q_row, blob, off, rows, p_e4, fmt, threads=threads)
        assert qoff[0] == 0 and qoff[-1] == text.shape[0]
        for k in range(nq):
            mine = hits[hits["query"] == k]
            got = text[qoff[k]:qoff[k + 1]].tobytes().decode()
            anns = []
            for h in mine:
                r = recs[h["row"]]
                anns.append([r["pos"], r["id"], r["ref"], r["alt"], r["vt"], p_e4[h["row"]] / 10000.0, value(h["packed"], True),
                             value(h["packed"], False), r["pos"] - recs[q_row[k]]["pos"]])                 # ld_area.py:264-272
            if fmt == AREA_TSV:
                want = "".join("\t".join(map(str, a)) + "\n" for a in anns)                               # :273-274
            elif fmt == AREA_RSIDS:
                want = "".join(a[1] + "\n" for a in anns)                                                 # :258-260
            else:
                head = [{"chr": "22"}, dict(zip(header, ["q"] * 9))]
                full = json.dumps(head + [dict(zip(header, a)) for a in anns], indent=4)                  # :275-283
                base = json.dumps(head, indent=4)
                assert full.startswith(base[:-2]) and full.endswith("\n]")
                want = full[len(base) - 2:-2]
            assert got == want, (fmt, k)
    # per-hit overrides {alt_freq, r2, D'} (general route): any magnitude, < 0 = default
    ov = np.full((len(hits), 3), -1, dtype=np.int32)
    ov[:, 0] = rng.integers(0, 30001, size=len(hits))
    big = rng.random(len(hits)) < 0.3
    ov[big, 1] = rng.choice([16383, 16384, 20001, 89099, 100000, 1234500, 2147480000 // 10000 * 10000 + 7], size=int(big.sum()))
    ov[big, 2] = rng.integers(0, 5_000_000, size=int(big.sum()))
    for fmt in (AREA_TSV, AREA_JSON):
        text, qoff = area_format(lib, hits, q_row, blob, off, rows, p_e4, fmt, overrides=ov, threads=threads)
        if fmt == AREA_TSV:
            cols = [ln.split("\t") for ln in text.tobytes().decode().split("\n")[:-1]]
        else:
            cols = [[None] * 5 + [c.split(": ")[1].rstrip(",") for c in el.split("\n")[6:9]] for el in text.tobytes().decode().split("\n    {")[1:]]
        for h, o, c in zip(hits, ov, cols):
            want = [str(o[0] / 10000.0), str(o[1] / 10000.0) if o[1] >= 0 else str(value(h["packed"], True)),
                    str(o[2] / 10000.0) if o[2] >= 0 else str(value(h["packed"], False))]
            assert c[5:8] == want, (fmt, c[5:8], want)


def test_area_format_rejects_bad_input():
    from ld_tools_b200 import LdxError, _lib
    from ld_tools_b200._lib import AREA_TSV, HIT_DTYPE
    from ld_tools_b200.engine import area_format
    lib = _lib.load()
    rows, blob, off, _ = make_table(np.random.default_rng(1), 10)
    p_e4 = np.zeros(10, np.int32)
    for bad in ([(1, 3, 0, 0), (0, 2, 0, 0)], [(0, 10, 0, 0)], [(2, 1, 0, 0)]):          # unsorted, row out of range, query out of range
        with pytest.raises(LdxError):
            area_format(lib, np.array(bad, dtype=HIT_DTYPE), np.array([1, 2]), blob, off, rows, p_e4, AREA_TSV)
    text, qoff = area_format(lib, np.zeros(0, dtype=HIT_DTYPE), np.array([1, 2]), blob, off, rows, p_e4, AREA_TSV)
    assert text.shape[0] == 0 and qoff.tolist() == [0, 0, 0]
