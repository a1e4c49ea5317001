"""The oracle is only trusted after it reproduces the reference's own outputs.

tests/golden/calc_ld_golden.json holds outputs of the unmodified reference calc_ld
(/root/reference/backend/calc_ld.py) -- rounded dict, pre-rounding values and D.  Every vector
is replayed through (1) the pure-Python port and (2) the C restatement.  Bit-exact, no
tolerance: both run the same IEEE operations and the same libm pow as the reference.
"""
import os

import numpy as np
import pytest

from conftest import ROOT, decode_genotypes, decode_raw
from oracle import calc_ld_port, ld_oracle

KEYS = ("r_square", "d_prime", "var_1_alt_freq", "var_2_alt_freq")
RAW_FIELD = {"r_square": "r2", "d_prime": "dprime", "var_1_alt_freq": "p_a", "var_2_alt_freq": "p_b"}


def same_obj(a, b):
    """Equal value AND equal Python type (int 0 vs float 0.0 prints differently downstream)."""
    return type(a) is type(b) and repr(a) == repr(b)


def lists_from_counts(n, n11, a, b):
    oa, ob = a - n11, b - n11
    rest = n - n11 - oa - ob
    return ([1] * n11 + [1] * oa + [0] * ob + [0] * rest,
            [1] * n11 + [0] * oa + [1] * ob + [0] * rest)


def check_port(g1, g2, out):
    got = calc_ld_port.calc_ld(g1, g2)
    assert list(got) == list(KEYS)
    for k, want in zip(KEYS, out[:4]):
        assert repr(got[k]) == want, (k, got[k], want)
    full = calc_ld_port.calc_ld_full(g1, g2)
    for k, want in zip(("r_square", "d_prime", "p_a", "p_b", "d"), out[4:9]):
        assert same_obj(full[k], decode_raw(want)), (k, full[k], want)


def check_c(res, out):
    ref = ld_oracle.as_reference_dict(res)
    for k, want in zip(KEYS, out[:4]):
        assert repr(ref[k]) == want, (k, ref[k], want)
    for k, want in zip(KEYS, out[4:8]):
        want = decode_raw(want)
        if isinstance(want, int):
            flag = {"r_square": "r2_is_int0", "d_prime": "dprime_is_int0"}[k]
            assert res[flag] == 1
        else:
            if k in ("r_square", "d_prime"):
                assert res[{"r_square": "r2_is_int0", "d_prime": "dprime_is_int0"}[k]] == 0
            assert float(res[RAW_FIELD[k]]).hex() == want.hex(), (k, res[RAW_FIELD[k]], want)
    assert float(res["d"]).hex() == float(decode_raw(out[8])).hex()


def test_list_cases_port_and_c(golden):
    for name, sa, sb, *out in golden["list_cases"]:
        g1, g2 = decode_genotypes(sa), decode_genotypes(sb)
        if name == "tuple_inputs":
            g1, g2 = tuple(g1), tuple(g2)
        check_port(g1, g2, out)
        res = ld_oracle.calc_ld_bytes(ld_oracle.encode_genotypes(g1), ld_oracle.encode_genotypes(g2))
        check_c(res, out)


def test_count_cases_c(golden):
    for tag, n, n11, a, b, *out in golden["count_cases"]:
        check_c(ld_oracle.finalise(n, n11, a, n - a, b, n - b), out)


def test_count_cases_port_sample(golden):
    cases = golden["count_cases"]
    for tag, n, n11, a, b, *out in cases[::7]:
        check_port(*lists_from_counts(n, n11, a, b), out)


def test_empty_input_raises_like_reference():
    with pytest.raises(ZeroDivisionError):
        calc_ld_port.calc_ld([], [])
    with pytest.raises(ZeroDivisionError):
        ld_oracle.calc_ld_bytes(np.zeros(0, np.uint8), np.zeros(0, np.uint8))


def test_round4_matches_python_round():
    rng = np.random.default_rng(7)
    xs = list(rng.random(20000)) + [k / 64 for k in range(65)] + [(k + 0.5) / 1e4 for k in range(0, 10001, 7)]
    xs += [np.nextafter((k + 0.5) / 1e4, 2.0) for k in range(0, 10001, 13)]
    xs += [np.nextafter((k + 0.5) / 1e4, -1.0) for k in range(0, 10001, 13)]
    xs += [1.0000000000001792, 1.0000000000003582, 0.79995, 0.03125, 0.09375]
    for x in xs:
        assert ld_oracle.round4(float(x)) == round(float(x), 4), x


def test_bitplane_routes_agree():
    """popcount(mask & a & b) == dense 0/1 matmul == the list-level count, incl. pad bits."""
    rng = np.random.default_rng(11)
    for n_hap in (2, 63, 64, 65, 198, 1006, 5008):
        h = (rng.random((12, n_hap)) < rng.random((12, 1))).astype(np.uint8)
        planes = ld_oracle.pack_bits(h)
        assert planes.shape[1] % 16 == 0
        assert (ld_oracle.unpack_bits(planes, n_hap) == h).all()
        sel = np.flatnonzero(rng.random(n_hap) < 0.4)
        if sel.size == 0:
            sel = np.array([0])
        mask = ld_oracle.mask_from_haplotypes(sel, n_hap)
        gram = ld_oracle.n11_matrix(planes, mask, n_hap)
        ia, ib = np.triu_indices(12, 1)
        res = ld_oracle.pairs(planes, mask, n_hap, ia, ib)
        assert (res["n_11"] == gram[ia, ib]).all()
        assert (res["n_hap"] == sel.size).all()
        for k in (0, 5, len(ia) - 1):
            full = calc_ld_port.calc_ld_full(list(map(int, h[ia[k], sel])), list(map(int, h[ib[k], sel])))
            assert full["n_11"] == res["n_11"][k] and full["n_a1"] == res["n_a1"][k]
            assert same_obj(ld_oracle.as_reference_dict(res[k])["r_square"],
                            calc_ld_port.calc_ld(list(map(int, h[ia[k], sel])),
                                                 list(map(int, h[ib[k], sel])))["r_square"])


def test_triangle_orientation_row_is_var_1():
    """ld_triangle.py:193 calls calc_ld(y = row variant, x = column variant) with row > col."""
    rng = np.random.default_rng(3)
    h = (rng.random((9, 130)) < 0.3).astype(np.uint8)
    planes = ld_oracle.pack_bits(h)
    mask = ld_oracle.mask_from_haplotypes(np.arange(130), 130)
    rows = np.array([4, 0, 7, 2, 8])
    tri = ld_oracle.triangle(planes, mask, 130, rows)
    for r in range(1, 5):
        for c in range(r):
            want = calc_ld_port.calc_ld(list(map(int, h[rows[r]])), list(map(int, h[rows[c]])))
            got = ld_oracle.as_reference_dict(tri[r * (r - 1) // 2 + c])
            assert all(same_obj(got[k], want[k]) for k in KEYS)


def test_pack_gt_text():
    rng = np.random.default_rng(5)
    n_samples = 37
    gt = (rng.random((6, n_samples, 2)) < 0.3).astype(np.uint8)
    rows, offs, text = [], [], b""
    for v in range(6):
        prefix = f"22\t{100 + v}\trs{v}\tA\tG\t100\tPASS\tVT=SNP\tGT\t".encode()
        body = "\t".join(f"{a}|{b}" for a, b in gt[v]).encode() + b"\n"
        offs.append(len(text) + len(prefix))
        text += prefix + body
    buf = np.frombuffer(text, dtype=np.uint8).copy()
    planes, status = ld_oracle.pack_gt(buf, np.array(offs), n_samples)
    assert (status == 0).all()
    assert (ld_oracle.unpack_bits(planes, 2 * n_samples) == gt.reshape(6, -1)).all()
    buf[offs[2] + 4 * 3] = ord(".")          # missing allele
    buf[offs[4] + 4 * 5 + 1] = ord("/")      # unphased
    _, status = ld_oracle.pack_gt(buf, np.array(offs), n_samples)
    assert status.tolist() == [0, 0, 1, 0, 1, 0]


def test_table_port_reproduces_reference_tables():
    """The table writer port (oracle/table_port.py, ld_triangle.py:114,:150,:223-230,:356-360) against the .tsv
    files the unmodified reference driver wrote: cells parsed back to the objects they print, lines re-created."""
    import glob
    from oracle import table_port
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "drivers", "triangle_*", "*", "*.tsv")))
    assert len(files) >= 2
    for path in files:
        with open(path) as fh:
            lines = fh.read().split("\n")
        assert lines[0].startswith("##General\tinfo:") and lines[1] == "" and lines[-1] == ""
        ids, poss = lines[2].split("\t")[2:], lines[3].split("\t")[2:]
        body = lines[4:-1]
        v = len(ids)
        assert len(body) == v and len(poss) == v
        cells = [ln.split("\t")[2:] for ln in body]
        assert all(len(row) == v for row in cells)
        assert all(cells[r][c] == "0" for r in range(v) for c in range(r, v))       # :150: only row > col is filled
        objs = [[0 if t == "0" else float(t) for t in row] for row in cells]
        assert any(isinstance(x, float) for row in objs for x in row)
        text = table_port.matrix_body(lambda r, c: objs[r][c], v, ids, poss)
        assert text == "\n".join(body) + "\n"
        # every float cell is a round(x, 4) value printed by str(): at most four decimals, no exponent
        for row in cells:
            for t in row:
                assert t == "0" or (t == str(round(float(t), 4)) and "e" not in t)


@pytest.mark.parametrize("n_hap,measure,thres", [(14, "d_prime", None), (198, "r_square", 0.3), (1006, "d_prime", 0.9)])
def test_packed_words_print_like_the_reference_objects(n_hap, measure, thres):
    """Host-side decode of the result word (engine.measure_value + BELOW_THRES) feeds str() the same objects the
    reference's matrix holds: table text from the words = table text from the oracle's calc_ld dicts."""
    from oracle import table_port
    from ld_tools_b200.engine import BELOW_THRES, measure_value, threshold_e4
    from ld_tools_b200.synth import synth_haplotypes
    v = 40
    planes = ld_oracle.pack_bits(synth_haplotypes(v, n_hap, seed=n_hap))
    mask = ld_oracle.mask_from_haplotypes(np.arange(n_hap), n_hap)
    res = ld_oracle.triangle(planes, mask, n_hap, np.arange(v))
    vals = [ld_oracle.as_reference_dict(x)[measure] for x in res]
    ids, poss = [f"rs{k}" for k in range(v)], [str(100 + k) for k in range(v)]
    want = table_port.matrix_body(lambda r, c: vals[r * (r - 1) // 2 + c], v, ids, poss, thres)
    words = ld_oracle.packed_of(res)
    if thres is not None:
        shift = 0 if measure == "r_square" else 16
        words = words | np.where(((words >> shift) & 0x3FFF) < threshold_e4(thres), np.uint32(BELOW_THRES), np.uint32(0))
    lines = []
    for r in range(v):
        cells = ["0" if (c >= r or words[r * (r - 1) // 2 + c] & BELOW_THRES) else str(measure_value(words[r * (r - 1) // 2 + c], measure))
                 for c in range(v)]
        lines.append(ids[r] + "\t" + poss[r] + "\t" + "\t".join(cells) + "\n")
    assert "".join(lines) == want
