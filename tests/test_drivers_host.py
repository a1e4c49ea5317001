"""Host-side logic of the driver adapters that needs no GPU: the conversion.db helpers (restated from the reference, and the
reference's own modules preferred when its tree is importable), window bookkeeping."""
import os
import sqlite3
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def _make_db(path):
    with sqlite3.connect(path) as conn:
        cur = conn.cursor()
        cur.execute("CREATE TABLE samples (sample TEXT, pop TEXT, super_pop TEXT, gender TEXT)")
        cur.executemany("INSERT INTO samples VALUES (?, ?, ?, ?)",
                        [("S1", "GBR", "EUR", "male"), ("S2", "GBR", "EUR", "female"), ("S3", "YRI", "AFR", "male"), ("S4", "CHB", "EAS", "female")])
        cur.execute("CREATE TABLE variants (CHROM TEXT, POS INTEGER, ID TEXT)")
        cur.executemany("INSERT INTO variants VALUES (?, ?, ?)", [("22", 100, "rs1"), ("22", 200, "rs2"), ("7", 50, "rs3")])
        conn.commit()


def test_conversion_db_helpers_follow_the_reference_semantics(tmp_path):
    """backend/get_sample_names.py:5-45 and backend/create_src_dict.py:5-64 restated: gender and (super-)population filters with
    the one-element-tuple quirk, leftmost rsID per line, meta lines skipped, unknown IDs dropped, grouping by chromosome."""
    from ld_tools_b200 import drivers
    db = str(tmp_path / "conversion.db")
    _make_db(db)
    assert drivers.get_sample_names(("male", "female"), ("ALL",), db) == ["S1", "S2", "S3", "S4"]
    assert drivers.get_sample_names(("male",), ("EUR", "YRI"), db) == ["S1", "S3"]
    assert drivers.get_sample_names(("female",), ("GBR",), db) == ["S2"]
    src = tmp_path / "src"
    src.mkdir()
    (src / "t.tsv").write_text("#meta rs2\nhead\tx\nchr22\trs1 and rs2 later\njunk\nfoo rs3\trs1\nrs999\nrs1 again\n")
    d = drivers.create_src_dict(str(src), "t.tsv", 2, db)
    assert sorted(d) == ["22", "7"] and d["22"] == [[100, "rs1"]] and d["7"] == [[50, "rs3"]]
    assert drivers.create_src_dict(str(src), "t.tsv", 7, db) == {}


def test_reference_helpers_are_preferred_when_the_reference_tree_is_importable(tmp_path, monkeypatch):
    from ld_tools_b200 import drivers
    monkeypatch.delenv("LD_TOOLS_REFERENCE", raising=False)
    for m in [k for k in sys.modules if k == "backend" or k.startswith("backend.")]:
        monkeypatch.delitem(sys.modules, m)
    own = drivers._reference_helpers()
    if own[0].__module__ != "backend.get_sample_names":          # (a reference tree already on sys.path is fine too)
        assert own == (drivers.get_sample_names, drivers.create_src_dict)
    ref = tmp_path / "ref"
    (ref / "backend").mkdir(parents=True)
    (ref / "backend" / "calc_ld.py").write_text("def calc_ld(a, b):\n    return {}\n")
    (ref / "backend" / "get_sample_names.py").write_text("def get_sample_names(g, p, db):\n    return ['from-the-reference']\n")
    (ref / "backend" / "create_src_dict.py").write_text("def create_src_dict(d, f, m, db):\n    return {'ref': []}\n")
    monkeypatch.setenv("LD_TOOLS_REFERENCE", str(ref))
    for m in [k for k in sys.modules if k == "backend" or k.startswith("backend.")]:
        monkeypatch.delitem(sys.modules, m)
    g, c = drivers._reference_helpers()
    assert g(None, None, None) == ["from-the-reference"] and c(None, None, None, None) == {"ref": []}
    assert str(ref) not in sys.path
    # somebody else's `backend` package (no calc_ld.py next to it) is not taken for the reference
    other = tmp_path / "other"
    (other / "backend").mkdir(parents=True)
    (other / "backend" / "get_sample_names.py").write_text("def get_sample_names(g, p, db):\n    return ['impostor']\n")
    (other / "backend" / "create_src_dict.py").write_text("def create_src_dict(d, f, m, db):\n    return {}\n")
    monkeypatch.delenv("LD_TOOLS_REFERENCE")
    for m in [k for k in sys.modules if k == "backend" or k.startswith("backend.")]:
        monkeypatch.delitem(sys.modules, m)
    monkeypatch.syspath_prepend(str(other))
    assert drivers._reference_helpers() == (drivers.get_sample_names, drivers.create_src_dict)
    for m in [k for k in sys.modules if k == "backend" or k.startswith("backend.")]:
        monkeypatch.delitem(sys.modules, m)
