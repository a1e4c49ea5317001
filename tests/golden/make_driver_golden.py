#!/usr/bin/env python3
"""Generate tests/golden/drivers/: the output trees of the UNMODIFIED reference drivers
(/root/reference/ld_area.py, ld_triangle.py, ld_lite.py) on the synthetic data set of
tests/driver_cases.py.  pysam and plotly are absent from this container, so the drivers run against
the pure-Python stand-ins in tests/refshim/ (see their docstrings).  Run here (needs /root/reference):

    python tests/golden/make_driver_golden.py
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import driver_cases as dc  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(HERE, "drivers")


def run_ref(script, argv, cwd):
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tests", "refshim"), PYTHONDONTWRITEBYTECODE="1", LANG="en_US.UTF-8",
               LC_ALL="en_US.UTF-8", PYTHONWARNINGS="ignore")
    r = subprocess.run([sys.executable, os.path.join(REF, script)] + argv, cwd=cwd, env=env, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(f"{script} {argv}: rc={r.returncode}\n{r.stdout}\n{r.stderr}")
    return r.stdout


def main():
    work = tempfile.mkdtemp(prefix="ldx_golden_")
    intgen, srcs = dc.build_dataset(work)
    if os.path.exists(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    index = {}
    for name, extra in dc.AREA_CASES:
        trg = os.path.join(work, "out_" + name)
        os.makedirs(trg)
        run_ref("ld_area.py", ["-S", srcs["area"], "-D", intgen, "-t", trg, "-f", "-p", "1"] + extra, work)
        shutil.copytree(trg, os.path.join(OUT, name))
        index[name] = sorted(dc.read_tree(trg))
    for name, extra in dc.TRIANGLE_CASES:
        trg = os.path.join(work, "out_" + name)
        os.makedirs(trg)
        run_ref("ld_triangle.py", ["-S", srcs["triangle"], "-D", intgen, "-t", trg, "-f", "-p", "1"] + extra, work)
        shutil.copytree(trg, os.path.join(OUT, name))
        index[name] = sorted(dc.read_tree(trg))
    for name, extra in dc.TRIANGLE_BIG_CASES:
        trg = os.path.join(work, "out_" + name)
        os.makedirs(trg)
        run_ref("ld_triangle.py", ["-S", srcs["triangle_big"], "-D", intgen, "-t", trg, "-f", "-p", "1"] + extra, work)
        shutil.copytree(trg, os.path.join(OUT, name))
        index[name] = sorted(dc.read_tree(trg))
    for name, extra in dc.LITE_CASES:
        os.makedirs(os.path.join(OUT, name))
        for k, (a, b) in enumerate(srcs["lite_pairs"]):
            text = run_ref("ld_lite.py", [a, b, "-D", intgen, "-f"] + extra, work)
            with open(os.path.join(OUT, name, f"pair{k}.txt"), "w") as fh:
                fh.write(text)
        index[name] = sorted(os.listdir(os.path.join(OUT, name)))
    # ---- the chrX-shaped data set (haploid males, missing calls): the general route
    intgen_x, srcs_x = dc.build_dataset_x(work)
    for cases, script, key in ((dc.AREA_X_CASES, "ld_area.py", "area"), (dc.TRIANGLE_X_CASES, "ld_triangle.py", "triangle")):
        for name, extra in cases:
            trg = os.path.join(work, "out_" + name)
            os.makedirs(trg)
            run_ref(script, ["-S", srcs_x[key], "-D", intgen_x, "-t", trg, "-f", "-p", "1"] + extra, work)
            shutil.copytree(trg, os.path.join(OUT, name))
            index[name] = sorted(dc.read_tree(trg))
    for name, extra in dc.LITE_X_CASES:
        os.makedirs(os.path.join(OUT, name))
        for k, (a, b) in enumerate(srcs_x["lite_pairs"]):
            text = run_ref("ld_lite.py", [a, b, "-D", intgen_x, "-f"] + extra, work)
            with open(os.path.join(OUT, name, f"pair{k}.txt"), "w") as fh:
                fh.write(text)
        index[name] = sorted(os.listdir(os.path.join(OUT, name)))
    # ---- records on the edges of a query's window, INFO/END intervals
    intgen_e, srcs_e, expect = dc.build_dataset_edge(work)
    for name, extra in dc.AREA_EDGE_CASES:
        trg = os.path.join(work, "out_" + name)
        os.makedirs(trg)
        run_ref("ld_area.py", ["-S", srcs_e["area"], "-D", intgen_e, "-t", trg, "-f", "-p", "1"] + extra, work)
        shutil.copytree(trg, os.path.join(OUT, name))
        index[name] = sorted(dc.read_tree(trg))
        # the crafted records decide by their intervals alone (they carry the query's haplotypes)
        text = b"".join(v for k, v in dc.read_tree(trg).items() if expect["query"] in os.path.basename(k)).decode()
        assert all(i in text for i in expect["kept"]) and not any(i in text for i in expect["not_fetched"]), (name, text[:2000])
    with open(os.path.join(OUT, "index.json"), "w") as fh:
        json.dump(index, fh, indent=1, sort_keys=True)
    n = sum(len(v) for v in index.values())
    print(f"wrote {n} golden files under {OUT}")
    shutil.rmtree(work)


if __name__ == "__main__":
    main()
