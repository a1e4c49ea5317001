"""End-to-end parity of the re-pointed drivers (ld_tools_b200/drivers.py: VCF ingest -> GPU bit packing
-> one library call per chromosome -> the reference's writers) against the output trees the UNMODIFIED
reference drivers produced on the same synthetic 1000G-format data (tests/golden/drivers/, made by
tests/golden/make_driver_golden.py where /root/reference exists).  Byte for byte: file names, headers,
row order, int-0 vs float, round(x, 4) formatting, empty-result files absent."""
import argparse
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import driver_cases as dc  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(HERE, "golden", "drivers")


@pytest.fixture(scope="module")
def data(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("ldx_drivers"))
    intgen, srcs = dc.build_dataset(root)
    return root, intgen, srcs


@pytest.fixture(scope="module")
def ctx():
    from ld_tools_b200 import Context
    c = Context(0)
    yield c
    c.close()


def parse(extra, kind):
    """The reference's own option names (cli/ld_area_cli_en.py:36-60, cli/ld_triangle_cli_en.py:40-74)."""
    ap = argparse.ArgumentParser()
    ap.add_argument("-m", dest="meta_lines_quan", type=int, default=0)
    ap.add_argument("-g", dest="gend_names", default="both")
    ap.add_argument("-e", dest="pop_names", default="all")
    if kind == "area":
        ap.add_argument("-w", dest="flank_size", type=int, default=100000)
        ap.add_argument("-l", dest="ld_thres_measure", default="r_square")
        ap.add_argument("-z", dest="ld_low_thres", type=float, default=0.8)
        ap.add_argument("-o", dest="trg_file_type", default="tsv")
    elif kind == "triangle":
        ap.add_argument("-l", dest="ld_measure", default="r_square")
        ap.add_argument("-z", dest="ld_low_thres", type=float, default=None)
        ap.add_argument("-o", dest="matrix_type", default="table")
    return vars(ap.parse_args(extra))


def assert_same_tree(got_root, name):
    want = dc.read_tree(os.path.join(GOLD, name))
    got = dc.read_tree(got_root)
    assert sorted(got) == sorted(want), (sorted(set(got) ^ set(want)))
    for rel in want:
        assert got[rel] == want[rel], f"{name}/{rel} differs from the reference driver's output"
    return len(want)


@pytest.mark.parametrize("name,extra", dc.AREA_CASES)
def test_ld_area_output_tree_identical(data, ctx, tmp_path, name, extra):
    from ld_tools_b200 import drivers
    root, intgen, srcs = data
    kw = parse(extra, "area")
    drivers.ld_area(srcs["area"], intgen, trg_top_dir_path=str(tmp_path), ctx=ctx, **kw)
    n = assert_same_tree(str(tmp_path), name)
    with open(os.path.join(GOLD, "index.json")) as fh:
        assert n == len(json.load(fh)[name])


@pytest.mark.parametrize("name,extra", dc.TRIANGLE_CASES)
def test_ld_triangle_table_identical(data, ctx, tmp_path, name, extra):
    from ld_tools_b200 import drivers
    root, intgen, srcs = data
    kw = parse(extra, "triangle")
    kw.pop("matrix_type")
    drivers.ld_triangle(srcs["triangle"], intgen, trg_top_dir_path=str(tmp_path), ctx=ctx, **kw)
    assert assert_same_tree(str(tmp_path), name) == 1


@pytest.mark.parametrize("tile_n", [0, 128])
@pytest.mark.parametrize("name,extra", dc.TRIANGLE_BIG_CASES)
def test_ld_triangle_320_variants_through_the_tensor_core_engine(data, ctx, tmp_path, name, extra, tile_n):
    """320 variants >= the 256 from which LDX_ENGINE_AUTO takes the tcgen05 engine: VCF -> store -> all-pairs kernel (64-wide
    tiles by default, 128-wide forced) -> settlement -> text kernel -> .tsv, byte for byte the UNMODIFIED reference driver's
    file (ld_triangle.py:133-230, :351-360)."""
    from ld_tools_b200 import drivers
    from ld_tools_b200._lib import TUNE_MMA_TILE_N
    root, intgen, srcs = data
    kw = parse(extra, "triangle")
    kw.pop("matrix_type")
    launches0 = ctx.launch_count
    timing0 = ctx.kernel_timing(True)
    ctx.set_tuning(TUNE_MMA_TILE_N, tile_n)
    try:
        drivers.ld_triangle(srcs["triangle_big"], intgen, trg_top_dir_path=str(tmp_path), ctx=ctx, **kw)
    finally:
        ctx.set_tuning(TUNE_MMA_TILE_N, 0)
    ms, n_allpairs = ctx.kernel_timing(False)
    assert n_allpairs >= 1 and ctx.launch_count > launches0
    assert assert_same_tree(str(tmp_path), name) == 1


@pytest.mark.parametrize("name,extra", [dc.AREA_CASES[0], dc.AREA_CASES[1]])
def test_ld_area_fan_out_over_devices_same_tree(data, tmp_path, name, extra):
    """Two workers with their own contexts (two GPUs when the box has them, else twice device 0) share the job's tables by
    chromosome -- here one chromosome, two source files -- and, for a lone table, its queries in slabs.  Same output tree."""
    import shutil
    from ld_tools_b200 import drivers
    from ld_tools_b200.engine import device_count
    root, intgen, srcs = data
    kw = parse(extra, "area")
    second = 1 if device_count() > 1 else 0
    drivers.ld_area(srcs["area"], intgen, trg_top_dir_path=str(tmp_path / "two"), devices=[0, second], **kw)
    assert_same_tree(str(tmp_path / "two"), name)
    # a lone table: region sharding of its queries over three workers
    one = tmp_path / "src_one"
    os.makedirs(one)
    shutil.copy(os.path.join(srcs["area"], "gwas_hits.tsv"), one)
    drivers.ld_area(str(one), intgen, trg_top_dir_path=str(tmp_path / "slabs"), devices=[0, second, 0], **kw)
    want = {k: v for k, v in dc.read_tree(os.path.join(GOLD, name)).items() if k.startswith("gwas_hits_in_LD")}
    got = dc.read_tree(str(tmp_path / "slabs"))
    assert sorted(got) == sorted(want) and all(got[k] == want[k] for k in want)


def test_ld_triangle_batched_tables_and_shared_matrix(data, ctx, tmp_path, monkeypatch):
    """Several (source file, chromosome) matrices of one run go through ONE batched launch (ldx_triangle_batch_dev) and the
    text kernel; a lone large matrix is cut into row slabs over the devices.  Both must write the files the one-table path
    writes -- which the goldens pin to the reference."""
    import numpy as np
    from ld_tools_b200 import drivers
    root, intgen, srcs = data
    with open(os.path.join(srcs["triangle_big"], "region.txt")) as fh:
        ids = fh.read().split()
    many = tmp_path / "src_many"
    os.makedirs(many)
    rng = np.random.default_rng(8)
    for k, n in enumerate((300, 40, 2, 129)):
        with open(many / f"set{k}.txt", "w") as fh:
            fh.write("\n".join(rng.permutation(ids)[:n]) + "\n")
    launches0 = ctx.launch_count
    drivers.ld_triangle(str(many), intgen, trg_top_dir_path=str(tmp_path / "batched"), ld_measure="d_prime", ld_low_thres=0.3, ctx=ctx)
    batched_launches = ctx.launch_count - launches0
    monkeypatch.setattr(drivers, "BATCH_MAX_VARIANTS", 1)                # every table alone, slab path
    drivers.ld_triangle(str(many), intgen, trg_top_dir_path=str(tmp_path / "single"), ld_measure="d_prime", ld_low_thres=0.3, ctx=ctx)
    a, b = dc.read_tree(str(tmp_path / "batched")), dc.read_tree(str(tmp_path / "single"))
    assert len(a) == 4 and sorted(a) == sorted(b) and all(a[k] == b[k] for k in a)
    assert batched_launches <= 4 + 3 * 4                                 # gather + all-pairs + deferred pairs (+ scatter), then 3 text kernels per table
    # the 320-variant golden table as a "large" matrix shared by three workers
    name, extra = dc.TRIANGLE_BIG_CASES[1]
    kw = parse(extra, "triangle")
    kw.pop("matrix_type")
    monkeypatch.setattr(drivers, "BATCH_MAX_VARIANTS", 100)
    from ld_tools_b200.engine import device_count
    second = 1 if device_count() > 1 else 0
    drivers.ld_triangle(srcs["triangle_big"], intgen, trg_top_dir_path=str(tmp_path / "shared"), devices=[0, second, 0], **kw)
    assert assert_same_tree(str(tmp_path / "shared"), name) == 1


# ------------------------------------------------------------------ the chrX-shaped data set: haploid males, missing calls
@pytest.fixture(scope="module")
def data_x(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("ldx_drivers_x"))
    intgen, srcs = dc.build_dataset_x(root)
    return root, intgen, srcs


@pytest.mark.parametrize("name,extra", dc.AREA_X_CASES)
def test_ld_area_chrx_general_route_tree_identical(data_x, ctx, tmp_path, name, extra):
    """Males haploid outside the pseudo-autosomal blocks, missing calls, unphased rows: the unmodified reference pairs the
    lists pysam hands it; the engine's general route must write the same files (alt_freq of a hit = var_2_alt_freq of the PAIR)."""
    from ld_tools_b200 import drivers
    root, intgen, srcs = data_x
    drivers.ld_area(srcs["area"], intgen, trg_top_dir_path=str(tmp_path), ctx=ctx, **parse(extra, "area"))
    assert assert_same_tree(str(tmp_path), name) > 3


@pytest.mark.parametrize("name,extra", dc.AREA_EDGE_CASES)
def test_ld_area_window_edges_and_info_end_tree_identical(ctx, tmp_path_factory, tmp_path, name, extra):
    """Records on the edges of a query's window and structural-variant records whose interval comes from INFO/END (the tabix
    index's definition, see tests/refshim/pysam): an indel straddling the left edge is reported, one ending exactly at it is
    not, pos0 = high - 1 is, pos0 = high is not, END= first or in the middle of INFO extends the record, CIEND= does not."""
    from ld_tools_b200 import drivers
    intgen, srcs, expect = dc.build_dataset_edge(str(tmp_path_factory.mktemp("ldx_drivers_edge")))
    drivers.ld_area(srcs["area"], intgen, trg_top_dir_path=str(tmp_path), ctx=ctx, **parse(extra, "area"))
    assert assert_same_tree(str(tmp_path), name) >= 2
    mine = b"".join(v for k, v in dc.read_tree(str(tmp_path)).items() if expect["query"] + "_" in os.path.basename(k)).decode()
    assert all(i in mine for i in expect["kept"]) and not any(i in mine for i in expect["not_fetched"])


@pytest.mark.parametrize("tile_n", [0, 128])
@pytest.mark.parametrize("name,extra", dc.TRIANGLE_X_CASES)
def test_ld_triangle_chrx_general_route_table_identical(data_x, ctx, tmp_path, name, extra, tile_n):
    from ld_tools_b200 import drivers
    from ld_tools_b200._lib import TUNE_MMA_TILE_N
    root, intgen, srcs = data_x
    kw = parse(extra, "triangle")
    kw.pop("matrix_type")
    ctx.set_tuning(TUNE_MMA_TILE_N, tile_n)
    try:
        drivers.ld_triangle(srcs["triangle"], intgen, trg_top_dir_path=str(tmp_path), ctx=ctx, **kw)
    finally:
        ctx.set_tuning(TUNE_MMA_TILE_N, 0)
    assert assert_same_tree(str(tmp_path), name) == 1


@pytest.mark.parametrize("name,extra", dc.LITE_X_CASES)
def test_ld_lite_chrx_printout_identical(data_x, ctx, name, extra):
    from ld_tools_b200 import drivers
    root, intgen, srcs = data_x
    kw = parse(extra, "lite")
    kw.pop("meta_lines_quan")
    for k, (a, b) in enumerate(srcs["lite_pairs"]):
        text = drivers.ld_lite(a, b, intgen, ctx=ctx, **kw)
        with open(os.path.join(GOLD, name, f"pair{k}.txt")) as fh:
            assert text + "\n" == fh.read(), (name, k)


def _pool_worker(args):
    """Runs in a forked multiprocessing.Pool worker: the drop-in calc_ld creates its context lazily, after the fork."""
    import os
    from ld_tools_b200 import calc_ld
    g1, g2 = args
    return os.getpid(), calc_ld(g1, g2)


def test_calc_ld_from_forked_pool_workers():
    """The reference fans out with multiprocessing.Pool (ld_area.py:336-339, ld_triangle.py:406-408; start method fork on
    Linux): the drop-in calc_ld must work in forked workers whose parent has NOT touched CUDA through libldx -- and, in a
    parent that has (this test process), in spawned workers.  Results equal the oracle's, object for object."""
    import multiprocessing as mp
    import subprocess
    import numpy as np
    from oracle import calc_ld_port
    rng = np.random.default_rng(5)
    jobs = []
    for _ in range(12):
        n = int(rng.integers(2, 400))
        jobs.append((list(map(int, rng.integers(0, 2, n))), list(map(int, rng.integers(0, 2, n)))))
    jobs.append(([0, 1, None, 1], [1, 1, 0, None]))
    want = [calc_ld_port.calc_ld(a, b) for a, b in jobs]
    # (1) a fresh parent that forks before anything initialises CUDA: exactly the reference's situation
    code = ("import sys, json, multiprocessing as mp\n"
            "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import test_drivers_gpu as t\n"
            "jobs = json.loads(sys.stdin.read())\n"
            "with mp.get_context('fork').Pool(2) as pool:\n"
            "    out = pool.map(t._pool_worker, [tuple(j) for j in jobs])\n"
            "print(json.dumps([[p, [[k, repr(v)] for k, v in d.items()]] for p, d in out]))\n") % (os.path.dirname(HERE), HERE)
    r = subprocess.run([sys.executable, "-c", code], input=json.dumps(jobs), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    got = json.loads(r.stdout.strip().splitlines()[-1])
    assert len({p for p, _ in got}) >= 1 and os.getpid() not in {p for p, _ in got}
    for (_, d), w in zip(got, want):
        assert d == [[k, repr(v)] for k, v in w.items()]
    # (2) spawned workers of a parent that already holds a CUDA context
    with mp.get_context("spawn").Pool(2) as pool:
        out = pool.map(_pool_worker, jobs)
    for (_, d), w in zip(out, want):
        assert list(d) == list(w) and all(type(d[k]) is type(w[k]) and d[k] == w[k] for k in w)


@pytest.mark.parametrize("name,extra", dc.LITE_CASES)
def test_ld_lite_printout_identical(data, ctx, name, extra):
    from ld_tools_b200 import drivers
    root, intgen, srcs = data
    kw = parse(extra, "lite")
    kw.pop("meta_lines_quan")
    for k, (a, b) in enumerate(srcs["lite_pairs"]):
        text = drivers.ld_lite(a, b, intgen, ctx=ctx, **kw)
        with open(os.path.join(GOLD, name, f"pair{k}.txt")) as fh:
            assert text + "\n" == fh.read()


def test_store_cache_round_trip_and_staleness(data, ctx, monkeypatch):
    """The packed store kept next to the VCF (ldx_store_save / ldx_store_load + the text columns): a second run
    loads it instead of re-reading the VCF and sees exactly the same data; a VCF that changed invalidates it."""
    import numpy as np
    from ld_tools_b200 import drivers
    root, intgen, srcs = data
    vcf = os.path.join(intgen, "22.vcf.gz")
    paths = drivers.ChromData._cache_paths(vcf)
    for p in paths:
        if os.path.exists(p):
            os.remove(p)
    fresh = drivers.ChromData(ctx, vcf, cache=False)
    assert not any(os.path.exists(p) for p in paths)
    first = drivers.ChromData(ctx, vcf)                      # ingests and writes the cache
    assert all(os.path.exists(p) for p in paths)
    ingests = []
    real_ingest = drivers.ChromData._ingest
    monkeypatch.setattr(drivers.ChromData, "_ingest", lambda self, c, p: (ingests.append(p), real_ingest(self, c, p))[1])
    cached = drivers.ChromData(ctx, vcf)                     # must come from the cache
    assert ingests == []
    for cd in (first, cached):
        assert (cd.store.download() == fresh.store.download()).all()
        assert cd.samples == fresh.samples and list(cd.ids) == list(fresh.ids) and list(cd.refs) == list(fresh.refs)
        assert list(cd.alts) == list(fresh.alts) and list(cd.vts) == list(fresh.vts) and list(cd.multi) == list(fresh.multi)
        assert (cd.pos == fresh.pos).all() and (cd.pos0 == fresh.pos0).all() and (cd.end0 == fresh.end0).all()
        assert cd.max_ref_len == fresh.max_ref_len and cd.n_samples == fresh.n_samples
    # the annotations travelled with the planes: the same window scan on both
    names = fresh.samples[::3]
    q = np.flatnonzero([bool(drivers.RS_RE.match(i)) and not m for i, m in zip(fresh.ids, fresh.multi)])[5:25]
    from ld_tools_b200 import shard
    lo, hi, ws, we = shard.window_bounds(fresh.pos0, fresh.max_ref_len, fresh.pos[q], 3000)
    hits = []
    for cd in (fresh, cached):
        cd.select_samples(names)
        hits.append(cd.store.window(q, lo, hi, ws, we, "r_square", 0)[0])
    assert len(hits[0]) > 0 and (hits[0] == hits[1]).all()
    # a VCF with another mtime: the cache is stale and the file is read again
    st = os.stat(vcf)
    os.utime(vcf, ns=(st.st_atime_ns, st.st_mtime_ns + 1_000_000_000))
    again = drivers.ChromData(ctx, vcf)
    assert ingests == [vcf]
    assert (again.store.download() == fresh.store.download()).all()
    # a truncated store file is refused (and rebuilt), never half-loaded
    with open(paths[0], "r+b") as fh:
        fh.truncate(os.path.getsize(paths[0]) // 2)
    rebuilt = drivers.ChromData(ctx, vcf)
    assert ingests == [vcf, vcf]
    assert (rebuilt.store.download() == fresh.store.download()).all()
    for cd in (fresh, first, cached, again, rebuilt):
        cd.close()


def _host_parse(raw, n_samples):
    """What the GPU ingest must reproduce, stated with Python string methods (the drivers' former per-line loop)."""
    import re
    rows = []
    for line in raw.split(b"\n"):
        line = line.rstrip(b"\r")
        if not line or line.startswith(b"#"):
            continue
        f = line.split(b"\t")
        ok = len(f) >= 9 + n_samples and len(b"\t".join(f[9:9 + n_samples])) >= 4 * n_samples - 1
        pos = int(f[1]) if len(f) > 1 and f[1].isdigit() else 0
        rid = f[2].decode() if len(f) > 2 else ""
        keys = [x.split("=")[0] for x in (f[7].decode().split(";") if len(f) > 7 else [])]
        rs = bool(re.match(r"rs\d+$", rid))
        multi = "MULTI_ALLELIC" in keys
        span = len(f[3]) if len(f) > 3 else 0
        info = f[7].decode() if len(f) > 7 else ""
        m = re.match(r"END=(\d+)", info) or re.search(r";END=(\d+)", info)
        if m and int(m.group(1)) > pos - 1:
            span = int(m.group(1)) - (pos - 1)
        rows.append({"pos": pos, "id": rid, "ref": f[3].decode() if len(f) > 3 else "", "alt": f[4].decode() if len(f) > 4 else "", "span": span,
                     "rs": rs, "multi": multi, "ok": ok, "gts": f[9:9 + n_samples] if ok else None,
                     "vt": ([x[3:] for x in f[7].decode().split(";") if x.startswith("VT=")] or [""])[0] if len(f) > 7 else ""})
    return rows


def _check_ingest(ctx, raw, n_samples):
    import numpy as np
    from ld_tools_b200 import Store
    want = _host_parse(raw, n_samples)
    st, rows = Store.ingest_vcf(ctx, raw, n_samples)
    assert len(rows) == len(want) == st.n_variants
    blob, off = Store.vcf_fixed_columns(ctx._lib, raw, rows)
    blob = blob.tobytes()
    planes = st.download() if len(want) else np.zeros((0, st.stride_words), dtype="<u8")
    for k, (r, w) in enumerate(zip(rows, want)):
        assert bool(r["status"] & 10) == (not w["ok"]), (k, r, w)       # bit 1: not a record line, bit 3: genotype fields no parser takes
        if not w["ok"]:
            assert r["eligible"] == 0 and not planes[k].any()
            continue
        assert r["pos"] == w["pos"] and r["ref_len"] == w["span"] and bool(r["multi"]) == w["multi"]
        assert bool(r["eligible"]) == (w["rs"] and not w["multi"])
        assert r["idnum"] == (int(w["id"][2:]) if w["rs"] else -1 - k)
        rec = blob[off[k]:off[k + 1]]
        assert rec[r["id_off"]:r["ref_off"] - 1].decode() == w["id"] and rec[r["ref_off"]:r["alt_off"] - 1].decode() == w["ref"]
        assert rec[r["alt_off"]:].split(b"\t")[0].decode() == w["alt"] and len(rec) == r["gt_off"]
        bits = np.zeros(st.stride_words * 64, dtype=np.uint8)
        bad = False
        for s_i, g in enumerate(w["gts"]):
            bad |= len(g) < 3 or g[0:1] not in (b"0", b"1") or g[2:3] not in (b"0", b"1") or g[1:2] != b"|"
            bits[2 * s_i] = g[0:1] == b"1"
            bits[2 * s_i + 1] = g[2:3] == b"1"
        assert bool(r["status"] & 1) == bad, (k, w["gts"][:4])
        assert (np.packbits(bits, bitorder="little").view("<u8") == planes[k]).all(), k
    st.close()
    return rows


def test_gpu_vcf_ingest_matches_a_host_parse(data, ctx):
    """ldx_store_ingest_vcf: newline index, field split, POS / rs number / MULTI_ALLELIC / len(REF) and the packed
    genotypes, against a parse written with Python string methods -- on the synthetic chr22 and on hand-made corner cases."""
    import gzip
    root, intgen, srcs = data
    with gzip.open(os.path.join(intgen, "22.vcf.gz"), "rb") as fh:
        raw = fh.read()
    hdr = [ln for ln in raw.split(b"\n") if ln.startswith(b"#CHROM")][0]
    rows = _check_ingest(ctx, raw, len(hdr.split(b"\t")) - 9)
    assert rows["eligible"].sum() > 0 and rows["multi"].sum() > 0 and (rows["idnum"] < 0).sum() > 0 and (rows["ref_len"] > 1).sum() > 0
    gt = lambda s: "\t".join(s.split())                                                     # noqa: E731
    fixed = "22\t{pos}\t{id}\t{ref}\t{alt}\t100\tPASS\t{info}\tGT\t"
    lines = ["##fileformat=VCFv4.1", "##INFO=<ID=VT>", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tA\tB\tC",
             fixed.format(pos=1, id="rs1", ref="A", alt="G", info="AC=1;VT=SNP") + gt("0|1 1|1 0|0"),
             fixed.format(pos=248956422, id="rs4294967296", ref="ACGT", alt="A", info="VT=INDEL;MULTI_ALLELIC") + gt("1|0 0|0 1|1"),
             fixed.format(pos=77, id="esv123", ref="A", alt="<CN0>", info="MULTI_ALLELIC=1;VT=SV") + gt("0|0 0|0 0|1"),
             fixed.format(pos=78, id=".", ref="A", alt="C,T", info="XMULTI_ALLELIC;MULTI_ALLELICX;VT=SNP") + gt("1|1 1|1 1|1"),
             fixed.format(pos=79, id="rs12x", ref="A", alt="C", info=".") + gt("0|1 .|1 0/1"),
             fixed.format(pos=500, id="rs21", ref="A", alt="<CN0>", info="END=1234;VT=SV") + gt("0|1 1|0 0|0"),     # interval from INFO/END
             fixed.format(pos=501, id="rs22", ref="ACG", alt="A", info="CIEND=-5,5;VT=SV;END=2000;SVLEN=9") + gt("0|1 1|0 0|0"),
             fixed.format(pos=502, id="rs23", ref="AC", alt="A", info="CIEND=0,9;XEND=7000;VT=SV") + gt("0|1 1|0 0|0"),    # no END key
             fixed.format(pos=503, id="rs24", ref="AC", alt="A", info="VT=SV;END=400") + gt("0|1 1|0 0|0"),            # END before POS: ignored
             fixed.format(pos=80, id="rs9", ref="A", alt="C", info="VT=SNP") + gt("0|1 1|0"),          # a column short
             "22\t81\trs10\tA",                                                                    # truncated line
             "",
             fixed.format(pos=82, id="rs0011", ref="AT", alt="A", info="VT=INDEL;AF=0.5") + gt("1|1 0|1 1|0")]
    for text in ("\n".join(lines) + "\n", "\n".join(lines), "\r\n".join(lines) + "\r\n", "\n".join(lines[:3]) + "\n"):
        _check_ingest(ctx, text.encode(), 3)


@pytest.mark.parametrize("which", ["chr22", "chrX"])
def test_slab_ingest_of_the_file_equals_the_whole_text_ingest(data, data_x, ctx, tmp_path, which):
    """ldx_store_ingest_vcf_file with slabs of 4 KiB (dozens of slabs, every one cut inside a line and carried over) builds
    the same store, record table and fixed columns as ldx_store_ingest_vcf on the whole text -- general-route rows included --
    and the drivers' first-run path (ChromData._ingest) goes through it."""
    import gzip
    import numpy as np
    from ld_tools_b200 import Store
    from ld_tools_b200.drivers import ChromData
    from ld_tools_b200.synth import BgzfWriter
    intgen, name = (data[1], "22.vcf.gz") if which == "chr22" else (data_x[1], "X.vcf.gz")
    with gzip.open(os.path.join(intgen, name), "rb") as fh:
        raw = fh.read()
    n_samples = len(ChromData._header_samples(os.path.join(intgen, name)))

    class SmallBlocks(BgzfWriter):           # members of 1,500 bytes: a 4 KiB slab holds two of them and six or seven lines
        BLOCK = 1500
    path = str(tmp_path / name)
    with SmallBlocks(path) as fh:
        fh.write(raw)
    assert len(raw) > 30 * 4096
    whole, rows_w = Store.ingest_vcf(ctx, raw, n_samples)
    blob_w, off_w = Store.vcf_fixed_columns(ctx._lib, raw, rows_w)
    for slab in (4096, 30_000, 0):
        st, rows, blob, off, text_bytes = Store.ingest_vcf_file(ctx, path, n_samples, slab_bytes=slab, threads=3)
        assert text_bytes == len(raw) and st.n_variants == whole.n_variants == len(rows)
        assert rows.tobytes() == rows_w.tobytes()                         # line_off is file-wide, idnum of unnamed records too
        assert (off == off_w).all() and blob.tobytes() == blob_w.tobytes()
        assert (st.download() == whole.download()).all()
        st.select_all(); whole.select_all()
        for a, b in zip(st.row_counts(), whole.row_counts()):
            assert np.array_equal(a, b)
        for a, b in zip(st.counts(), whole.counts()):
            assert np.array_equal(a, b)
        rows_all = np.arange(min(st.n_variants, 700), dtype=np.int64)
        assert np.array_equal(st.triangle(rows_all, "r_square")[0], whole.triangle(rows_all, "r_square")[0])
        q = np.arange(5, st.n_variants, 97, dtype=np.int64)
        lo, hi = np.maximum(q - 150, 0), np.minimum(q + 150, st.n_variants)
        ws, we = np.zeros(len(q), np.int32), np.full(len(q), 2**31 - 1, np.int32)
        h1, _ = st.window(q, lo, hi, ws, we, "r_square", 2000)
        h2, _ = whole.window(q, lo, hi, ws, we, "r_square", 2000)
        assert np.array(h1).tobytes() == np.array(h2).tobytes() and len(h1) > 0
        st.close()
    whole.close()
    cd = ChromData(ctx, path, cache=False)
    assert cd.rows.tobytes() == rows_w.tobytes() and cd._blob == blob_w.tobytes()
    cd.close()


def test_command_line_front_end_writes_the_reference_tree(data, tmp_path, capsys):
    """`python -m ld_tools_b200 ld_area ...` with the reference's own option letters (cli/ld_area_cli_en.py:36-60): the same
    output tree as the function; -p picks the number of GPUs (1 here, or however many the box has)."""
    from ld_tools_b200.__main__ import main
    root, intgen, srcs = data
    name, extra = dc.AREA_CASES[0]
    main(["ld_area", "-S", srcs["area"], "-D", intgen, "-t", str(tmp_path), "-f", "-p", "2"] + extra)
    assert "computation time" in capsys.readouterr().out
    assert assert_same_tree(str(tmp_path), name) > 3
    a, b = srcs["lite_pairs"][0]
    main(["ld_lite", a, b, "-D", intgen, "-f"])
    with open(os.path.join(GOLD, "lite_all", "pair0.txt")) as fh:
        assert capsys.readouterr().out == fh.read()


def test_slab_ingest_edge_cases(data, ctx, tmp_path):
    """ldx_store_ingest_vcf_file at its edges: a file with no record, a last line without its newline, a plain (non-BGZF) gzip
    file, CRLF line ends, and a slab smaller than a BGZF member or a line (the buffer grows)."""
    import gzip
    import numpy as np
    from ld_tools_b200 import Store
    from ld_tools_b200.synth import BgzfWriter
    root, intgen, srcs = data
    with gzip.open(os.path.join(intgen, "22.vcf.gz"), "rb") as fh:
        raw = fh.read()
    lines = raw.split(b"\n")
    n_samples = len([ln for ln in lines if ln.startswith(b"#CHROM")][0].split(b"\t")) - 9
    header = b"\n".join(ln for ln in lines if ln.startswith(b"#")) + b"\n"
    body = [ln for ln in lines if ln and not ln.startswith(b"#")]

    def write(name, text, bgzf=True):
        path = str(tmp_path / name)
        if bgzf:
            with BgzfWriter(path) as fh:
                fh.write(text)
        else:
            with gzip.open(path, "wb") as fh:
                fh.write(text)
        return path

    whole, rows_w = Store.ingest_vcf(ctx, raw, n_samples)
    planes_w = whole.download()
    # no record at all
    st, rows, blob, off, n_text = Store.ingest_vcf_file(ctx, write("empty.vcf.gz", header), n_samples)
    assert st.n_variants == 0 and len(rows) == 0 and off.tolist() == [0] and n_text == len(header)
    st.close()
    # the last line without its newline; small slabs
    text = header + b"\n".join(body)
    st, rows, blob, off, n_text = Store.ingest_vcf_file(ctx, write("nonl.vcf.gz", text), n_samples, slab_bytes=5000)
    assert st.n_variants == len(body) == whole.n_variants and (st.download() == planes_w).all() and n_text == len(text)
    assert (rows["pos"] == rows_w["pos"]).all() and (rows["status"] == rows_w["status"]).all()
    st.close()
    # a plain gzip stream (no block table): read at once, same store
    st, rows, blob, off, n_text = Store.ingest_vcf_file(ctx, write("plain.vcf.gz", raw, bgzf=False), n_samples, slab_bytes=5000)
    assert st.n_variants == whole.n_variants and (st.download() == planes_w).all() and rows.tobytes() == rows_w.tobytes()
    st.close()
    # CRLF line ends
    crlf = raw.replace(b"\n", b"\r\n")
    st, rows, blob, off, n_text = Store.ingest_vcf_file(ctx, write("crlf.vcf.gz", crlf), n_samples, slab_bytes=7000)
    assert st.n_variants == whole.n_variants and (st.download() == planes_w).all() and (rows["pos"] == rows_w["pos"]).all()
    assert (rows["ref_len"] == rows_w["ref_len"]).all() and (rows["eligible"] == rows_w["eligible"]).all()
    st.close()
    # a slab smaller than a BGZF member and than a record line: the buffer grows to what a step needs
    wide = header + b"\n".join(b + b"\t" + b"0|0\t" * 4000 for b in body[:3]) + b"\n"            # ~16 KB lines
    st, rows, blob, off, n_text = Store.ingest_vcf_file(ctx, write("wide.vcf.gz", wide), n_samples, slab_bytes=4096)
    assert st.n_variants == 3 and (st.download() == planes_w[:3]).all() and n_text == len(wide)
    st.close()
    whole.close()


def test_a_second_device_from_fresh_host_threads(data, tmp_path):
    """Every library call selects its context's device itself: a host thread that has never touched CUDA starts on device 0, and
    an allocation made there for a context of device 1 is unreadable from device 1.  Each call below runs in its OWN fresh thread
    against a context on the last device of the box (skipped on a one-GPU box)."""
    import threading
    import numpy as np
    from ld_tools_b200 import Context, Store, drivers
    from ld_tools_b200.engine import device_count, threshold_e4
    from ld_tools_b200.synth import random_planes
    from oracle import ld_oracle
    if device_count() < 2:
        pytest.skip("needs two GPUs")
    dev = device_count() - 1
    state, errors = {}, []

    def in_thread(fn):
        def run():
            try:
                fn()
            except BaseException as e:      # noqa: BLE001
                errors.append(e)
        t = threading.Thread(target=run)
        t.start()
        t.join()
        if errors:
            raise errors[0]

    n_var, n_hap = 900, 1300
    planes = random_planes(n_var, n_hap, seed=3)
    sel = np.sort(np.random.default_rng(1).choice(n_hap, 600, replace=False))
    mask = ld_oracle.mask_from_haplotypes(sel, n_hap)
    rows = np.arange(n_var, dtype=np.int64)
    want = ld_oracle.packed_of(ld_oracle.triangle(planes, mask, n_hap, rows))
    in_thread(lambda: state.update(ctx=Context(dev)))
    in_thread(lambda: state.update(st=Store.from_planes(state["ctx"], planes, n_hap)))
    in_thread(lambda: state["st"].set_mask(mask))
    in_thread(lambda: state.update(tri=state["st"].triangle(rows)[0]))
    assert (state["tri"] == want).all()
    in_thread(lambda: state.update(vals=state["st"].triangle_values(rows, "d_prime")))
    assert (state["vals"] == (want >> 16).astype(np.uint16)).all()
    in_thread(lambda: state.update(hits=state["st"].triangle_hits(rows, "r_square", threshold_e4(0.3))))
    assert len(state["hits"]) == int(((want & 0x3FFF) >= 3000).sum())
    in_thread(lambda: state.update(sub=state["st"].subset(sel)))
    pos0 = (np.arange(n_var) * 20 + 50).astype(np.int32)

    def scan():
        sub = state["sub"]
        sub.set_annotations(pos0, pos0 + 1, np.arange(n_var, dtype=np.int64), np.ones(n_var, np.uint8))
        q = np.arange(10, n_var, 37, dtype=np.int64)
        lo, hi = np.maximum(q - 100, 0), np.minimum(q + 100, n_var)
        state["win"] = sub.window(q, lo, hi, np.zeros(len(q), np.int32), np.full(len(q), 2**31 - 1, np.int32), "r_square", 0)[0]
        state["q"] = q
    in_thread(scan)
    k = 3
    q = int(state["q"][k])
    got = state["win"][state["win"]["query"] == k]
    idx = [max(q, r) * (max(q, r) - 1) // 2 + min(q, r) for r in got["row"]]
    assert got["row"].tolist() == [r for r in range(q - 100, q + 100) if r != q]           # threshold 0: every candidate but the query
    assert ((got["packed"] & 0x3FFF) == (want[idx] & 0x3FFF)).all()                          # r2 is symmetric in the pair
    # the drivers on that device alone
    root, intgen, srcs = data
    name, extra = dc.AREA_CASES[0]
    in_thread(lambda: drivers.ld_area(srcs["area"], intgen, trg_top_dir_path=str(tmp_path), devices=[dev], **parse(extra, "area")))
    assert assert_same_tree(str(tmp_path), name) > 3
    in_thread(lambda: (state["sub"].close(), state["st"].close(), state["ctx"].close()))
