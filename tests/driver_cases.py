"""Shared by tests/golden/make_driver_golden.py (runs the UNMODIFIED reference drivers here, where
/root/reference exists) and tests/test_drivers_gpu.py (runs ld_tools_b200.drivers on the GPU box):
one deterministic synthetic 1000G-format data set and the list of driver invocations."""
import os

import numpy as np

N_SAMPLES, N_VARIANTS, SEED = 120, 520, 77
N_TRIANGLE_BIG = 320


def build_dataset(root):
    """-> (intgen_dir, {name: src_dir}).  Writes <root>/intgen/{22.vcf.gz, panel, conversion.db} and the
    source-table directories the -S option points at."""
    from ld_tools_b200.synth import conversion_rows, make_panel, make_records, synth_haplotypes, write_intgen_dir
    panel = make_panel(N_SAMPLES, seed=SEED)
    sp = sorted({p[2] for p in panel})
    pop_of_hap = np.repeat([sp.index(p[2]) for p in panel], 2)
    haps = synth_haplotypes(N_VARIANTS, 2 * N_SAMPLES, seed=SEED, n_founders=24, switch_rate=0.01, pop_of_hap=pop_of_hap)
    # LD blocks: a third of the variants are noisy copies of a close neighbour, so that r2 >= 0.8 sets exist
    rng0 = np.random.default_rng(SEED + 5)
    for i in range(12, N_VARIANTS):
        if rng0.random() < 0.35:
            src = i - int(rng0.integers(1, 12))
            flips = rng0.random(haps.shape[1]) < rng0.choice([0.0, 0.004, 0.02, 0.08])
            haps[i] = haps[src] ^ flips.astype(haps.dtype)
    recs = make_records(N_VARIANTS, seed=SEED, mean_gap=40)
    intgen = os.path.join(root, "intgen")
    write_intgen_dir(intgen, panel, recs, haps)
    addressable = [r[2] for r in conversion_rows(recs)]                 # what conversion.db knows
    rng = np.random.default_rng(SEED)
    srcs = {}
    # ld_area sources: two tables, one with meta lines, junk lines and a repeated / unknown rsID
    d = os.path.join(root, "src_area")
    os.makedirs(d)
    pick = rng.choice(len(addressable), 30, replace=False)
    with open(os.path.join(d, "gwas_hits.tsv"), "w") as fh:
        fh.write("# meta line 1\n# meta line 2\n")
        for k in pick[:18]:
            fh.write(f"chr22\t{addressable[k]}\t0.001\n")
        fh.write("no identifier on this line\nrs999999999\tunknown to the cache\n")
        fh.write(f"{addressable[pick[0]]}\trepeated\n")
    with open(os.path.join(d, "second_set.txt"), "w") as fh:
        fh.write("header\nheader\n")
        for k in pick[18:]:
            fh.write(f"{addressable[k]} and rs1 on the same line\n")
    srcs["area"] = d
    # ld_triangle source: 24 variants of one neighbourhood
    d = os.path.join(root, "src_triangle")
    os.makedirs(d)
    start = int(rng.integers(0, len(addressable) - 60))
    with open(os.path.join(d, "locus.txt"), "w") as fh:
        for k in rng.permutation(np.arange(start, start + 60))[:24]:
            fh.write(addressable[k] + "\n")
    srcs["triangle"] = d
    srcs["lite_pairs"] = [(addressable[start + 1], addressable[start + 4]), (addressable[3], addressable[-2]),
                          (addressable[start + 10], addressable[start + 11])]
    # ld_triangle source large enough for the tcgen05 engine (LDX_ENGINE_AUTO uses it from 256 variants up): 320 variants of the
    # chromosome in shuffled order.  Its own generator: the draws above (and the goldens made from them) do not move.
    d = os.path.join(root, "src_triangle_big")
    os.makedirs(d)
    rng_big = np.random.default_rng(SEED + 1000)
    with open(os.path.join(d, "region.txt"), "w") as fh:
        for k in rng_big.permutation(len(addressable))[:N_TRIANGLE_BIG]:
            fh.write(addressable[k] + "\n")
    srcs["triangle_big"] = d
    return intgen, srcs


N_SAMPLES_X, N_VARIANTS_X = 90, 420


def build_dataset_x(root):
    """A chrX-shaped data set for the general route: males haploid outside two pseudo-autosomal blocks,
    missing calls ('.', '.|.', '.|1'), a few unphased rows.  -> (intgen_dir, {name: src_dir}, lite pairs)."""
    from ld_tools_b200.synth import conversion_rows, make_panel, make_records, synth_haplotypes, write_intgen_dir
    panel = make_panel(N_SAMPLES_X, seed=SEED + 7)
    male = np.array([p[3] == "male" for p in panel])
    haps = synth_haplotypes(N_VARIANTS_X, 2 * N_SAMPLES_X, seed=SEED + 7, n_founders=16, switch_rate=0.01)
    rng = np.random.default_rng(SEED + 8)
    for i in range(8, N_VARIANTS_X):
        if rng.random() < 0.35:
            src = i - int(rng.integers(1, 8))
            haps[i] = haps[src] ^ (rng.random(haps.shape[1]) < rng.choice([0.0, 0.01, 0.05])).astype(haps.dtype)
    recs = make_records(N_VARIANTS_X, chrom="X", seed=SEED + 7, mean_gap=40)
    gt_text = []
    for v in range(N_VARIANTS_X):
        a = haps[v]
        par = v < 50 or v >= N_VARIANTS_X - 30                                   # PAR1 / PAR2: everybody diploid
        fields = [f"{a[2 * s]}|{a[2 * s + 1]}" for s in range(N_SAMPLES_X)]
        if not par:
            for s in np.flatnonzero(male):
                fields[s] = str(a[2 * s])
        u = rng.random()
        if u < 0.12:
            for s in rng.choice(N_SAMPLES_X, int(rng.integers(1, 4)), replace=False):
                fields[s] = "." if (not par and male[s]) else str(rng.choice([".|.", f".|{a[2 * s]}", f"{a[2 * s]}|."]))
        elif u < 0.16:
            fields = [f.replace("|", "/") for f in fields]
        gt_text.append("\t".join(fields).encode())
    intgen = os.path.join(root, "intgen_x")
    write_intgen_dir(intgen, panel, recs, haps, chrom="X", gt_text=gt_text)
    addressable = [r[2] for r in conversion_rows(recs)]
    srcs = {}
    d = os.path.join(root, "src_area_x")
    os.makedirs(d)
    with open(os.path.join(d, "hits_x.txt"), "w") as fh:
        for k in rng.choice(len(addressable), 24, replace=False):
            fh.write(addressable[k] + "\n")
    srcs["area"] = d
    d = os.path.join(root, "src_triangle_x")
    os.makedirs(d)
    with open(os.path.join(d, "region_x.txt"), "w") as fh:
        for k in rng.permutation(len(addressable))[:280]:
            fh.write(addressable[k] + "\n")
    srcs["triangle"] = d
    srcs["lite_pairs"] = [(addressable[3], addressable[10]), (addressable[20], addressable[200]), (addressable[150], addressable[151]),
                          (addressable[-3], addressable[100])]
    return intgen, srcs


def build_dataset_edge(root):
    """Records placed on the edges of one query's window (ld_area.py:174-177: fetch(chrom, pos - flank, pos + flank), 0-based
    half-open, overlap semantics of the tabix index): an indel straddling the left edge, one ending exactly at it, records at
    high - 1 and high, structural-variant style records whose interval comes from INFO/END (first key, middle key, END equal to
    the window start, CIEND without END).  All of them carry the query's haplotypes, so only the interval decides.
    -> (intgen_dir, {"area": src_dir}, flank)"""
    from ld_tools_b200.synth import conversion_rows, make_panel, make_records, synth_haplotypes, write_intgen_dir
    n, n_s, flank = 240, 60, 500
    panel = make_panel(n_s, seed=SEED + 21)
    haps = synth_haplotypes(n, 2 * n_s, seed=SEED + 21, n_founders=12, switch_rate=0.02)
    recs = make_records(n, chrom="7", start=1_000_000, mean_gap=30, seed=SEED + 21)
    qi = next(i for i in range(118, n) if recs[i]["id"].startswith("rs") and not recs[i]["multi"] and len(recs[i]["ref"]) == 1
              and 0.2 < haps[i].mean() < 0.8)
    q = recs[qi]
    low, high = q["pos"] - flank, q["pos"] + flank
    crafted = [
        dict(pos=low - 5, id="rs900001", ref="ACGTACGTACGT", alt="A", vt="INDEL"),                      # [low-6, low+6): straddles the left edge -> kept
        dict(pos=low - 9, id="rs900002", ref="ACGTACGTAC", alt="A", vt="INDEL"),                        # [low-10, low): ends at the edge -> not fetched
        dict(pos=high, id="rs900003", ref="A", alt="C", vt="SNP"),                                      # pos0 = high - 1 -> kept
        dict(pos=high + 1, id="rs900004", ref="A", alt="C", vt="SNP"),                                  # pos0 = high -> not fetched
        dict(pos=low - 300, id="rs900005", ref="A", alt="<CN0>", vt="SV", info_prefix=f"END={low + 20};"),          # END first in INFO -> kept
        dict(pos=low - 250, id="rs900006", ref="A", alt="<CN0>", vt="SV", info_prefix=f"END={low};"),               # interval ends at the edge -> not fetched
        dict(pos=low - 200, id="rs900007", ref="A", alt="<CN2>", vt="SV", info_prefix="CIEND=-50,400;"),            # no END key -> len(REF) -> not fetched
        dict(pos=low - 150, id="rs900008", ref="AC", alt="A", vt="SV", info_suffix=f";CIEND=0,5;END={low + 1};SVLEN=-9"),   # END in the middle -> kept
        dict(pos=high - 40, id="rs900009", ref="A", alt="<CN0>", vt="SV", info_prefix=f"END={high - 400};"),        # END before POS: ignored -> kept by len(REF)
    ]
    for c in crafted:
        c.update(chrom="7", multi=False)
    order = sorted(range(n + len(crafted)), key=lambda k: ((recs + crafted)[k]["pos"], k))
    all_recs = [(recs + crafted)[k] for k in order]
    all_haps = np.concatenate([haps, np.repeat(haps[qi:qi + 1], len(crafted), axis=0)])[order]
    intgen = os.path.join(root, "intgen_edge")
    write_intgen_dir(intgen, panel, all_recs, all_haps, chrom="7")
    addressable = [r[2] for r in conversion_rows(all_recs)]
    rng = np.random.default_rng(SEED + 22)
    d = os.path.join(root, "src_area_edge")
    os.makedirs(d)
    with open(os.path.join(d, "edge.txt"), "w") as fh:
        fh.write(q["id"] + "\n")
        for k in rng.choice(len(addressable), 9, replace=False):
            fh.write(addressable[k] + "\n")
        fh.write("rs900005\n")                         # the END= record as a query: its own window is around its POS
    return intgen, {"area": d}, {"query": q["id"], "kept": ["rs900001", "rs900003", "rs900005", "rs900008", "rs900009"],
                                 "not_fetched": ["rs900002", "rs900004", "rs900006", "rs900007"]}


AREA_EDGE_CASES = [
    ("area_edge_r2_tsv", ["-w", "500", "-l", "r_square", "-z", "0.9", "-o", "tsv"]),
    ("area_edge_dp_rsids", ["-w", "500", "-l", "d_prime", "-z", "0.9", "-o", "rsids", "-e", "eur,eas"]),
]
AREA_X_CASES = [
    ("area_x_r2_tsv", ["-w", "3000", "-l", "r_square", "-z", "0.2", "-o", "tsv"]),
    ("area_x_dp_json_male", ["-w", "2500", "-l", "d_prime", "-z", "0.8", "-o", "json", "-g", "male"]),
]
TRIANGLE_X_CASES = [
    ("triangle_x_r2", ["-o", "table"]),
    ("triangle_x_dp_thres_female", ["-o", "table", "-l", "d_prime", "-z", "0.3", "-g", "female"]),
]
LITE_X_CASES = [("lite_x_all", []), ("lite_x_afr", ["-e", "afr"])]

# name, driver, extra CLI arguments (as the reference's argparse takes them)
AREA_CASES = [
    ("area_r2_tsv", ["-m", "2", "-w", "2500", "-l", "r_square", "-z", "0.3", "-o", "tsv"]),
    ("area_dp_json_eur_afr_female", ["-m", "2", "-w", "4000", "-l", "d_prime", "-z", "0.9", "-o", "json", "-e", "eur,afr", "-g", "female"]),
    ("area_r2_rsids_default_thres", ["-m", "2", "-w", "6000", "-o", "rsids"]),
    ("area_r2_zero_thres_eas", ["-m", "2", "-w", "800", "-z", "0.0", "-e", "eas"]),
]
TRIANGLE_CASES = [
    ("triangle_r2", ["-o", "table"]),
    ("triangle_dp_thres_sas_male", ["-o", "table", "-l", "d_prime", "-z", "0.5", "-e", "sas", "-g", "male"]),
]
# 320 variants: files in -> tcgen05 all-pairs kernel -> settlement -> text kernel -> file out, against the reference's own file
TRIANGLE_BIG_CASES = [
    ("triangle_big_r2", ["-o", "table"]),
    ("triangle_big_dp_thres_eur", ["-o", "table", "-l", "d_prime", "-z", "0.4", "-e", "eur"]),
]
LITE_CASES = [("lite_all", []), ("lite_eur", ["-e", "eur"])]


def read_tree(root):
    """{relative path: bytes} of every file under root."""
    out = {}
    for base, _, files in os.walk(root):
        for f in files:
            p = os.path.join(base, f)
            with open(p, "rb") as fh:
                out[os.path.relpath(p, root)] = fh.read()
    return out
