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
