"""configs[1] through the DRIVER, files in and file out: `ld_triangle -o table` for 2,000 rsIDs, 2504 samples.

    python tools/bench_driver_triangle.py [--variants 6000] [--pick 2000]

Builds a synthetic 1000G-format directory (<chrom>.vcf.gz as BGZF, panel, conversion.db) and a source table of
`--pick` rsIDs, then times ld_tools_b200.drivers.ld_triangle -- create_src_dict (SQLite), chromosome data (first
run: inflate + GPU ingest + store cache; later runs: the cache), the all-pairs kernel + settlement + the table
writer in one library call, and the .tsv on disk -- cold and warm.  A bounded sample of the table's cells is
checked against the oracle's calc_ld on the haplotypes the VCF was written from.  For scale: the reference's own
loop costs two tabix fetches, 2 x 2504 pysam lookups and one pure-Python calc_ld (~0.4 ms alone) PER PAIR.
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", type=int, default=6000)
    ap.add_argument("--pick", type=int, default=2000)
    ap.add_argument("--samples", type=int, default=2504)
    ap.add_argument("--runs", type=int, default=4)
    args = ap.parse_args()
    from ld_tools_b200 import Context, drivers
    from ld_tools_b200.synth import conversion_rows, make_panel, make_records, synth_haplotypes, write_intgen_dir
    from oracle import calc_ld_port                                       # checker only

    out = {"variants_in_vcf": args.variants, "samples": args.samples}
    with tempfile.TemporaryDirectory() as root:
        t0 = time.perf_counter()
        panel = make_panel(args.samples)
        haps = synth_haplotypes(args.variants, 2 * args.samples, seed=11)
        recs = make_records(args.variants, seed=11)
        intgen = os.path.join(root, "intgen")
        write_intgen_dir(intgen, panel, recs, haps)
        addressable = conversion_rows(recs)
        rng = np.random.default_rng(5)
        picked = sorted(rng.choice(len(addressable), min(args.pick, len(addressable)), replace=False).tolist())
        src = os.path.join(root, "src")
        os.makedirs(src)
        with open(os.path.join(src, "locus.txt"), "w") as fh:
            for k in rng.permutation(picked):
                fh.write(addressable[k][2] + "\n")
        out["dataset_s"] = time.perf_counter() - t0
        out["vcf_gz_bytes"] = os.path.getsize(os.path.join(intgen, "22.vcf.gz"))
        v = len(picked)
        out["variants_in_matrix"], out["pairs"] = v, v * (v - 1) // 2

        ctx = Context(0)
        runs = []
        for k in range(args.runs):
            trg = os.path.join(root, f"out{k}")
            t0 = time.perf_counter()
            drivers.ld_triangle(src, intgen, trg, ld_measure="r_square", ctx=ctx)
            runs.append(time.perf_counter() - t0)
            if k + 1 < args.runs:
                os.remove(os.path.join(trg, "locus_LD_matr", "locus_chr22_r.tsv"))      # a 20,000-variant table is 1.2 GB
        out["driver_s"] = {"first_run_inflate_ingest_cache": runs[0], "later_runs_from_store_cache": runs[1:]}
        tsv = os.path.join(root, f"out{args.runs - 1}", "locus_LD_matr", "locus_chr22_r.tsv")
        out["tsv_bytes"] = os.path.getsize(tsv)
        out["pairs_per_s_files_in_file_out"] = out["pairs"] / min(runs[1:])

        # ---- sampled parity against the oracle, from the haplotypes the VCF was written from
        with open(tsv) as fh:
            lines = fh.read().split("\n")
        ids = lines[2].split("\t")[2:]
        row_of_id = {r["id"]: i for i, r in enumerate(recs)}
        checked = 0
        for r in rng.choice(np.arange(1, v), 12, replace=False).tolist():
            cells = lines[4 + r].split("\t")[2:]
            for c in rng.choice(r, min(r, 20), replace=False).tolist():
                ref = calc_ld_port.calc_ld(haps[row_of_id[ids[r]]].tolist(), haps[row_of_id[ids[c]]].tolist())
                assert cells[c] == str(ref["r_square"]), (r, c, cells[c], ref)
                checked += 1
        out["cells_checked_against_oracle"] = checked
        ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
