"""configs[2]'s job through the DRIVER, files in and files out: `ld_area -w 500000 -l r_square -z 0.8 -e eur -o tsv`.

    python tools/bench_driver_area.py [--variants 40000] [--queries 1000]

The synthetic chromosome is shorter than configs[2]'s (40,000 variants instead of 1.1 M: writing 11 GB of VCF text from
Python is not a benchmark of anything) but as dense, so every query's +/-500 kb window still holds ~30,000 records and
the job is the same ~3e7 candidate pairs.  Timed: ld_tools_b200.drivers.ld_area cold (BGZF inflate, GPU ingest, store
cache) and warm (SQLite lookups, store cache, ONE window scan for all queries, one .tsv per query with kept pairs).
About two dozen queries' files are then re-derived on the host from the haplotypes the VCF was written from: the reference's
filters (ld_area.py:215-249) restated + the oracle's finalise_counts -- same rows, same values, same order.
"""
import argparse
import json
import os
import re
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", type=int, default=40000)
    ap.add_argument("--queries", type=int, default=1000)
    ap.add_argument("--samples", type=int, default=2504)
    ap.add_argument("--flank", type=int, default=500000)
    ap.add_argument("--thres", type=float, default=0.8, help="-z of the job; 0.0 keeps every record in the windows (the writers at scale)")
    args = ap.parse_args()
    from ld_tools_b200 import Context, drivers
    from ld_tools_b200.synth import conversion_rows, make_panel, make_records, synth_haplotypes, write_intgen_dir
    from oracle import calc_ld_port                                       # checker only

    out = {"variants_in_vcf": args.variants, "samples": args.samples, "flank": args.flank}
    with tempfile.TemporaryDirectory() as root:
        t0 = time.perf_counter()
        panel = make_panel(args.samples)
        sp = sorted({p[2] for p in panel})
        pop_of_hap = np.repeat([sp.index(p[2]) for p in panel], 2)
        haps = synth_haplotypes(args.variants, 2 * args.samples, seed=21, pop_of_hap=pop_of_hap)
        rng = np.random.default_rng(9)
        for i in range(12, args.variants):                                # LD blocks: noisy copies of close neighbours
            if rng.random() < 0.3:
                flips = rng.random(haps.shape[1]) < rng.choice([0.0, 0.004, 0.02])
                haps[i] = haps[i - int(rng.integers(1, 12))] ^ flips.astype(haps.dtype)
        recs = make_records(args.variants, seed=21)
        intgen = os.path.join(root, "intgen")
        write_intgen_dir(intgen, panel, recs, haps)
        addressable = conversion_rows(recs)
        picked = rng.choice(len(addressable), min(args.queries, len(addressable)), replace=False).tolist()
        src = os.path.join(root, "src")
        os.makedirs(src)
        with open(os.path.join(src, "hits.txt"), "w") as fh:
            for k in picked:
                fh.write(addressable[k][2] + "\n")
        out["dataset_s"] = time.perf_counter() - t0
        out["vcf_gz_bytes"] = os.path.getsize(os.path.join(intgen, "22.vcf.gz"))

        ctx = Context(0)
        runs = []
        for k in range(3):
            trg = os.path.join(root, f"out{k}")
            t0 = time.perf_counter()
            drivers.ld_area(src, intgen, trg, pop_names="eur", flank_size=args.flank, ld_thres_measure="r_square",
                            ld_low_thres=args.thres, trg_file_type="tsv", ctx=ctx)
            runs.append(time.perf_counter() - t0)
        ctx.close()
        out["driver_s"] = {"first_run_inflate_ingest_cache": runs[0], "later_runs_from_store_cache": runs[1:]}
        chr_dir = os.path.join(root, "out2", "hits_in_LD", "22")
        files = sorted(os.listdir(chr_dir))
        out["files_written"] = len(files)
        out["kept_pairs"] = sum(sum(1 for _ in open(os.path.join(chr_dir, f))) - 3 for f in files)

        # ---- candidate pairs of the job and two queries' files re-derived on the host
        pos = np.array([r["pos"] for r in recs])
        pos0, end0 = pos - 1, pos - 1 + np.array([len(r["ref"]) for r in recs])
        eligible = np.array([bool(re.match(r"rs\d+$", r["id"])) and not r["multi"] for r in recs])      # ld_area.py:223-224
        rec_of = {(r["pos"], r["id"]): i for i, r in enumerate(recs)}
        addr_by_id = {a[2]: a for a in addressable}
        n_cand = 0
        for k in picked:
            q = rec_of[(addressable[k][1], addressable[k][2])]
            ws, we = max(0, int(pos[q]) - args.flank), int(pos[q]) + args.flank
            n_cand += int(((pos0 < we) & (end0 > ws)).sum())
        out["records_in_windows"] = n_cand
        out["records_in_windows_per_s_files_in_files_out"] = n_cand / min(runs[1:])
        eur = np.flatnonzero(np.repeat([p[2] == "EUR" for p in panel], 2))
        h = haps[:, eur].astype(np.int64)
        n_hap = len(eur)
        n1 = h.sum(axis=1)
        checked = 0
        for f in files[::max(1, len(files) // 24)]:
            qid = f.split("_chr")[0]
            q = rec_of[(addr_by_id[qid][1], qid)]
            ws, we = max(0, int(pos[q]) - args.flank), int(pos[q]) + args.flank
            cand = np.flatnonzero((pos0 < we) & (end0 > ws) & eligible)                                   # :215-224
            n11 = h[cand] @ h[q]
            want = []
            for j, c11 in zip(cand.tolist(), n11.tolist()):
                if recs[j]["id"] == qid:                                                                  # :222
                    continue
                r2, dp, p_a, p_b, _ = calc_ld_port.finalise_counts(n_hap, c11, int(n1[q]), n_hap - int(n1[q]), int(n1[j]), n_hap - int(n1[j]))
                if r2 < args.thres:                                                                       # :248
                    continue
                want.append("\t".join(map(str, [recs[j]["pos"], recs[j]["id"], recs[j]["ref"], recs[j]["alt"], recs[j]["vt"],
                                                 p_b, r2, dp, recs[j]["pos"] - recs[q]["pos"]])))
            with open(os.path.join(chr_dir, f)) as fh:
                got = fh.read().split("\n")[3:-1]
            assert got == want, (f, len(got), len(want))
            checked += len(want)
        out["rows_checked_against_oracle"] = checked
    print(json.dumps(out))


if __name__ == "__main__":
    main()
