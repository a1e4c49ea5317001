"""The ingest path end to end on one synthetic 1000G-format chromosome file:

    <chrom>.vcf.gz (BGZF)  --ldx_inflate_gz_file-->  text  --ldx_store_ingest_vcf-->  store + records
                           --ldx_store_save-->  <chrom>.vcf.gz.ldxstore  --ldx_store_load-->  store

    python tools/bench_ingest.py [--variants 100000] [--samples 2504]

Reports the wall time and rate of every stage, Python's gzip module on the same file for comparison, and what
drivers.ChromData costs on the first run (inflate + GPU ingest + cache write) and on the second (cache load).
"""
import argparse
import gzip
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", type=int, default=100_000)
    ap.add_argument("--samples", type=int, default=2504)
    args = ap.parse_args()
    from ld_tools_b200 import Context, HostText, Store, drivers
    from ld_tools_b200.synth import BgzfWriter

    nv, ns = args.variants, args.samples
    rng = np.random.default_rng(3)
    out = {"variants": nv, "samples": ns}
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "22.vcf.gz")
        t0 = time.perf_counter()
        with BgzfWriter(path) as fh:
            fh.write(b"##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + b"\t".join(b"S%d" % i for i in range(ns)) + b"\n")
            for a in range(0, nv, 2000):
                n = min(2000, nv - a)
                bits = (rng.random((n, ns, 2)) < 0.2).astype(np.uint8)
                body = np.empty((n, 4 * ns), dtype=np.uint8)
                body[:, 0::4] = bits[:, :, 0] + 48; body[:, 1::4] = 124; body[:, 2::4] = bits[:, :, 1] + 48; body[:, 3::4] = 9
                body[:, -1] = 10
                fh.write(b"".join(f"22\t{16050000 + 31 * (a + k)}\trs{1000 + a + k}\tA\tG\t100\tPASS\tAC=1;AF=0.2;AN={2 * ns};VT=SNP\tGT\t".encode()
                                  + body[k].tobytes() for k in range(n)))
        out["write_s"] = time.perf_counter() - t0
        out["gz_bytes"] = os.path.getsize(path)

        t0 = time.perf_counter()
        host = HostText(path)
        t_inf = time.perf_counter() - t0
        out["text_bytes"] = host.nbytes
        out["inflate"] = {"s": t_inf, "text_GBps": host.nbytes / t_inf / 1e9, "bgzf_parallel": host.was_bgzf, "host_cores": os.cpu_count()}
        t0 = time.perf_counter()
        host1 = HostText(path, threads=1)
        out["inflate_one_thread"] = {"s": time.perf_counter() - t0}
        host1.close()
        t0 = time.perf_counter()
        with gzip.open(path, "rb") as fh:
            ref = fh.read()
        out["python_gzip"] = {"s": time.perf_counter() - t0}
        assert len(ref) == host.nbytes and ref[:4096] == host.array[:4096].tobytes() and ref[-4096:] == host.array[-4096:].tobytes()
        del ref

        ctx = Context(0)
        Store.ingest_vcf(ctx, host.array[:1 << 24].tobytes().rsplit(b"\n", 1)[0] + b"\n", ns)[0].close()        # warm-up: arena, first launches
        t0 = time.perf_counter()
        st, rows = Store.ingest_vcf(ctx, host.array, ns)
        t_ing = time.perf_counter() - t0
        assert len(rows) == nv and not rows["status"].any() and rows["eligible"].all()
        out["gpu_ingest"] = {"s": t_ing, "variants_per_s": nv / t_ing, "text_GBps": host.nbytes / t_ing / 1e9, "from": "pageable host memory"}
        t0 = time.perf_counter(); st.save(os.path.join(d, "s.ldxstore")); t_save = time.perf_counter() - t0
        t0 = time.perf_counter(); back = Store.load(ctx, os.path.join(d, "s.ldxstore")); t_load = time.perf_counter() - t0
        assert (back.download(nv - 100, 100) == st.download(nv - 100, 100)).all()
        out["store_file"] = {"bytes": os.path.getsize(os.path.join(d, "s.ldxstore")), "save_s": t_save, "load_s": t_load, "load_variants_per_s": nv / t_load}
        back.close(); st.close(); host.close()

        t0 = time.perf_counter(); cd = drivers.ChromData(ctx, path); t_first = time.perf_counter() - t0
        cd.close()
        t0 = time.perf_counter(); cd = drivers.ChromData(ctx, path); t_second = time.perf_counter() - t0
        assert cd.n_variants == nv and cd.ids[nv - 1] == f"rs{1000 + nv - 1}"
        cd.close()
        out["drivers_ChromData"] = {"first_run_s": t_first, "second_run_from_cache_s": t_second}
        ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
