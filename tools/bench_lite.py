"""BASELINE configs[0]: ld_lite -- r2 / D' for 100 variant pairs (5008 haplotypes), latency and parity.

    python tools/bench_lite.py [--pairs 100]

Three ways to the same 100 results, each checked against the reference algorithm (oracle/calc_ld_port.py, the
pure-Python port that the golden vectors of backend/calc_ld.py pin):
  scalar    the drop-in calc_ld(list, list) -> dict, one call per pair (what ld_lite.py:143 does; genotype lists
            encoded and uploaded every call)
  store     Store.pairs on a resident bit-plane store, one call per pair (ld_lite once the chromosome is loaded)
  batch     Store.pairs, all pairs in one call
and the reference port itself on one host core.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=100)
    args = ap.parse_args()
    from ld_tools_b200 import Context, Store, calc_ld
    from ld_tools_b200.engine import dprime_value, r2_value
    from ld_tools_b200.synth import pack_bits, synth_haplotypes
    from oracle import calc_ld_port

    n_hap, n_var = 5008, 4000
    h = synth_haplotypes(n_var, n_hap, seed=20130502)
    rng = np.random.default_rng(5)
    ia = rng.integers(0, n_var, args.pairs)
    ib = np.clip(ia + rng.integers(-30, 31, args.pairs), 0, n_var - 1)        # neighbours: a spread of LD values
    lists = {int(v): list(map(int, h[v])) for v in set(ia.tolist()) | set(ib.tolist())}

    t0 = time.perf_counter()
    want = [calc_ld_port.calc_ld(lists[int(a)], lists[int(b)]) for a, b in zip(ia, ib)]
    t_ref = (time.perf_counter() - t0) / args.pairs

    calc_ld(lists[int(ia[0])], lists[int(ib[0])])                              # context creation, first launch
    t0 = time.perf_counter()
    got = [calc_ld(lists[int(a)], lists[int(b)]) for a, b in zip(ia, ib)]
    t_scalar = (time.perf_counter() - t0) / args.pairs
    same_scalar = all(repr(g) == repr(w) for g, w in zip(got, want))

    ctx = Context(0)
    st = Store.from_planes(ctx, pack_bits(h), n_hap)
    st.select_all()
    st.pairs(ia[:1], ib[:1], raw=False)
    t0 = time.perf_counter()
    one = [st.pairs(ia[k:k + 1], ib[k:k + 1], raw=False)["packed"][0] for k in range(args.pairs)]
    t_store = (time.perf_counter() - t0) / args.pairs
    t0 = time.perf_counter()
    res = st.pairs(ia, ib, raw=False)
    t_batch = time.perf_counter() - t0
    n1, p_e4, _ = st.counts()

    def as_dict(word, a, b):
        return {'r_square': r2_value(word), 'd_prime': dprime_value(word), 'var_1_alt_freq': int(p_e4[a]) / 10000.0, 'var_2_alt_freq': int(p_e4[b]) / 10000.0}

    same_store = all(repr(as_dict(w, a, b)) == repr(x) for w, a, b, x in zip(one, ia, ib, want))
    same_batch = all(repr(as_dict(w, a, b)) == repr(x) for w, a, b, x in zip(res["packed"], ia, ib, want))
    print(json.dumps({"workload": f"ld_lite: {args.pairs} variant pairs x {n_hap} haplotypes (BASELINE configs[0])",
                      "reference_port_us_per_pair_one_core": t_ref * 1e6,
                      "scalar_calc_ld_us_per_pair": t_scalar * 1e6, "store_pairs_us_per_pair": t_store * 1e6,
                      "batch_us_total": t_batch * 1e6, "batch_us_per_pair": t_batch * 1e6 / args.pairs,
                      "identical_to_reference": {"scalar": same_scalar, "store": same_store, "batch": same_batch},
                      "r2_ge_0.8_pairs": int(sum(1 for w in want if w['r_square'] >= 0.8))}))
    st.close()
    ctx.close()


if __name__ == "__main__":
    main()
