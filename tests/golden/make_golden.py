#!/usr/bin/env python3
"""Generate tests/golden/calc_ld_golden.json from the UNMODIFIED reference calc_ld.

Run in the dev container (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference (`/root/reference/backend/calc_ld.py`) is imported by path and never edited.
Three things are recorded per case, all produced by the reference's own code:

  * the dict it returns (values stored as repr() strings so int `0` vs float `0.0` survives);
  * the same four values BEFORE round(., 4): obtained by running the same function object with
    a module-global `round` that returns its argument (name lookup finds the global before the
    builtin; the source file is untouched);
  * D itself (a local of calc_ld, `calc_ld.py:50`), read from the frame with a profile hook.

Floats are stored as float.hex() so the fixture is bit-exact.
"""
import itertools
import json
import os
import platform
import random
import sys
from fractions import Fraction

sys.dont_write_bytecode = True
REF = os.environ.get("LD_TOOLS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from backend import calc_ld as ref_mod   # noqa: E402

ref_calc_ld = ref_mod.calc_ld
HERE = os.path.dirname(os.path.abspath(__file__))


def run_reference(g1, g2):
    """-> (rounded dict, raw dict incl. 'd')."""
    rounded = ref_calc_ld(g1, g2)
    captured = {}

    def hook(frame, event, arg):
        if event == "return" and frame.f_code is ref_calc_ld.__code__:
            captured["d"] = frame.f_locals["d"]

    ref_mod.round = lambda x, nd: x          # shadows the builtin for this module only
    sys.setprofile(hook)
    try:
        raw = ref_calc_ld(g1, g2)
    finally:
        sys.setprofile(None)
        del ref_mod.round
    raw = dict(raw)
    raw["d"] = captured["d"]
    return rounded, raw


KEYS = ("r_square", "d_prime", "var_1_alt_freq", "var_2_alt_freq")


def enc_raw(v):
    """int 0 (the reference's ZeroDivisionError / D'==0 sentinels) stays the string "0"."""
    if isinstance(v, bool) or not isinstance(v, (int, float)):
        raise TypeError(v)
    return "0" if isinstance(v, int) else v.hex()


def enc_out(rounded, raw):
    """[4 x repr(rounded value)] + [4 x raw value] + [raw D]; repr keeps `0` vs `0.0` apart."""
    return [repr(rounded[k]) for k in KEYS] + [enc_raw(raw[k]) for k in KEYS] + [enc_raw(raw["d"])]


CHARS = {0: "0", 1: "1", 2: "2", None: "N"}


def enc_genotypes(g):
    return "".join(CHARS[x] for x in g)


def lists_from_counts(n, n11, n1a, n1b):
    """A genotype arrangement with the requested counts (order is irrelevant to calc_ld)."""
    only_a, only_b = n1a - n11, n1b - n11
    rest = n - n11 - only_a - only_b
    assert min(only_a, only_b, rest) >= 0
    g1 = [1] * n11 + [1] * only_a + [0] * only_b + [0] * rest
    g2 = [1] * n11 + [0] * only_a + [1] * only_b + [0] * rest
    return g1, g2


def exact_r2(n, n11, a, b):
    den = a * (n - a) * b * (n - b)
    return None if den == 0 else Fraction((n11 * n - a * b) ** 2, den)


def exact_dprime(n, n11, a, b):
    num = n11 * n - a * b
    den = min(a * (n - b), (n - a) * b) if num >= 0 else max(-a * b, -(n - a) * (n - b))
    return None if den == 0 else Fraction(num, den)


def is_tie(fr):
    return fr is not None and (fr * 20000).denominator == 1 and (fr * 20000).numerator % 2 == 1


def main():
    rng = random.Random(20130502)
    list_cases, count_cases = [], []

    def add_list(name, g1, g2):
        rounded, raw = run_reference(g1, g2)
        list_cases.append([name, enc_genotypes(g1), enc_genotypes(g2)] + enc_out(rounded, raw))

    def add_counts(tag, n, n11, a, b):
        rounded, raw = run_reference(*lists_from_counts(n, n11, a, b))
        count_cases.append([tag, n, n11, a, b] + enc_out(rounded, raw))

    # ---- list-level cases: the shapes and quirks SURVEY.md section 8(a)/(c) lists
    add_list("identical_4", [1, 1, 0, 0], [1, 1, 0, 0])
    add_list("complementary_4", [1, 1, 0, 0], [0, 0, 1, 1])
    add_list("mono_ref_a", [0, 0, 0, 0], [1, 0, 1, 0])
    add_list("mono_alt_a", [1, 1, 1, 1], [1, 0, 1, 0])
    add_list("mono_both", [0, 0, 0, 0], [1, 1, 1, 1])
    add_list("d_zero_polymorphic", [1, 1, 0, 0], [1, 0, 1, 0])
    add_list("singletons_same", [1] + [0] * 7, [1] + [0] * 7)
    add_list("singletons_diff", [1] + [0] * 7, [0, 1] + [0] * 6)
    add_list("pair_of_2", [1, 0], [1, 0])
    add_list("tuple_inputs", (1, 0, 1, 1, 0, 0), (1, 0, 0, 1, 0, 1))
    add_list("r2_tie_1_over_32", *lists_from_counts(64, 20, 32, 32))
    add_list("unequal_len_a_longer", [1, 0, 1, 1, 0, 1, 1, 1], [1, 0, 0, 1, 0])
    add_list("unequal_len_b_longer", [1, 0, 0, 1], [1, 1, 0, 1, 0, 0, 1])
    add_list("missing_none", [1, None, 0, 1, 0, 1, None, 0], [1, 1, 0, None, 0, 1, 0, 0])
    add_list("allele_2_present", [1, 2, 0, 1, 0, 1, 2, 0], [1, 1, 0, 2, 0, 1, 0, 0])
    for n in (2, 6, 64, 65, 127, 128, 198, 1006, 5008):
        for rep in range(6 if n >= 1000 else 10):
            fa, fb = rng.choice([0.01, 0.05, 0.2, 0.5, 0.9]), rng.choice([0.01, 0.05, 0.3, 0.5])
            g1 = [int(rng.random() < fa) for _ in range(n)]
            link = rng.random()
            g2 = [x if rng.random() < link else int(rng.random() < fb) for x in g1]
            add_list(f"random_n{n}_{rep}", g1, g2)
    for rep in range(6):   # rare x rare at full 1000G width
        g1 = [0] * 5008
        g2 = [0] * 5008
        for i in rng.sample(range(5008), rep + 1):
            g1[i] = 1
        for i in rng.sample(range(5008), 2 * rep + 1):
            g2[i] = 1
        if rep % 2:
            for i in range(5008):
                g2[i] |= g1[i]
        add_list(f"rare_n5008_{rep}", g1, g2)

    # ---- count-level cases (lists are synthesised from the counts, then run through the reference)
    for n in (8, 16, 32, 64, 100, 128):   # every exact rounding tie for small N
        ties = 0
        for a, b in itertools.product(range(1, n), repeat=2):
            for n11 in range(max(0, a + b - n), min(a, b) + 1):
                if is_tie(exact_r2(n, n11, a, b)) or is_tie(exact_dprime(n, n11, a, b)):
                    if ties < 150 or rng.random() < 0.02:
                        add_counts("exact_tie", n, n11, a, b)
                    ties += 1
    for n in (32, 64, 128, 160, 320, 640, 5008):   # alt-freq ties such as 1/32 = 0.03125
        for a in range(1, n):
            if is_tie(Fraction(a, n)):
                b = max(1, n // 3)
                add_counts("freq_tie", n, max(0, a + b - n, min(a, b) // 2), a, b)
    for n in (2, 3, 10, 198, 1006, 5008):
        for _ in range(700 if n >= 198 else 60):
            a, b = rng.randint(0, n), rng.randint(0, n)
            if rng.random() < 0.4:     # rare-variant corner of the spectrum
                a = min(n, rng.randint(0, 12))
            if rng.random() < 0.3:
                b = min(n, rng.randint(0, 12))
            n11 = rng.randint(max(0, a + b - n), min(a, b))
            if rng.random() < 0.25:
                n11 = min(a, b)        # complete LD: D' = 1
            add_counts("random", n, n11, a, b)
    for n in (1006, 5008):             # threshold neighbourhood of r2 = 0.8 (ld_area default -z 0.8)
        found = 0
        while found < 150:
            a, b = rng.randint(1, n - 1), rng.randint(1, n - 1)
            lo, hi = max(0, a + b - n), min(a, b)
            n11 = rng.randint(lo, hi)
            r2 = exact_r2(n, n11, a, b)
            if r2 is not None and abs(r2 - Fraction(4, 5)) < Fraction(1, 400):
                add_counts("near_0.8", n, n11, a, b)
                found += 1

    doc = {"meta": {"generator": "tests/golden/make_golden.py", "seed": 20130502,
                    "reference": "PlatonB/ld-tools backend/calc_ld.py " + ref_mod.__version__,
                    "python": platform.python_version(), "libc": " ".join(platform.libc_ver()),
                    "encoding": "genotype strings: 0,1,2 = ints, N = None; raw floats as "
                                "float.hex(), int sentinel 0 as \"0\"; rounded values as repr()",
                    "out_columns": ["rounded:" + k for k in KEYS] + ["raw:" + k for k in KEYS]
                    + ["raw:d"],
                    "list_case_columns": ["name", "var_1_genotypes", "var_2_genotypes", "*out"],
                    "count_case_columns": ["tag", "n_hap", "n_11", "n_a1", "n_b1", "*out"]},
           "list_cases": list_cases, "count_cases": count_cases}
    out = os.path.join(HERE, "calc_ld_golden.json")
    with open(out, "w") as fh:
        json.dump(doc, fh, separators=(",", ":"))
    print(f"{len(list_cases)} list cases, {len(count_cases)} count cases -> {out} "
          f"({os.path.getsize(out) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
