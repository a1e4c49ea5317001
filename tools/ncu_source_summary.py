"""Summarise the source page of an .ncu-rep (ncu --set full --import-source on) of the tcgen05 all-pairs kernel:
warp-stall samples per warp role (the roles are separated by their USETMAXREG / first UTCIMMA instructions in the SASS
listing) and the instructions with the most samples.

    python tools/ncu_source_summary.py REPORT.ncu-rep "header line" > profiles/NAME.txt
"""
import csv
import io
import subprocess
import sys


def main():
    rep, header = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    kernel, hdr, data = rows[0][1], rows[1], rows[2:]
    ix = {n: i for i, n in enumerate(hdr)}
    stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    src = [r[ix["Source"]] for r in data]
    setmax = [i for i, s in enumerate(src) if "USETMAXREG" in s]
    mma = [i for i, s in enumerate(src) if "UTCIMMA" in s]
    # producer | (spare) | wideners | epilogue | MMA issuer, in SASS order
    bounds = [("prologue", 0, setmax[0]), ("producer", setmax[0], setmax[1]), ("wideners", setmax[2], setmax[3]),
              ("epilogue", setmax[3], setmax[4] if len(setmax) > 4 else mma[0] - 150), ("mma issuer + exit", setmax[4] if len(setmax) > 4 else mma[0] - 150, len(data))]
    total = sum(int(r[ix["# Samples"]] or 0) for r in data)
    print("# " + header)
    print(f"# kernel: {kernel}")
    print(f"# {total} warp-stall samples over {len(data)} SASS instructions")
    print()
    print(f"{'role':<20} {'samples':>8} {'share':>6}  {'instr executed':>15}  top stall reasons")
    for name, a, b in bounds:
        tot, agg = 0, {}
        for r in data[a:b]:
            tot += int(r[ix["# Samples"]] or 0)
            for n in stalls:
                agg[n] = agg.get(n, 0) + int(r[ix[n]] or 0)
        ex = sum(int(r[ix["Instructions Executed"]] or 0) for r in data[a:b])
        top = ", ".join(f"{n[6:]} {v}" for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6])
        print(f"{name:<20} {tot:>8} {100 * tot / max(total, 1):>5.1f}%  {ex:>15}  {top}")
    print()
    print("instructions with the most samples:")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:30]:
        s = int(r[ix["# Samples"]] or 0)
        top = ", ".join(f"{n[6:]} {int(r[ix[n]] or 0)}" for n in sorted(stalls, key=lambda n: -int(r[ix[n]] or 0))[:2])
        print(f"  {r[ix['Address']][-5:]} {s:>6} {100 * s / max(total, 1):>5.1f}%  exec {r[ix['Instructions Executed']]:>10}  {r[ix['Source']].strip()[:70]:<70} {top}")


if __name__ == "__main__":
    main()
