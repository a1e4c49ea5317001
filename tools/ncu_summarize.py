"""Summarise one kernel of an .ncu-rep (ncu --set full) into the text format kept under profiles/.

    python tools/ncu_summarize.py REPORT.ncu-rep KERNEL_SUBSTRING "header line" > profiles/NAME.txt

Reads the report with `ncu -i REPORT --page raw --csv` and prints the metrics the DESIGN.md rooflines
quote (duration, pipe utilisation, issue slots, DRAM bytes = roofline.traffic, L2 hit rate, launch shape)
for the LAST launch whose name contains KERNEL_SUBSTRING (or, with a fourth argument "longest", the launch of that kernel
with the longest duration -- the full-size one when a script also runs the kernel on small inputs).
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "sm__cycles_elapsed.max",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "lts__t_sector_hit_rate.pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__occupancy_limit_registers",
]


def main():
    rep, kernel, header = sys.argv[1], sys.argv[2], sys.argv[3]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    names, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(names)}
    hit = [r for r in rows[2:] if kernel in r[col["Kernel Name"]]]
    if not hit:
        sys.exit(f"no launch of a kernel matching {kernel!r} in {rep}")
    r = hit[-1]
    if len(sys.argv) > 4 and sys.argv[4] == "longest":
        def dur(x):
            v, u = float(x[col["gpu__time_duration.sum"]]), units[col["gpu__time_duration.sum"]]
            return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u.replace("second", "s").strip(), 1.0)
        r = max(hit, key=dur)
        header += f" [longest of {len(hit)} launches]"
    print("# " + header)
    print(f"{'Kernel Name':<90} {r[col['Kernel Name']]}")
    for m in METRICS:
        if m in col:
            print(f"{m:<90} {r[col[m]]} {units[col[m]]}")


if __name__ == "__main__":
    main()
