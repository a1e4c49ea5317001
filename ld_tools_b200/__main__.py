"""`python -m ld_tools_b200 {ld_area,ld_triangle,ld_lite} ...` -- the reference's command lines on the GPU engine.

Option letters, long names, defaults and meanings are those of cli/ld_area_cli_en.py:36-60,
cli/ld_triangle_cli_en.py:40-74 and cli/ld_lite_cli_en.py:37-49, so an existing invocation keeps working:

    python ld_area.py -S src -D intgen -f -w 500000 -z 0.8 -e eur      (reference)
    python -m ld_tools_b200 ld_area -S src -D intgen -f -w 500000 -z 0.8 -e eur

-f (skip the download/verification of 1000 Genomes data) is accepted and implied: the network path of
prep_intgen_data.py is out of scope (the upstream data is gone, reference README.md:1-2).  -p, the reference's
number of parallel processes over source files, is the number of GPUs to fan the job's tables / matrices out to
(at most those present; one worker thread and one context per GPU).  ld_triangle writes the table output
(-o table|both); the Plotly heatmap is out of scope.
"""
import argparse
import datetime
import sys

from . import drivers


def common(ap, with_src=True):
    if with_src:
        ap.add_argument("-S", "--src-dir-path", dest="src_dir_path", required=True)
    ap.add_argument("-D", "--intgen-dir-path", dest="intgen_dir_path", required=True)
    if with_src:
        ap.add_argument("-t", "--trg-top-dir-path", dest="trg_top_dir_path", default=None)
        ap.add_argument("-m", "--meta-lines-quan", dest="meta_lines_quan", type=int, default=0)
        ap.add_argument("-p", "--max-proc-quan", dest="max_proc_quan", type=int, default=4)
    ap.add_argument("-f", "--skip-intgen-data-ver", dest="skip_intgen_data_ver", action="store_true")
    ap.add_argument("-g", "--gend-names", dest="gend_names", choices=["male", "female", "both"], default="both")
    ap.add_argument("-e", "--pop-names", dest="pop_names", default="all")


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m ld_tools_b200", description=__doc__,
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="tool", required=True)
    a = sub.add_parser("ld_area")
    common(a)
    a.add_argument("-w", "--flank-size", dest="flank_size", type=int, default=100000)
    a.add_argument("-l", "--ld-thres-measure", dest="ld_thres_measure", choices=["r_square", "d_prime"], default="r_square")
    a.add_argument("-z", "--ld-low-thres", dest="ld_low_thres", type=float, default=0.8)
    a.add_argument("-o", "--trg-file-type", dest="trg_file_type", choices=["tsv", "json", "rsids"], default="tsv")
    t = sub.add_parser("ld_triangle")
    common(t)
    t.add_argument("-l", "--ld-measure", dest="ld_measure", choices=["r_square", "d_prime"], default="r_square")
    t.add_argument("-z", "--ld-low-thres", dest="ld_low_thres", type=float, default=None)
    t.add_argument("-o", "--matrix-type", dest="matrix_type", choices=["heatmap", "table", "both"], default="table")
    li = sub.add_parser("ld_lite")
    li.add_argument("rs_id_1")
    li.add_argument("rs_id_2")
    common(li, with_src=False)
    args = ap.parse_args(argv)
    t0 = datetime.datetime.now()
    devices = None
    if args.tool != "ld_lite":
        from .engine import device_count
        n = min(max(args.max_proc_quan, 1), device_count())
        devices = n if n > 1 else None
    if args.tool == "ld_area":
        drivers.ld_area(args.src_dir_path, args.intgen_dir_path, args.trg_top_dir_path, args.meta_lines_quan, args.gend_names,
                        args.pop_names, args.flank_size, args.ld_thres_measure, args.ld_low_thres, args.trg_file_type, devices=devices)
    elif args.tool == "ld_triangle":
        if args.matrix_type == "heatmap":
            sys.exit("the Plotly heatmap output is out of scope of this engine: use -o table")
        drivers.ld_triangle(args.src_dir_path, args.intgen_dir_path, args.trg_top_dir_path, args.meta_lines_quan, args.gend_names,
                            args.pop_names, args.ld_measure, args.ld_low_thres, devices=devices)
    else:
        print(drivers.ld_lite(args.rs_id_1, args.rs_id_2, args.intgen_dir_path, args.gend_names, args.pop_names))
        return
    print(f"\tcomputation time: {datetime.datetime.now() - t0}")


if __name__ == "__main__":
    main()
