#!/bin/bash
# final ncu --set full pass of round 2: the all-pairs kernel (one-wave direct, multi-wave pair), the row-per-thread window kernel,
# the subset kernel; every ncu run follows a plain run of the same command that exited 0
TAG=${1:-r2f}
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
S="python tools/ncu_summarize.py"
P="python tools/prof_target.py"
$P 2000 --direct 1 > $O/${TAG}_plain1.log 2>&1 && ncu --set full --clock-control none -k regex:triangle_mma -s 2 -c 1 -f -o /tmp/${TAG}_direct $P 2000 --direct 1 > $O/${TAG}_ncu1.log 2>&1; echo "rc=$?"
$S /tmp/${TAG}_direct.ncu-rep triangle_mma "ncu --set full, tcgen05 all-pairs kernel, one wave, DIRECT mode, 2,000 variants x 5008 haplotypes (round 2 final: streaming result stores)" > $O/${TAG}_ncu_full_mma_direct_v2000.txt
$P 16384 > $O/${TAG}_plain3.log 2>&1 && ncu --set full --clock-control none -k regex:triangle_mma -s 2 -c 1 -f -o /tmp/${TAG}_pair $P 16384 > $O/${TAG}_ncu3.log 2>&1; echo "rc=$?"
$S /tmp/${TAG}_pair.ncu-rep triangle_mma "ncu --set full, tcgen05 all-pairs kernel, multi-wave CTA pairs, 16,384 variants x 5008 haplotypes (round 2 final: streaming result stores)" > $O/${TAG}_ncu_full_mma_pair_v16384.txt
P="python bench.py --workload ld_area --steps 3 --warmup 3"
$P > $O/${TAG}_area_plain.json 2> $O/${TAG}_area_plain.err && ncu --set full --clock-control none -k regex:'window_rows1|subset_rows' -c 8 -f -o /tmp/${TAG}_area $P > $O/${TAG}_area_ncu.log 2>&1; echo "rc=$?"
$S /tmp/${TAG}_area.ncu-rep window_rows1 "ncu --set full, window_rows1_kernel (round 2 final), configs[2]: 1.1 M variants, 1006 of 5008 haplotypes gathered into 16-word rows, 1,000 queries, +/-500 kb, r2 >= 0.8" longest > $O/${TAG}_ncu_full_window_rows1.txt
$S /tmp/${TAG}_area.ncu-rep subset_rows "ncu --set full, subset_rows_kernel: 1.1 M rows, 1006 of 5008 columns" longest > $O/${TAG}_ncu_full_subset_rows.txt
ls -la $O/${TAG}_*.txt
