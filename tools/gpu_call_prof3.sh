#!/bin/bash
# ncu --set full: the store kernels at full size (longest launch of each), and the row-per-thread window kernel at configs[2]
TAG=${1:-r2s}
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
S="python tools/ncu_summarize.py"
P="python tools/bench_store.py --variants 200000"
$P > $O/${TAG}_store_plain.json 2> $O/${TAG}_store_plain.err && \
  ncu --set full --clock-control none -k regex:'pack_gt_kernel|variant_freq_kernel|pairs_kernel|parse_lines_kernel|subset_kernel' \
      -c 64 -f -o /tmp/${TAG}_store $P > $O/${TAG}_store_ncu.log 2>&1
echo "store rc=$?"
for k in pack_gt_kernel variant_freq_kernel pairs_kernel parse_lines_kernel subset_kernel; do
  $S /tmp/${TAG}_store.ncu-rep $k "ncu --set full, $k, tools/bench_store.py --variants 200000 (5008 haplotypes; K1 on 50,000 rows of GT text, ingest on 20,000 records of 2504 samples)" longest > $O/${TAG}_ncu_full_$k.txt 2>> $O/${TAG}_summ.err
done
P="python bench.py --steps 3 --warmup 3 --no-batched --no-sharded --no-steady"
$P > $O/${TAG}_area_plain.json 2> $O/${TAG}_area_plain.err && \
  ncu --set full --clock-control none --import-source on -k regex:window_rows1 -s 2 -c 1 -f -o $O/${TAG}_window_rows1 $P > $O/${TAG}_area_ncu.log 2>&1
echo "area rc=$?"
$S $O/${TAG}_window_rows1.ncu-rep window_rows1 "ncu --set full, window_rows1_kernel (configs[2]: 100 k variants, 1006 of 5008 haplotypes gathered into 16-word rows, 5,000 queries, +/-500 kb, r2 >= 0.8)" > $O/${TAG}_ncu_full_window_rows1.txt 2>> $O/${TAG}_summ.err
cat $O/${TAG}_summ.err
