#!/bin/bash
# ncu --set full of the store-side kernels (K1 pack_gt, K2 variant_freq, K3 pairs, the VCF line parser), the table writer,
# the window kernel on the subset store (configs[2]) and a 16-set batch of the all-pairs kernel; then the launch list of the
# default bench command.  Every ncu run follows a plain run of the same command that exited 0.  Summaries are made on the box
# (tools/ncu_summarize.py); only those and the small reports come back.
TAG=${1:-r2q}
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
S="python tools/ncu_summarize.py"
P="python tools/bench_store.py --variants 200000"
$P > $O/${TAG}_store_plain.json 2> $O/${TAG}_store_plain.err && \
  ncu --set full --clock-control none -k regex:'pack_gt_kernel|variant_freq_kernel|pairs_kernel|parse_lines_kernel|count_newlines_kernel|compact_rows_kernel|subset_kernel' \
      -c 64 -f -o /tmp/${TAG}_store $P > $O/${TAG}_store_ncu.log 2>&1
echo "store rc=$?"
for k in pack_gt_kernel variant_freq_kernel pairs_kernel parse_lines_kernel count_newlines_kernel compact_rows_kernel subset_kernel; do
  $S /tmp/${TAG}_store.ncu-rep $k "ncu --set full, $k, tools/bench_store.py --variants 200000 (5008 haplotypes; K1/ingest on 50,000 / 20,000 records of 2504 samples)" > $O/${TAG}_ncu_full_$k.txt 2>> $O/${TAG}_summ.err
done
P="python tools/bench_text.py 2000 20000"
$P > $O/${TAG}_text_plain.json 2> $O/${TAG}_text_plain.err && \
  ncu --set full --clock-control none -k regex:'matrix_text_kernel|matrix_line_bytes_kernel' -c 12 -f -o /tmp/${TAG}_text $P > $O/${TAG}_text_ncu.log 2>&1
echo "text rc=$?"
for k in matrix_text_kernel matrix_line_bytes_kernel; do
  $S /tmp/${TAG}_text.ncu-rep $k "ncu --set full, $k, tools/bench_text.py 2000 20000 (last launch: 20,000 variants)" > $O/${TAG}_ncu_full_$k.txt 2>> $O/${TAG}_summ.err
done
P="python bench.py --steps 3 --warmup 3 --no-batched --no-sharded --no-steady"
$P > $O/${TAG}_area_plain.json 2> $O/${TAG}_area_plain.err && \
  ncu --set full --clock-control none --import-source on -k regex:window_mq -s 2 -c 1 -f -o $O/${TAG}_window_mq_subset $P > $O/${TAG}_area_ncu.log 2>&1
echo "area rc=$?"
$S $O/${TAG}_window_mq_subset.ncu-rep window_mq "ncu --set full, window_mq_kernel on the subset store (configs[2]: 100 k variants, 1006 of 5008 haplotypes gathered into 16-word rows, 5,000 queries)" > $O/${TAG}_ncu_full_window_mq_subset.txt 2>> $O/${TAG}_summ.err
P="python tools/prof_target.py 2000 --batch 16"
$P > $O/${TAG}_batch_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:triangle_mma -s 2 -c 1 -f -o $O/${TAG}_mma_batch16 $P > $O/${TAG}_batch_ncu.log 2>&1
echo "batch rc=$?"
$S $O/${TAG}_mma_batch16.ncu-rep triangle_mma "ncu --set full, tcgen05 all-pairs kernel, ONE launch over 16 sets of 2,000 variants x 5008 haplotypes (ldx_triangle_batch_dev)" > $O/${TAG}_ncu_full_mma_batch16.txt 2>> $O/${TAG}_summ.err
P="python bench.py --steps 5 --warmup 3"
$P > $O/${TAG}_bench_plain.json 2> $O/${TAG}_bench_plain.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${TAG}_launches_bench.csv $P > $O/${TAG}_bench_ncu.log 2>&1
echo "launches rc=$?"
ls -la $O | tail -40
