#!/bin/bash
# one-GPU measurement pass (run under gpurun): genome-scale ld_area tool, bench line, launch list, ncu captures, trace
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 300 python tools/bench_area_genome.py --variants 2000000 --queries 2000 --reps 2 --check 4 > $O/r01v_area_small.log 2>&1
echo "area small rc=$?"
timeout 900 python tools/bench_area_genome.py > $O/r01v_area_genome_n1.log 2>&1
echo "area genome rc=$?"
timeout 600 python bench.py > $O/r01v_bench.log 2>&1
echo "bench rc=$?"
timeout 300 python tools/bench_large.py 2000 --tiles 128 --trace > $O/r01v_trace.log 2>&1
timeout 300 python tools/bench_large.py 2000 8192 32768 --tiles 128 > $O/r01v_large.log 2>&1
B="python bench.py --steps 8 --warmup 3 --no-steady --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01v_launches.csv $B > $O/r01v_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:triangle_mma_kernel -s 4 -c 1 -f -o $O/r01v_prof_mma_v2000 $B > $O/r01v_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:triangle_mma_kernel -s 1 -c 1 -f -o $O/r01v_prof_mma_v32768 python tools/bench_large.py 32768 --tiles 128 --reps 1 > $O/r01v_ncu3.log 2>&1
tail -2 $O/r01v_area_small.log | cut -c1-1500
tail -1 $O/r01v_area_genome_n1.log | cut -c1-1800
cat $O/r01v_large.log
