#!/bin/bash
# window-kernel captures (run under gpurun): bash tools/gpu_call_area.sh TAG
TAG=$1
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
B="python bench.py --workload ld_area --steps 3 --warmup 3"
timeout 600 $B > $O/${TAG}_area_plain.log 2>&1; echo "plain rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${TAG}_area_launches.csv $B > $O/${TAG}_area_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:window_mq_kernel -s 3 -c 1 -f -o $O/${TAG}_prof_window_mq $B > $O/${TAG}_area_ncu2.log 2>&1; echo "ncu rc=$?"
grep "^{" $O/${TAG}_area_plain.log | cut -c1-400
