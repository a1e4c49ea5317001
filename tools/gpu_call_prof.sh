#!/bin/bash
# ncu --set full of the all-pairs kernel: one-wave direct, one-wave gathered, multi-wave pair (32,768 variants)
TAG=${1:-r2p}
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
P="python tools/prof_target.py"
$P 2000 --direct 1 > $O/${TAG}_plain1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:triangle_mma -s 2 -c 1 -f -o $O/${TAG}_direct_v2000 $P 2000 --direct 1 > $O/${TAG}_ncu1.log 2>&1; echo "rc=$?"
$P 2000 --direct 0 > $O/${TAG}_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:triangle_mma -s 2 -c 1 -f -o $O/${TAG}_gather_v2000 $P 2000 --direct 0 > $O/${TAG}_ncu2.log 2>&1; echo "rc=$?"
$P 16384 > $O/${TAG}_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:triangle_mma -s 2 -c 1 -f -o $O/${TAG}_pair_v16384 $P 16384 > $O/${TAG}_ncu3.log 2>&1; echo "rc=$?"
ls -la $O/*.ncu-rep
