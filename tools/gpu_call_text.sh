#!/bin/bash
# one-GPU pass for the matrix text kernels (run under gpurun): bash tools/gpu_call_text.sh TAG
TAG=$1
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} > $O/${TAG}_pytest_gpu.log 2>&1; tail -3 $O/${TAG}_pytest_gpu.log
timeout 600 python tools/bench_text.py 2000 20000 > $O/${TAG}_text.log 2>&1; echo "bench_text rc=$?"
B="python tools/bench_text.py 2000"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${TAG}_text_launches.csv $B > $O/${TAG}_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:matrix_text_kernel -s 2 -c 1 -f -o $O/${TAG}_prof_matrix_text_v2000 $B > $O/${TAG}_ncu2.log 2>&1
cat $O/${TAG}_text.log | cut -c1-900
