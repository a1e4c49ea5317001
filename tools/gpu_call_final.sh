#!/bin/bash
# whole one-GPU pass (run under gpurun): bash tools/gpu_call_final.sh TAG -- tests, benches, smoke, launch lists, ncu captures
TAG=$1
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; tail -3 $O/${TAG}_pytest_gpu.log
timeout 600 python bench.py > $O/${TAG}_bench.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --workload ld_area --steps 5 --warmup 3 > $O/${TAG}_bench_area.log 2>&1; echo "bench area rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.log 2>&1; echo "bench ref rc=$?"
timeout 300 python tools/bench_large.py 2000 8192 32768 --tiles 128 > $O/${TAG}_large.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
timeout 600 python tools/bench_text.py 2000 20000 > $O/${TAG}_text.log 2>&1; echo "bench_text rc=$?"
B="python bench.py --steps 8 --warmup 3 --no-steady --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv $B > $O/${TAG}_ncu1.log 2>&1
T="python tools/bench_text.py 20000"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:matrix_ -c 40 --csv --log-file $O/${TAG}_text_launches.csv $T > $O/${TAG}_ncu3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:matrix_text_kernel -s 2 -c 1 -f -o $O/${TAG}_prof_matrix_text_v8000 python tools/bench_text.py 8000 > $O/${TAG}_ncu4.log 2>&1
for f in bench bench_area bench_ref; do grep "^{" $O/${TAG}_$f.log | cut -c1-600; done; cat $O/${TAG}_large.log; cut -c1-900 $O/${TAG}_text.log
