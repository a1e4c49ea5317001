#!/bin/bash
# one-GPU measurement pass (run under gpurun): bash tools/gpu_call_single.sh TAG
TAG=$1
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; tail -1 $O/${TAG}_pytest_gpu.log
timeout 600 python bench.py > $O/${TAG}_bench.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --workload ld_area --steps 5 --warmup 3 > $O/${TAG}_bench_area.log 2>&1; echo "bench area rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.log 2>&1; echo "bench ref rc=$?"
timeout 600 python tools/bench_sharded.py --variants 100000 > $O/${TAG}_sharded_n1.log 2>&1; echo "sharded rc=$?"
timeout 300 python tools/bench_large.py 2000 8192 32768 --tiles 128 > $O/${TAG}_large.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
B="python bench.py --steps 8 --warmup 3 --no-steady --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv $B > $O/${TAG}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:triangle_mma_kernel -s 4 -c 1 -f -o $O/${TAG}_prof_mma_v2000 $B > $O/${TAG}_ncu2.log 2>&1
for f in bench bench_area bench_ref sharded_n1; do grep "^{" $O/${TAG}_$f.log | cut -c1-700; done; cat $O/${TAG}_large.log
