#!/bin/bash
# round 2, call A: the rewritten tcgen05 engine -- new mode tests first, then the old triangle tests, smoke, mode timings
TAG=${1:-r2a}
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests/test_triangle_modes_gpu.py -x -q -m gpu > $O/${TAG}_modes.log 2>&1; echo "modes rc=$?"; tail -15 $O/${TAG}_modes.log
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "triangle or resolve or timing" > $O/${TAG}_parity_tri.log 2>&1; echo "parity-tri rc=$?"; tail -8 $O/${TAG}_parity_tri.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/${TAG}_smoke.log
timeout 600 python tools/bench_modes.py > $O/${TAG}_modes_bench.log 2>&1; echo "bench_modes rc=$?"; cat $O/${TAG}_modes_bench.log | cut -c1-400
timeout 300 python tools/bench_large.py 2000 --tiles 128 --trace > $O/${TAG}_trace2000.log 2>&1; head -30 $O/${TAG}_trace2000.log
