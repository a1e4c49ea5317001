#!/bin/bash
# round 2, call C: the whole GPU test suite, smoke, the new bench.py (N=1) with direct mode on and off
TAG=${1:-r2c}
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $O/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/${TAG}_smoke.log | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.log 2>&1; echo "bench rc=$?"; tail -5 $O/${TAG}_bench.log | cut -c1-6000
LDX_MMA_DIRECT=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-sharded --no-area --no-steady --no-batched --no-cpu-baseline > $O/${TAG}_bench_gather.log 2>&1; echo "bench(gather) rc=$?"; tail -2 $O/${TAG}_bench_gather.log | cut -c1-2500
timeout 300 python tools/bench_modes.py > $O/${TAG}_modes_bench.log 2>&1; cut -c1-330 $O/${TAG}_modes_bench.log
