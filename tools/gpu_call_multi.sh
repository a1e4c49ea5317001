#!/bin/bash
# multi-GPU measurement pass (run under gpurun --gpus N): bash tools/gpu_call_multi.sh N TAG
N=$1; TAG=$2
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nproc > $O/${TAG}_nproc.log
timeout 900 $T --master-port 29512 tools/bench_area_genome.py > $O/${TAG}_area_genome.log 2>&1; echo "area rc=$?"
timeout 600 $T --master-port 29513 tools/bench_sharded.py --variants 100000 > $O/${TAG}_sharded.log 2>&1; echo "sharded rc=$?"
timeout 600 $T --master-port 29514 bench.py --gpus $N > $O/${TAG}_bench.log 2>&1; echo "bench rc=$?"
timeout 600 $T --master-port 29515 bench.py --gpus $N --workload ld_area --steps 5 --warmup 3 > $O/${TAG}_bench_area.log 2>&1; echo "bench area rc=$?"
for f in area_genome sharded bench bench_area; do grep "^{" $O/${TAG}_$f.log | cut -c1-900; done
