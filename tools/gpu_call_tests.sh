#!/bin/bash
# the GPU test suite (optionally a -k expression) and smoke
TAG=${1:-t}; shift
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest ${@:-tests} -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $O/${TAG}_pytest_gpu.log
