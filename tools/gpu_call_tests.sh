#!/bin/bash
# the GPU test suite (optionally files and a -k expression) : bash tools/gpu_call_tests.sh TAG [pytest args]
TAG=${1:-t}; shift
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
if [ $# -eq 0 ]; then set -- tests; fi
timeout 1500 python -m pytest "$@" -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $O/${TAG}_pytest_gpu.log
