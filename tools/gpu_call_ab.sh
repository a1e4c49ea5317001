#!/bin/bash
# A/B timing of library builds: bash tools/gpu_call_ab.sh TAG A B C ...  (ld_tools_b200/lib/libldx_<X>.so)
TAG=$1; shift
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
for v in "$@"; do
  echo "== variant $v"
  LDX_LIB=$PWD/ld_tools_b200/lib/libldx_$v.so timeout 300 python tools/bench_modes.py --batches 16 --reps 12 > $O/${TAG}_ab_$v.log 2>&1 || tail -3 $O/${TAG}_ab_$v.log
  python - "$O/${TAG}_ab_$v.log" <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line)
        k=d.get('allpairs_kernel_us') or d.get('allpairs_kernel_ms')
        print(' ', d['case'], d.get('variants'), d.get('direct', d.get('sets','')), 'kernel', round(k,3), 'call', round(d.get('call_us_median') or d.get('call_ms_median'),3), 'frac', round(d['roofline_frac_kernel'],3))
PY
done
