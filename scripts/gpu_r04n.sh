#!/usr/bin/env bash
# compute-sanitizer memcheck over smoke() and one launch of each of the four bench calls
set -u
OUT=gpurun_out/r04n; mkdir -p $OUT
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/memcheck_smoke.log 2>&1; echo "smoke memcheck exit $?"; tail -4 $OUT/memcheck_smoke.log
for w in sceneflow_576x960 kitti_384x1280; do
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/one_kernel.py corr_fwd,corr_bwd,warp_fwd,warp_bwd $w 2 > $OUT/memcheck_$w.log 2>&1; echo "$w memcheck exit $?"; tail -3 $OUT/memcheck_$w.log
done
