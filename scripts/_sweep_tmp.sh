for w in corr_micro sceneflow_576x960 kitti_384x1280; do
for cb in 0 8 12 16; do
  echo "== $w CB=$cb"; DSSMD_CORR_BWD_CB=$cb DSSMD_CORR_DEBUG=1 python scripts/kbench.py corr_bwd_gL $w X=1 2>&1 | grep -E "cfg:|X=" | sort -u | head -3
done; done
