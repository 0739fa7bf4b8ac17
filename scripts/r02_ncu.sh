#!/bin/bash
# one `ncu --set full` capture of the hot kernels of a steady-state frame (c2, and the weight kernel of c4)
TAG=${1:-r02a}
O=gpurun_out
mkdir -p $O
ncu --set full --import-source on --clock-control none -k regex:"weight_lists_kernel|cand_build_kernel|cand_octant_kernel|cand_build_far_kernel|cand_mark_kernel" \
    --launch-skip 60 --launch-count 10 -o $O/${TAG}_c2_hot -f python scripts/frame_breakdown.py c2 14 > $O/${TAG}_ncu_c2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"weight_lists_kernel" \
    --launch-skip 6 --launch-count 1 -o $O/${TAG}_c4_hot -f python scripts/frame_breakdown.py c4 5 > $O/${TAG}_ncu_c4.log 2>&1
ls -la $O | grep $TAG
