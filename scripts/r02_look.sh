#!/bin/bash
# tuning: per-occurrence frame breakdown, list statistics and one full ncu capture of the hot kernels of c2
TAG=${1:-r02u}
O=gpurun_out
mkdir -p $O
timeout 300 python scripts/frame_breakdown.py c2 100 > $O/${TAG}_breakdown_c2.txt 2>&1
head -22 $O/${TAG}_breakdown_c2.txt | cut -c1-400
PFT_LIB=$PWD/pcl_tracking_b200/lib/libpft_stats.so timeout 300 python scripts/stats_run.py > $O/${TAG}_stats.txt 2>&1
cat $O/${TAG}_stats.txt
if [ "${NCU:-1}" = 1 ]; then bash scripts/r02_ncu.sh $TAG; fi
