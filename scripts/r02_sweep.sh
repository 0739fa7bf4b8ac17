#!/bin/bash
# frame breakdown of c2 / c4 for a list of tuning builds of libpft (pcl_tracking_b200/lib/libpft_<tag>.so); "default" = libpft.so
TAG=${TAG:-r02b}
O=gpurun_out
mkdir -p $O
if [ "${PYTEST:-1}" = 1 ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1
  echo "pytest rc=$?" >> $O/${TAG}_pytest.log
  tail -4 $O/${TAG}_pytest.log
fi
for v in "$@"; do
  if [ "$v" = default ]; then unset PFT_LIB; else export PFT_LIB=$PWD/pcl_tracking_b200/lib/libpft_$v.so; fi
  timeout 300 python scripts/frame_breakdown.py c2 100 > $O/${TAG}_breakdown_c2_$v.txt 2>&1
  timeout 300 python scripts/frame_breakdown.py c4 10 > $O/${TAG}_breakdown_c4_$v.txt 2>&1
  echo "== $v"; head -4 $O/${TAG}_breakdown_c2_$v.txt | cut -c1-110; head -4 $O/${TAG}_breakdown_c4_$v.txt | cut -c1-110
done
unset PFT_LIB
