#!/bin/bash
# round-2 first GPU pass: parity tests, then the frame breakdown of c2 / c4 for the CTA sizes of weight_lists_kernel
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02a_pytest.log
tail -5 $O/r02a_pytest.log
for v in "" l768 l1024; do
  if [ -z "$v" ]; then unset PFT_LIB; tag=l512; else export PFT_LIB=$PWD/pcl_tracking_b200/lib/libpft_$v.so; tag=$v; fi
  timeout 300 python scripts/frame_breakdown.py c2 100 > $O/r02a_breakdown_c2_$tag.txt 2>&1
  timeout 300 python scripts/frame_breakdown.py c4 10 > $O/r02a_breakdown_c4_$tag.txt 2>&1
  head -8 $O/r02a_breakdown_c2_$tag.txt; head -5 $O/r02a_breakdown_c4_$tag.txt
done
unset PFT_LIB
