#!/bin/bash
# kernel-tuning sweep: weight-kernel CTA size x index cell level on the c2 bench workload
for v in "" _t768 _t512; do
 for lvl in ${LEVELS:-1 2}; do
  if [ -z "$v" ]; then unset PFT_LIB; name=t1024; else export PFT_LIB=$PWD/pcl_tracking_b200/lib/libpft$v.so; name=$v; fi
  PFT_INDEX_LEVEL=$lvl python bench.py --steps 60 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('$name level $lvl', 'ms/step %.3f'%d['ms_per_step'], 'weight ms %.4f'%r['ms_per_launch'], 'kernel Gevals/s %.2f'%(r['evals_per_s_in_kernel']/1e9), 'share %.2f'%r['share_of_compute'], d['scene_index'])"
 done
done
