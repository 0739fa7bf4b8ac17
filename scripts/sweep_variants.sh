#!/bin/bash
# kernel-tuning sweep: index cell level (auto / fixed) on the c2 bench workload
for lvl in ${LEVELS:-auto 1 2}; do
  if [ "$lvl" = auto ]; then unset PFT_INDEX_LEVEL; else export PFT_INDEX_LEVEL=$lvl; fi
  python bench.py --steps ${STEPS:-100} --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('level $lvl', 'ms/step %.3f'%d['ms_per_step'], 'e2e %.3f'%d['e2e']['ms_per_step'], 'weight ms %.4f'%r['ms_per_launch'], 'kernel Gevals/s %.2f'%(r['evals_per_s_in_kernel']/1e9), 'share %.2f'%r['share_of_compute'], 'clk', d['clocks']['sm_mhz'], d['scene_index'])"
done
