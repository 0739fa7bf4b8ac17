#!/bin/bash
# per-kernel device times of the c2 frame as a function of the particle count (fixed overheads vs per-query cost)
O=gpurun_out; mkdir -p $O
for n in 125 250 500 1000 2000 4000 8000 16000; do
  timeout 200 python scripts/frame_breakdown.py c2 40 $n > $O/r02_nscan_$n.txt 2>&1
  echo "== N=$n"; head -8 $O/r02_nscan_$n.txt | cut -c1-105
done
