#!/bin/bash
# Collects the evidence committed under profiles/ (run on the GPU box: gpurun -- 'bash scripts/collect_profiles.sh TAG').
# Bench lines first (never under a profiler), then the ncu launch list and one `--set full` capture of the hot kernels.
TAG=${1:-r01f}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${TAG}_bench_c2_1gpu.json 2> $O/${TAG}_bench_c2_1gpu.err
python bench.py --impl reference --steps 10 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_reference_arm.err
for w in c3 c4 c5; do
  python bench.py --workload $w --steps 60 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_${w}_1gpu.json 2> $O/${TAG}_bench_${w}_1gpu.err
done
python scripts/frame_breakdown.py c2 100 > $O/${TAG}_frame_breakdown_c2.txt 2>&1
python scripts/frame_breakdown.py c4 10 > $O/${TAG}_frame_breakdown_c4.txt 2>&1
# launch list of the bench command (cold caches, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_launches.log 2>&1
# one full capture of each hot kernel of a steady-state frame
ncu --set full --import-source on --clock-control none -k regex:"weight_kernel|cand_build|cand_mark|cand_collect|aabb_kernel|k1_insert|k1_accum" \
    --launch-skip 60 --launch-count 9 -o $O/${TAG}_hot_kernels -f python scripts/frame_breakdown.py c2 12 > $O/${TAG}_ncu_full.log 2>&1
ls -la $O | grep $TAG
