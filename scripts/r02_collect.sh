#!/bin/bash
# Final evidence of round 2 on one B200: parity tests, smoke, the default bench line, the reference arm, frame breakdowns,
# the ncu launch list of the bench command and one `ncu --set full` capture of the hot kernels (c2 and c4).
# Everything lands in gpurun_out/r02f_*; the summaries are copied into profiles/ afterwards.
O=gpurun_out; T=r02f
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest_gpu.log; tail -3 $O/${T}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${T}_smoke.log; tail -2 $O/${T}_smoke.log
timeout 900 python bench.py > $O/${T}_bench_1gpu.json 2> $O/${T}_bench_1gpu.err; echo "bench rc=$?"
python scripts/show_bench.py $O/${T}_bench_1gpu.json | cut -c1-260
timeout 600 python bench.py --impl reference > $O/${T}_bench_reference_c2.json 2> $O/${T}_bench_reference_c2.err; echo "reference c2 rc=$?"; tail -1 $O/${T}_bench_reference_c2.json | cut -c1-300
timeout 600 python bench.py --impl reference --workload c1 > $O/${T}_bench_reference_c1.json 2> $O/${T}_bench_reference_c1.err; echo "reference c1 rc=$?"; tail -1 $O/${T}_bench_reference_c1.json | cut -c1-300
timeout 300 python scripts/frame_breakdown.py c2 100 > $O/${T}_frame_breakdown_c2.txt 2>&1
timeout 300 python scripts/frame_breakdown.py c4 10 > $O/${T}_frame_breakdown_c4.txt 2>&1
head -3 $O/${T}_frame_breakdown_c2.txt | cut -c1-140; head -3 $O/${T}_frame_breakdown_c4.txt | cut -c1-140
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_launches.csv python bench.py --steps 3 --warmup 3 --also none --no-cpu-baseline > $O/${T}_ncu_launches.log 2>&1; echo "ncu launch list rc=$?"
ncu --set full --import-source on --clock-control none -k regex:"weight_lists_kernel|cand_build_kernel|cand_octant_kernel|cand_build_far_kernel|cand_mark_kernel" \
    --launch-skip 60 --launch-count 10 -o $O/${T}_c2_hot -f python scripts/frame_breakdown.py c2 14 > $O/${T}_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
ncu --set full --import-source on --clock-control none -k regex:"weight_lists_kernel|cand_build_kernel|cand_octant_kernel" \
    --launch-skip 18 --launch-count 3 -o $O/${T}_c4_hot -f python scripts/frame_breakdown.py c4 5 > $O/${T}_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
ls -la $O | grep ${T}_
