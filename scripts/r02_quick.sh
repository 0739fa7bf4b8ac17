#!/bin/bash
# tuning: breakdown of c2 / c4 with the default build + the quick parity tests
O=gpurun_out; TAG=${1:-r02w}
timeout 300 python scripts/frame_breakdown.py c2 100 > $O/${TAG}_c2.txt 2>&1
timeout 300 python scripts/frame_breakdown.py c4 10 > $O/${TAG}_c4.txt 2>&1
head -18 $O/${TAG}_c2.txt | cut -c1-600; head -8 $O/${TAG}_c4.txt | cut -c1-110
if [ "${PYTEST:-1}" = 1 ]; then timeout 900 python -m pytest tests/test_gpu_weight.py tests/test_gpu_compute.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3; fi
