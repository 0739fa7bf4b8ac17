#!/bin/bash
# two GPUs: the peer tests, then bench.py under torchrun (c2 + c4), with each scene mode
O=gpurun_out; TAG=${1:-r02v}; N=${2:-2}
timeout 900 python -m pytest tests/test_gpu_peer.py tests/test_gpu_shard.py -m gpu -x -q 2>&1 | tail -5 | cut -c1-250
for scene in peer replicate; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --scene $scene > $O/${TAG}_bench_${N}gpu_$scene.json 2> $O/${TAG}_bench_${N}gpu_$scene.err
  python scripts/show_bench.py $O/${TAG}_bench_${N}gpu_$scene.json 2>&1 | cut -c1-250
done
