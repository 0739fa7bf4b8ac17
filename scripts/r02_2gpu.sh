#!/bin/bash
# N GPUs: the peer tests, then bench.py under torchrun (c2 + c4), and the single-process multi-device mode
O=gpurun_out; TAG=${1:-r02v}; N=${2:-2}
timeout 900 python -m pytest tests/test_gpu_peer.py tests/test_gpu_shard.py tests/test_gpu_shim.py -m gpu -x -q 2>&1 | tail -5 | cut -c1-250
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 > $O/${TAG}_bench_${N}gpu.json 2> $O/${TAG}_bench_${N}gpu.err
python scripts/show_bench.py $O/${TAG}_bench_${N}gpu.json 2>&1 | cut -c1-250
DEVS=$(python -c "print(','.join(str(i) for i in range($N)))")
timeout 600 python bench.py --devices $DEVS --steps 100 --also c4 --no-cpu-baseline > $O/${TAG}_bench_${N}dev_one_process.json 2> $O/${TAG}_bench_${N}dev_one_process.err
python scripts/show_bench.py $O/${TAG}_bench_${N}dev_one_process.json 2>&1 | cut -c1-250
tail -3 $O/${TAG}_bench_${N}dev_one_process.err | cut -c1-300
