#!/bin/bash
# N GPUs: bench.py under torchrun (c2 weak + c4 strong) as the driver launches it; at N=2 also the peer tests, the
# per-kernel breakdown of the sharded frame and the single-process multi-device mode
O=gpurun_out; T=r02f; N=${1:-2}
if [ "$N" = 2 ]; then
  timeout 900 python -m pytest tests/test_gpu_peer.py tests/test_gpu_shard.py tests/test_gpu_shim.py -m gpu -q > $O/${T}_pytest_2gpu.log 2>&1; tail -2 $O/${T}_pytest_2gpu.log
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/frame_breakdown.py c2 60 > $O/${T}_frame_breakdown_c2_2gpu.txt 2>&1; grep -v "^\[W\|^\*\|OMP_NUM" $O/${T}_frame_breakdown_c2_2gpu.txt | head -4 | cut -c1-150
  timeout 600 python bench.py --devices 0,1 --steps 200 --also c4 --no-cpu-baseline > $O/${T}_bench_2dev_one_process.json 2> $O/${T}_bench_2dev_one_process.err
  python scripts/show_bench.py $O/${T}_bench_2dev_one_process.json 2>&1 | cut -c1-250
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --exchange nccl > $O/${T}_bench_2gpu_nccl.json 2> $O/${T}_bench_2gpu_nccl.err
  python scripts/show_bench.py $O/${T}_bench_2gpu_nccl.json 2>&1 | cut -c1-250
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 > $O/${T}_bench_${N}gpu.json 2> $O/${T}_bench_${N}gpu.err
python scripts/show_bench.py $O/${T}_bench_${N}gpu.json 2>&1 | cut -c1-250
