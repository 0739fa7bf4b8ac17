"""The multi-GPU algorithm on CPU with world_size 2 over gloo: each rank weights the particles
i % R == rank with the CPU oracle, the crop box is all-reduced (min/max), the raw-weight slices are
all-gathered in the library's [R][slice_cap] layout, and normalise runs replicated.  The result must equal
the single-process oracle bit for bit.  (The NCCL path itself is exercised on GPUs by bench.py --gpus N and
emulated on one GPU by tests/test_gpu_shard.py.)"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from pcl_tracking_b200 import sharding
from tests import util


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case():
    scene, model, centre = util.small_case(31, n_scene=3000, n_model=200)
    parts = util.particles_around(centre, 77, seed=5)
    return scene, model, parts


def _tracker(parts_subset, model, scene):
    t = oracle.Tracker(kld=False)
    oracle.configure_like_reference(t, particle_num=len(parts_subset), use_hsv=True, nn_mode=oracle.NN_EXACT_GRID)
    t.set_reference(model)
    t.set_input(scene)
    t.set_particles(parts_subset)
    return t


def _rank_main(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene, model, parts = _case()
    n = len(parts)
    mine = sharding.local_particles(n, world, rank)
    t = _tracker(parts[mine], model, scene)
    t.weight()                                   # pass 1: only to get this rank's share of the crop box
    box = torch.from_numpy(t.local_aabb().copy())
    lo, hi = box[:3].clone(), box[3:].clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    t.set_particles(parts[mine])
    t.set_crop_box(torch.cat([lo, hi]).numpy())
    t.weight()                                   # pass 2: index + coherence with the global crop box
    cap = sharding.slice_cap(n, world)
    local = torch.zeros(cap, dtype=torch.float32)
    local[: len(mine)] = torch.from_numpy(t.raw_weights())
    gathered = [torch.zeros(cap, dtype=torch.float32) for _ in range(world)]
    dist.all_gather(gathered, local)
    raw = sharding.assemble(torch.stack(gathered).numpy(), n, world)
    # replicated normalise on every rank
    full = oracle.Tracker(kld=False)
    oracle.configure_like_reference(full, particle_num=n)
    p = parts.copy()
    p["weight"] = raw
    full.set_particles(p)
    full.normalize()
    if rank == 0:
        np.save(out, np.stack([raw, full.get_particles()["weight"]]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_oracle_world_size_2(tmp_path):
    out = str(tmp_path / "shard.npy")
    port = _free_port()
    mp.start_processes(_rank_main, args=(2, port, out), nprocs=2, join=True, start_method="spawn")
    got = np.load(out)
    scene, model, parts = _case()
    ref = _tracker(parts, model, scene)
    ref.weight()
    assert np.array_equal(got[0].view(np.uint32), ref.raw_weights().view(np.uint32))
    assert np.array_equal(got[1].view(np.uint32), ref.get_particles()["weight"].view(np.uint32))


def test_sharding_layout_roundtrip():
    for n, r in ((1, 1), (7, 2), (100, 8), (101, 3), (1000, 8)):
        cap = sharding.slice_cap(n, r)
        buf = np.full(r * cap, -1, dtype=np.int64)
        for rank in range(r):
            mine = sharding.local_particles(n, r, rank)
            assert all(sharding.owner(i, r) == rank for i in mine)
            buf[rank * cap: rank * cap + len(mine)] = mine
        assert np.array_equal(sharding.assemble(buf, n, r), np.arange(n))
        assert all(buf[sharding.raw_slot(i, r, cap)] == i for i in range(n))
