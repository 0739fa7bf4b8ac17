"""Per-kernel device times of one tracked frame (bench.py's c2 / c4 workload), from CUDA events recorded after
every kernel of compute() on the library's stream (stream-launched, no graph, warm caches).  Complements the
ncu launch list (cold caches, serialised): usage  python scripts/frame_breakdown.py [c2|c4] [frames] [particles]"""
import collections
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
from pcl_tracking_b200 import pcl  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "c2"
    n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    n_particles = int(sys.argv[3]) if len(sys.argv) > 3 else (1000 if workload == "c2" else 100000)
    rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    if world > 1:  # torchrun: one tracker sharded by particle over the ranks, NVLink peer exchange (as bench.py)
        import torch.distributed as dist
        dist.init_process_group("gloo")
        if workload == "c2" and len(sys.argv) <= 3:
            n_particles *= world
    ctx = pcl.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    # c5:<k> = object k of the eight-object scene of bench.py's c5 workload, tracked alone (one of the batch's trackers)
    obj_k = int(workload.split(":")[1]) if workload.startswith("c5") and ":" in workload else 0
    if workload.startswith("c5"):
        frames, oid0 = bench.make_frames(bench.N_FRAMES, bench.C5_OBJECTS)
        workload = "c2"
    else:
        frames, oid0 = bench.make_frames(bench.N_FRAMES)
    model_cloud, centroid = pcl.prepare_model(pcl.PointCloud(bench.raw_model(frames, oid0, obj_k), ctx=ctx), bench.LEAF, ctx=ctx)
    M = model_cloud.size()
    tracker = pcl.ParticleFilterOMPTracker(16, ctx=ctx)
    pcl.configure_like_reference(tracker, particle_num=n_particles, use_hsv=True, iteration_num=bench.ITERATIONS)
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = centroid
    tracker.setTrans(m)
    tracker.seed(1234)
    if world > 1:
        tracker.setShard(world, rank)
        handles = [None] * world
        dist.all_gather_object(handles, tracker.peerExport())
        tracker.peerAttach(handles)
        dist.barrier()
    tracker.setReferenceCloud(model_cloud)
    dev_frames = [pcl.PointCloud(f, ctx=ctx) for f in frames]
    ds = pcl.PointCloud(ctx=ctx)
    vg = pcl.ApproximateVoxelGrid(ctx=ctx)
    vg.setLeafSize(bench.LEAF, bench.LEAF, bench.LEAF)
    vg.setPassThrough("z", 0.0, 10.0)

    def step(k):
        vg.setInputCloud(dev_frames[bench.frame_order(k, bench.N_FRAMES)])
        vg.filter(ds)
        tracker.setInputCloud(ds)
        tracker.compute()

    for k in range(5):
        step(k)
    ctx.synchronize()
    # graph-replayed frame time for reference
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(n_frames):
        step(5 + k)
    e1.record(stream)
    ctx.synchronize()
    graph_ms = e0.elapsed_time(e1) / n_frames
    tracker.enableTiming(True)
    agg = collections.OrderedDict()
    by_occ = collections.OrderedDict()  # (kernel, n-th launch of it in the frame): the iterations of a frame differ (index reuse)
    k1_ms = 0.0
    total = 0.0
    for k in range(n_frames):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        vg.setInputCloud(dev_frames[bench.frame_order(k, bench.N_FRAMES)])
        vg.filter(ds)
        b.record(stream)
        tracker.setInputCloud(ds)
        tracker.compute()
        times = tracker.kernelTimes()
        k1_ms += a.elapsed_time(b)
        occ = collections.Counter()
        for name, ms in times:
            e = agg.setdefault(name, [0, 0.0])
            e[0] += 1
            e[1] += ms
            total += ms
            occ[name] += 1
            o = by_occ.setdefault((name, occ[name]), [0, 0.0])
            o[0] += 1
            o[1] += ms
    out = {"workload": workload, "particles": n_particles, "model_points": M, "frames": n_frames,
           "frame_ms_graph_replay": graph_ms, "frame_ms_stream_launched_sum": (total + k1_ms) / n_frames,
           "k1_downsample_ms_per_frame": k1_ms / n_frames, "index": tracker.indexInfo(), "kernels": {}}
    if rank != 0:
        if world > 1:
            dist.barrier()
        return
    print("%s: %d particles x %d pts; frame %.3f ms (graph replay), %.3f ms (sum of stream-launched kernels incl. K1 %.3f ms)"
          % (workload, n_particles, M, graph_ms, (total + k1_ms) / n_frames, k1_ms / n_frames))
    for name, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("  %-28s %5.1f launches/frame  %8.2f us/launch  %8.2f us/frame  %5.1f%%"
              % (name, cnt / n_frames, 1e3 * ms / cnt, 1e3 * ms / n_frames, 100.0 * ms / (total + k1_ms)))
        out["kernels"][name] = {"launches_per_frame": cnt / n_frames, "us_per_launch": 1e3 * ms / cnt, "us_per_frame": 1e3 * ms / n_frames}
    print("  launch order (us per launch, by occurrence in the frame):")
    print("   " + "  ".join("%s#%d %.1f" % (name.replace("_kernel", ""), k, 1e3 * ms / cnt) for (name, k), (cnt, ms) in by_occ.items()))
    print(json.dumps(out))
    if world > 1:
        dist.barrier()


if __name__ == "__main__":
    main()
