"""Tuning aid: run the c2 bench workload on a PFT_STATS build and print the list / search statistics of weight()."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from pcl_tracking_b200 import pcl, _capi
n_particles = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
frames, oid0 = bench.make_frames(6)
ctx = pcl.Context(0)
model_cloud, centroid = pcl.prepare_model(pcl.PointCloud(bench.raw_model(frames, oid0), ctx=ctx), 0.01, ctx=ctx)
M = model_cloud.size()
t = pcl.ParticleFilterOMPTracker(16, ctx=ctx)
pcl.configure_like_reference(t, particle_num=n_particles, use_hsv=True)
m = np.eye(4, dtype=np.float32); m[:3, 3] = centroid
t.setTrans(m); t.seed(1234); t.setReferenceCloud(model_cloud)
vg = pcl.ApproximateVoxelGrid(ctx=ctx); vg.setLeafSize(0.01); vg.setPassThrough("z", 0.0, 10.0)
ds = pcl.PointCloud(ctx=ctx)
dev = [pcl.PointCloud(f, ctx=ctx) for f in frames]
lib = _capi.load()
out = (C.c_ulonglong * 48)()
for k in range(12):
    vg.setInputCloud(dev[bench.frame_order(k, 6)]); vg.filter(ds)
    t.setInputCloud(ds); t.compute()
    ctx.synchronize()
    lib.pft_debug_stats(out)
    v = list(out)
    if k in (1, 5, 11):
        print("frame %d (2 x weight()): list lookups %d, mean count %.2f; with pool groups %d (%.3f%%, mean count %.1f); warp-scanned queries %d (%.3f%% of all), "
              "brute-force queries %d, mean extended scan length %.0f" % (k, v[12], v[5] / max(v[12], 1), v[3], 100.0 * v[3] / max(v[12], 1), v[6] / max(v[3], 1),
                                                                   v[13], 100.0 * v[13] / max(v[12] + v[13], 1), v[4], v[14] / max(v[13], 1)))
        print("   build: octant lists %d, mean length %.2f, longer than seven %d; cells built %d, far cells %d (extended %d)" % (v[2], v[1] / max(v[2], 1), v[0], v[15], v[10], v[11]))
        print("   marked cells by queries per cell [1, 2-3, 4-7, ... >=512]: cells %s  queries %s; blocks %d" % (v[16:26], v[26:36], v[36]))
