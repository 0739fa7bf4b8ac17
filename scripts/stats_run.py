"""Tuning aid: run the c2 bench workload on a PFT_STATS build and print the search statistics."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from pcl_tracking_b200 import pcl, _capi
frames, oid0 = bench.make_frames(6)
ctx = pcl.Context(0)
model_cloud, centroid = pcl.prepare_model(pcl.PointCloud(bench.raw_model(frames, oid0), ctx=ctx), 0.01, ctx=ctx)
M = model_cloud.size()
t = pcl.ParticleFilterOMPTracker(16, ctx=ctx)
pcl.configure_like_reference(t, particle_num=1000, use_hsv=True)
m = np.eye(4, dtype=np.float32); m[:3, 3] = centroid
t.setTrans(m); t.seed(1234); t.setReferenceCloud(model_cloud)
vg = pcl.ApproximateVoxelGrid(ctx=ctx); vg.setLeafSize(0.01); vg.setPassThrough("z", 0.0, 10.0)
ds = pcl.PointCloud(ctx=ctx)
dev = [pcl.PointCloud(f, ctx=ctx) for f in frames]
lib = _capi.load()
out = (C.c_ulonglong * 16)()
names = ["patch calls", "passes", "sum list total", "valid lanes", "fallback lanes", "unusable groups", "nn_search calls", "rows visited", "rows scanned",
         "nn cands", "sum delta", "fallback w/o match"]
for k in range(12):
    vg.setInputCloud(dev[bench.frame_order(k, 6)]); vg.filter(ds)
    t.setInputCloud(ds); t.compute()
    ctx.synchronize()
    lib.pft_debug_stats(out)
    v = list(out)
    if k in (1, 11):
        print("   lists: lookups %d, overflow %.4f%%, mean list length %.1f, cells built %d; lookups with length >16: %d, >32: %d, >64: %d, >128: %d, >512: %d, ==0: %d"
              % (v[12], 100.0 * v[13] / max(v[12], 1), v[14] / max(v[12] - v[13], 1), v[15], v[0], v[1], v[2], v[3], v[4], v[5]))
        print("   far cells queued %d, of which rebuilt as extended lists %d" % (v[10], v[11]))
    if k in (0, 1, 5, 11):
        print("frame", k, {n: v[i] for i, n in enumerate(names)})
        pc = max(v[0], 1)
        print("   per patch-call: passes %.2f, list/group-pass %.1f, delta %.2f; fallback %.1f%% of lanes (%.1f%% without match); nn_search: rows visited %.1f scanned %.1f cands %.1f per call"
              % (v[1] / pc, v[2] / max(v[1], 1) / (32 // 8), v[10] / max(v[1], 1), 100.0 * v[4] / max(v[3], 1), 100.0 * v[11] / max(v[3], 1),
                 v[7] / max(v[6], 1), v[8] / max(v[6], 1), v[9] / max(v[6], 1)))
