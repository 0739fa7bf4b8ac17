import sys; sys.path.insert(0, '.')
import numpy as np
from pcl_tracking_b200 import pcl, synth
import bench
frames, oid0 = bench.make_frames(3)
ctx = pcl.Context(0)
raw_model_cloud = pcl.PointCloud(bench.raw_model(frames, oid0), ctx=ctx)
model_cloud, centroid = pcl.prepare_model(raw_model_cloud, 0.01, ctx=ctx)
print("model", model_cloud.size(), centroid)
dev = [pcl.PointCloud(f, ctx=ctx) for f in frames]
print("dev sizes", [d.size() for d in dev])
ds = pcl.PointCloud(ctx=ctx)
vg = pcl.ApproximateVoxelGrid(ctx=ctx); vg.setLeafSize(0.01,0.01,0.01); vg.setPassThrough("z",0.0,10.0)
for k in range(4):
    vg.setInputCloud(dev[k%3]); vg.filter(ds)
    print("ds", k, ds.size())
ds2 = pcl.PointCloud(ctx=ctx)
vg.setInputCloud(dev[0]); vg.filter(ds2); print("ds2", ds2.size())
vg2 = pcl.ApproximateVoxelGrid(ctx=ctx); vg2.setLeafSize(0.01); vg2.setPassThrough("z",0.0,10.0)
vg2.setInputCloud(pcl.PointCloud(frames[0], ctx=ctx)); print("fresh", vg2.filter().size())
