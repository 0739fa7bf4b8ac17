"""Frame ingest overlapped with compute: pft_cloud_upload_async / pft_cloud_upload_pointcloud2_async.
The asynchronous uploads must deliver the same bytes as the synchronous ones, and a double-buffered frame loop
(frame k+1 copied while frame k is tracked) must give the same poses as the serial loop, bit for bit."""
import numpy as np
import pytest

from pcl_tracking_b200 import pcl, synth
from pcl_tracking_b200._capi import POINT, POINT_PCL32

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32).reshape(len(a), 4)


def _cloud(n, seed):
    rng = np.random.default_rng(seed)
    out = np.zeros(n, dtype=POINT)
    for k in ("x", "y", "z"):
        out[k] = rng.uniform(-2, 2, n).astype(np.float32)
    out["rgba"] = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    return out


@pytest.mark.parametrize("n", [0, 1, 4097, 217088])
def test_async_upload_delivers_the_same_points(n):
    pts = _cloud(n, n + 3)
    buf = pcl.PinnedBuffer.of(pts)
    c = pcl.PointCloud()
    c.upload_raw_async(buf.ptr or 1, n)
    assert c.size() == n
    assert np.array_equal(_bits(c.to_numpy()), _bits(pts))
    # 32-byte PCL records through the per-cloud staging buffer
    p32 = np.zeros(n, dtype=POINT_PCL32)
    for k in ("x", "y", "z", "rgba"):
        p32[k] = pts[k]
    b32 = pcl.PinnedBuffer.of(p32)
    c2 = pcl.PointCloud()
    c2.upload_raw_async(b32.ptr or 1, n, pcl.capi.LAYOUT_PCL32)
    c2.waitUpload()
    assert np.array_equal(_bits(c2.to_numpy()), _bits(pts))
    # re-upload into the same cloud while nothing else is queued: the second frame wins
    pts_b = _cloud(n, n + 4)
    buf_b = pcl.PinnedBuffer.of(pts_b)
    c.upload_raw_async(buf_b.ptr or 1, n)
    assert np.array_equal(_bits(c.to_numpy()), _bits(pts_b))


def test_async_pointcloud2_ingest_matches_synchronous():
    width, height, point_step = 512, 424, 32
    pts = _cloud(width * height, 9)
    rec = np.zeros((width * height, point_step), dtype=np.uint8)
    for name, off in zip(("x", "y", "z", "rgba"), (0, 4, 8, 16)):
        rec[:, off:off + 4] = np.ascontiguousarray(pts[name]).view(np.uint8).reshape(-1, 4)
    buf = pcl.PinnedBuffer.of(rec)
    a = pcl.PointCloud().fromPointCloud2(buf.ptr, width, height, point_step)
    b = pcl.PointCloud().fromPointCloud2(buf.ptr, width, height, point_step, asynchronous=True)
    assert np.array_equal(_bits(a.to_numpy()), _bits(b.to_numpy()))
    assert np.array_equal(_bits(b.to_numpy()), _bits(pts))
    with pytest.raises(pcl.PftError):
        pcl.PointCloud().fromPointCloud2(buf.ptr, width, height, 10, asynchronous=True)


def _frame_loop(frames, model, centre, pipelined):
    t = pcl.KLDAdaptiveParticleFilterOMPTracker(16)
    pcl.configure_like_reference(t, particle_num=300, max_particle_num=400, use_hsv=True)
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = centre
    t.setTrans(m)
    t.seed(99)
    t.setReferenceCloud(model)
    vg = pcl.ApproximateVoxelGrid()
    vg.setLeafSize(0.01, 0.01, 0.01)
    vg.setPassThrough("z", 0.0, 10.0)
    bufs = [pcl.PinnedBuffer.of(f) for f in frames]
    clouds = [pcl.PointCloud(), pcl.PointCloud()]
    ds = pcl.PointCloud()
    poses, sizes = [], []
    if pipelined:
        clouds[0].upload_raw_async(bufs[0].ptr, len(frames[0]))
    for k in range(len(frames)):
        cur = clouds[k % 2]
        if pipelined:
            if k + 1 < len(frames):   # the next frame travels while this one is tracked
                clouds[(k + 1) % 2].upload_raw_async(bufs[k + 1].ptr, len(frames[k + 1]))
        else:
            cur.upload_raw(bufs[k].ptr, len(frames[k]))
        vg.setInputCloud(cur)
        vg.filter(ds)
        t.setInputCloud(ds)
        t.compute()
        poses.append(np.array(list(t.getResult().tolist()), dtype=np.float64))
        sizes.append(ds.size())
    return np.array(poses), sizes


def test_double_buffered_frame_loop_equals_serial_loop():
    objs = synth.default_objects(1)
    frames = [synth.render(f, objs)[0] for f in range(6)]
    pts0, oid0 = synth.render(0, objs)
    model, centre = pcl.prepare_model(pcl.PointCloud(synth.model_points(pts0, oid0, 0)), 0.01)
    serial, sizes_a = _frame_loop(frames, model, centre, pipelined=False)
    piped, sizes_b = _frame_loop(frames, model, centre, pipelined=True)
    assert sizes_a == sizes_b
    np.testing.assert_array_equal(serial, piped)
