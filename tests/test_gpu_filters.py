"""Parity of kernel K1 (PassThrough / voxel-grid downsampling / model preparation) with the CPU oracle,
through the C ABI.  Integer/byte results (order, counts, colours) are bit-exact; centroids are
bit-exact too because both sides accumulate fp32 coordinates in fp64 (exact, order independent)."""
import numpy as np
import pytest

import oracle
from pcl_tracking_b200 import pcl, synth

pytestmark = pytest.mark.gpu


def _random_cloud(n, seed, nan_frac=0.05, span=2.0):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-span, span, (n, 3)).astype(np.float32)
    xyz[:, 2] = rng.uniform(-1.0, 11.0, n)
    bad = rng.random(n) < nan_frac
    xyz[bad, rng.integers(0, 3, bad.sum())] = np.nan
    xyz[rng.random(n) < 0.01, 1] = np.inf
    return oracle.make_points(xyz, rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32))


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32).reshape(len(a), 4)


@pytest.mark.parametrize("n,seed", [(0, 0), (1, 1), (37, 2), (5000, 3), (70001, 4)])
def test_passthrough_bit_exact(n, seed):
    pts = _random_cloud(n, seed)
    for field, lo, hi in ((2, 0.0, 10.0), (0, -0.5, 0.25), (1, -1e9, 1e9)):
        pt = pcl.PassThrough()
        pt.setFilterFieldName("xyz"[field])
        pt.setFilterLimits(lo, hi)
        pt.setKeepOrganized(False)
        pt.setInputCloud(pcl.PointCloud(pts))
        got = pt.filter().to_numpy()
        want = oracle.passthrough(pts, field, lo, hi)
        assert len(got) == len(want)
        assert np.array_equal(_bits(got), _bits(want))


def test_passthrough_inclusive_limits():
    pts = oracle.make_points([[0, 0, 0.0], [0, 0, 10.0], [0, 0, np.nextafter(np.float32(10.0), np.float32(11.0))], [0, 0, -1e-30]])
    pt = pcl.PassThrough()
    pt.setFilterFieldName("z")
    pt.setFilterLimits(0.0, 10.0)
    pt.setInputCloud(pcl.PointCloud(pts))
    assert len(pt.filter()) == 2


@pytest.mark.parametrize("n,seed,leaf", [(0, 0, 0.01), (1, 1, 0.01), (500, 2, 0.05), (20000, 3, 0.01), (60000, 4, 0.1)])
def test_voxel_grid_bit_exact(n, seed, leaf):
    pts = _random_cloud(n, seed, span=0.6 if leaf < 0.05 else 2.0)
    vg = pcl.ApproximateVoxelGrid()
    vg.setLeafSize(leaf, leaf, leaf)
    vg.setPassThrough("z", 0.0, 10.0)
    vg.setInputCloud(pcl.PointCloud(pts))
    got = vg.filter().to_numpy()
    want = oracle.voxel_grid_exact(pts, leaf, 2, 0.0, 10.0)
    assert len(got) == len(want)
    assert np.array_equal(_bits(got), _bits(want))  # same voxels, same order (first appearance), same centroids, same colours


def test_voxel_grid_without_passthrough_and_dense_voxels():
    rng = np.random.default_rng(9)
    xyz = rng.uniform(0, 0.03, (30000, 3)).astype(np.float32)  # 27 voxels, ~1100 points each
    pts = oracle.make_points(xyz, rng.integers(0, 1 << 24, 30000).astype(np.uint32))
    vg = pcl.VoxelGrid()
    vg.setLeafSize(0.01)
    vg.setInputCloud(pcl.PointCloud(pts))
    got = vg.filter().to_numpy()
    want = oracle.voxel_grid_exact(pts, 0.01, -1, 0, 0)
    assert len(got) == 27 == len(want)
    assert np.array_equal(_bits(got), _bits(want))


def test_voxel_grid_kinect_frame_full_size_properties():
    """BASELINE size (217 088 points): bit-exact against the oracle, plus size-independent properties."""
    pts, _ = synth.render(0)
    vg = pcl.ApproximateVoxelGrid()
    vg.setLeafSize(0.01, 0.01, 0.01)
    vg.setPassThrough("z", 0.0, 10.0)
    cloud = pcl.PointCloud(pts)
    vg.setInputCloud(cloud)
    ds = vg.filter()
    got = ds.to_numpy()
    want = oracle.voxel_grid_exact(pts, 0.01, 2, 0.0, 10.0)
    assert np.array_equal(_bits(got), _bits(want))
    # one output per occupied voxel
    inv = np.float32(1.0) / np.float32(0.01)
    ok = np.isfinite(pts["x"]) & np.isfinite(pts["y"]) & np.isfinite(pts["z"]) & (pts["z"] >= 0) & (pts["z"] <= 10)
    keys = np.stack([np.floor(pts[k][ok] * inv).astype(np.int64) for k in "xyz"], 1)
    assert len(got) == len(np.unique(keys, axis=0))
    # idempotence: a centroid stays in its voxel, so downsampling the downsampled cloud is the identity
    vg2 = pcl.ApproximateVoxelGrid()
    vg2.setLeafSize(0.01)
    vg2.setInputCloud(ds)
    again = vg2.filter().to_numpy()
    gk = np.stack([np.floor(got[k] * inv).astype(np.int64) for k in "xyz"], 1)
    if len(np.unique(gk, axis=0)) == len(got):  # (a centroid can round onto a voxel face; then two merge)
        assert np.array_equal(_bits(again)[:, :3], _bits(got)[:, :3])
    # against PCL's own 512-slot ApproximateVoxelGrid (restated): same voxel set, PCL emits duplicates
    approx = oracle.approx_voxel_grid_pcl(oracle.passthrough(pts, 2, 0.0, 10.0), 0.01)
    assert len(approx) >= len(got)
    ak = np.unique(np.stack([np.floor(approx[k] * inv).astype(np.int64) for k in "xyz"], 1), axis=0)
    assert abs(len(ak) - len(got)) <= 0.001 * len(got)


def test_pcl32_layout_roundtrip():
    pts = _random_cloud(1000, 5, nan_frac=0)
    p32 = synth.to_pcl32(pts)
    cloud = pcl.PointCloud(np.ascontiguousarray(p32).view(pcl.POINT_PCL32).reshape(-1))
    back16 = cloud.to_numpy()
    assert np.array_equal(_bits(back16), _bits(pts))
    back32 = cloud.to_numpy(pcl32=True)
    assert np.array_equal(back32["x"], pts["x"]) and np.array_equal(back32["rgba"], pts["rgba"]) and np.all(back32["w"] == 1.0)


def test_prepare_model_matches_reference_flow():
    """removeZeroPoints -> centroid -> translate -> VoxelGrid (ref: src/auto_tracking.cpp:656-674)."""
    pts, oid = synth.render(0)
    raw = synth.model_points(pts, oid, 0)
    raw = np.concatenate([raw, oracle.make_points([[0.001, -0.002, 0.003], [0, 0, 0]])])  # "zero points" to be removed
    model, c = pcl.prepare_model(pcl.PointCloud(raw), 0.01)
    got = model.to_numpy()
    r = oracle.remove_zero_points(raw)
    assert len(r) == len(raw) - 2
    c_ref = oracle.centroid(r)  # PCL accumulates the centroid in fp32; the GPU in fp64
    np.testing.assert_allclose(c, c_ref, atol=2e-5)
    r["x"] -= c[0]; r["y"] -= c[1]; r["z"] -= c[2]  # same translation as the GPU used
    want = oracle.voxel_grid_exact(r, 0.01, -1, 0, 0)
    assert len(got) == len(want)
    assert np.array_equal(_bits(got), _bits(want))
    # and against PCL's exact VoxelGrid (sorted output order, fp32 sums): same set within float tolerance
    pclvg = oracle.voxel_grid_pcl(r, 0.01)
    assert len(pclvg) == len(got)
    key = lambda a: np.lexsort((np.floor(a["z"] * 100), np.floor(a["y"] * 100), np.floor(a["x"] * 100)))
    a, b = got[key(got)], pclvg[key(pclvg)]
    for k in "xyz":
        np.testing.assert_allclose(a[k], b[k], atol=1e-6)


def test_all_points_rejected():
    pts = oracle.make_points(np.full((100, 3), np.nan))
    vg = pcl.ApproximateVoxelGrid()
    vg.setInputCloud(pcl.PointCloud(pts))
    assert len(vg.filter()) == 0


def _pointcloud2_payload(pts, width, height, point_step, row_step, offs, seed):
    """Serialise packed points into a sensor_msgs/PointCloud2 byte payload with the given strides/offsets; padding
    bytes are random so that a parser reading the wrong bytes is caught."""
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 256, (height, row_step), dtype=np.uint8)
    rec = np.zeros((height * width, point_step), dtype=np.uint8)
    rec[:] = rng.integers(0, 256, rec.shape, dtype=np.uint8)
    for name, off in zip(("x", "y", "z", "rgba"), offs):
        if off >= 0:
            rec[:, off:off + 4] = np.ascontiguousarray(pts[name]).view(np.uint8).reshape(-1, 4)
    raw[:, :width * point_step] = rec.reshape(height, width * point_step)
    return raw


@pytest.mark.parametrize("width,height,point_step,pad,offs", [
    (512, 424, 32, 0, (0, 4, 8, 16)),     # kinect2_bridge: the PointXYZRGBA record as PCL serialises it
    (640, 480, 16, 64, (0, 4, 8, 12)),    # packed xyz + rgb, padded rows
    (97, 3, 24, 8, (8, 12, 16, 0)),       # colour first, ragged sizes
    (33, 1, 12, 0, (0, 4, 8, -1)),        # xyz only
    (0, 0, 32, 0, (0, 4, 8, 16)),         # empty message
])
def test_pointcloud2_ingest_bit_exact(width, height, point_step, pad, offs):
    """Frame ingest: PointCloud2 payload -> device cloud == pcl::fromPCLPointCloud2 restated."""
    n = width * height
    pts = _random_cloud(n, 5 + width)
    row_step = width * point_step + pad
    raw = _pointcloud2_payload(pts, width, height, point_step, row_step, offs, 11)
    want = oracle.from_pointcloud2(raw.tobytes(), width, height, point_step, row_step, *offs)
    if offs[3] < 0:
        assert not want["rgba"].any()
    else:
        assert np.array_equal(_bits(want), _bits(pts))  # the oracle parser recovers what was serialised
    got = pcl.PointCloud().fromPointCloud2(raw, width, height, point_step, row_step, *offs).to_numpy()
    assert len(got) == n
    assert np.array_equal(_bits(got), _bits(want))
    # and the ingest feeds the frame pipeline: PassThrough(z) of the ingested cloud == of the original points
    if n:
        pt = pcl.PassThrough()
        pt.setFilterFieldName("z")
        pt.setFilterLimits(0.0, 10.0)
        pt.setInputCloud(pcl.PointCloud().fromPointCloud2(raw, width, height, point_step, row_step, *offs))
        assert np.array_equal(_bits(pt.filter().to_numpy()), _bits(oracle.passthrough(want, 2, 0.0, 10.0)))


def test_pointcloud2_ingest_rejects_bad_layouts():
    c = pcl.PointCloud()
    data = np.zeros(64, dtype=np.uint8)
    for kw in (dict(point_step=10), dict(point_step=16, row_step=8), dict(point_step=16, off_x=14), dict(point_step=16, off_z=3),
               dict(point_step=16, is_bigendian=True)):
        args = dict(width=2, height=1, point_step=16, off_rgb=12)
        args.update(kw)
        with pytest.raises(pcl.capi.PftError):
            c.fromPointCloud2(data, **args)


def _cluster_scene(seed, n_blobs, n_noise):
    rng = np.random.default_rng(seed)
    parts = []
    for b in range(n_blobs):
        c = rng.uniform(-0.6, 0.6, 3) + np.array([0, 0, 1.2])
        m = int(rng.integers(150, 900))
        q = rng.uniform(-0.5, 0.5, (m, 3)) * rng.uniform(0.04, 0.12, 3)   # dense blob: 1-2 cm spacing
        parts.append(c + q)
    parts.append(rng.uniform(-1.0, 1.0, (n_noise, 3)) + np.array([0, 0, 1.2]))   # sparse clutter: singletons and small groups
    xyz = np.concatenate(parts).astype(np.float32)
    xyz = xyz[rng.permutation(len(xyz))]
    xyz[rng.random(len(xyz)) < 0.01] = np.nan
    return oracle.make_points(xyz, rng.integers(0, 1 << 32, len(xyz), dtype=np.uint64).astype(np.uint32))


@pytest.mark.parametrize("seed,n_blobs,n_noise,tol,mn,mx", [(1, 6, 2000, 0.02, 100, 25000), (2, 12, 500, 0.02, 200, 600), (3, 3, 0, 0.015, 1, 25000),
                                                            (4, 0, 300, 0.02, 500, 25000)])
def test_euclidean_clusters_match_oracle(seed, n_blobs, n_noise, tol, mn, mx):
    """Model acquisition: EuclideanClusterExtraction on the GPU == the BFS restatement: same components,
    same size filter, same ranking; each cluster cloud = the input points of its (sorted) indices."""
    pts = _cluster_scene(seed, n_blobs, n_noise)
    want_labels, want_sizes = oracle.euclidean_clusters(pts, tol, mn, mx)
    cloud = pcl.PointCloud(pts)
    ec = pcl.EuclideanClusterExtraction()
    ec.setClusterTolerance(tol); ec.setMinClusterSize(mn); ec.setMaxClusterSize(mx); ec.setInputCloud(cloud)
    clusters = ec.extract()
    assert ec.sizes.tolist() == want_sizes.tolist()
    assert np.array_equal(ec.labels, want_labels)
    for k, idx in enumerate(clusters):
        assert len(idx) == want_sizes[k]
        got = ec.cluster_cloud(k).to_numpy()
        assert np.array_equal(_bits(got), _bits(pts[idx]))


def test_euclidean_clusters_edge_cases():
    ec = pcl.EuclideanClusterExtraction()
    ec.setClusterTolerance(0.02); ec.setMinClusterSize(1); ec.setMaxClusterSize(10)
    ec.setInputCloud(pcl.PointCloud(np.zeros(0, dtype=pcl.POINT)))
    assert ec.extract() == []
    one = oracle.make_points(np.array([[0.1, 0.2, 0.3]], dtype=np.float32))
    ec.setInputCloud(pcl.PointCloud(one))
    assert [c.tolist() for c in ec.extract()] == [[0]]
    with pytest.raises(pcl.capi.PftError):
        ec.setClusterTolerance(0.0); ec.extract()


def _plane_scene(seed, n_plane=30000, n_obj=6000):
    rng = np.random.default_rng(seed)
    xy = rng.uniform(-1, 1, (n_plane, 2))
    plane = np.c_[xy, 0.25 * xy[:, 0] + 0.4 * xy[:, 1] + 1.2 + rng.normal(0, 0.004, n_plane)]
    objs = [rng.uniform(-0.5, 0.5, (n_obj // 3, 3)) * [0.2, 0.2, 0.15] + [cx, cy, 0.25 * cx + 0.4 * cy + 1.2 - 0.12] for cx, cy in ((-0.4, 0.1), (0.2, -0.3), (0.5, 0.5))]
    xyz = np.concatenate([plane] + objs).astype(np.float32)
    xyz = xyz[rng.permutation(len(xyz))]
    xyz[rng.random(len(xyz)) < 0.01, 0] = np.nan
    return oracle.make_points(xyz, rng.integers(0, 1 << 32, len(xyz), dtype=np.uint64).astype(np.uint32)), rng


@pytest.mark.parametrize("seed,optimize", [(1, False), (2, False), (3, True), (4, True)])
def test_segment_plane_matches_oracle(seed, optimize):
    """Plane variant of the model builder (ref: src/create_model_planar_segmentation.cpp:157-174): same winning draw,
    same iteration count, same plane / rest split as the sequential RANSAC restatement for the same draws."""
    pts, rng = _plane_scene(seed)
    samples = rng.integers(0, len(pts), (1126, 3)).astype(np.int32)
    samples[3] = (7, 7, 11)                                     # degenerate draws are skipped
    want = oracle.segment_plane(pts, samples, 0.015, 1000, 0.99, optimize)
    seg = pcl.SACSegmentation()
    seg.setModelType("SACMODEL_PLANE"); seg.setMethodType("SAC_RANSAC"); seg.setMaxIterations(1000); seg.setDistanceThreshold(0.015)
    seg.setOptimizeCoefficients(optimize); seg.setSamples(samples); seg.setInputCloud(pcl.PointCloud(pts))
    coeff, n_in = seg.segment()
    assert seg.iterations == want["iterations"]
    if not optimize:
        assert np.array_equal(coeff.view(np.uint32), want["coeff"].view(np.uint32))      # the winning hypothesis, bit for bit
        m = want["inliers"]
    else:
        np.testing.assert_allclose(coeff, want["coeff"], atol=2e-6)                      # fp64 sums + eigen-solve on both sides
        # points whose distance is within 1e-5 of the threshold may fall either way
        x, y, z = pts["x"].astype(np.float64), pts["y"].astype(np.float64), pts["z"].astype(np.float64)
        with np.errstate(invalid="ignore"):
            d = np.abs(coeff[0] * x + coeff[1] * y + coeff[2] * z + coeff[3])
        m = want["inliers"]
        sure = np.abs(d - 0.015) > 1e-5
        got_mask = d < 0.015
        assert np.array_equal(got_mask[sure], m[sure])
        m = None
    plane, rest = seg.plane.to_numpy(), seg.rest.to_numpy()
    assert len(plane) == n_in and len(plane) + len(rest) == len(pts)
    if m is not None:
        assert np.array_equal(_bits(plane), _bits(pts[m])) and np.array_equal(_bits(rest), _bits(pts[~m]))
    assert n_in > 0.95 * 30000 * 0.99 and n_in < 30000 + 3000                            # the table, not the objects


def test_segment_plane_feeds_the_clustering():
    """The whole plane variant of create_model: plane removal, then Euclidean clustering of what is left -> 3 objects."""
    pts, rng = _plane_scene(9)
    seg = pcl.SACSegmentation()
    seg.setMaxIterations(1000); seg.setDistanceThreshold(0.015); seg.setSeed(42); seg.setInputCloud(pcl.PointCloud(pts))
    coeff, n_in = seg.segment()
    assert abs(abs(coeff[2]) - 1.0 / np.sqrt(1 + 0.25 ** 2 + 0.4 ** 2)) < 5e-3
    ec = pcl.EuclideanClusterExtraction()
    ec.setClusterTolerance(0.02); ec.setMinClusterSize(500); ec.setMaxClusterSize(25000); ec.setInputCloud(seg.rest)
    assert len(ec.extract()) == 3


@pytest.mark.parametrize("n,seed,leaf", [(0, 1, 0.01), (5, 2, 0.01), (20000, 3, 0.05), (60000, 4, 0.01)])
def test_pcl_approximate_voxel_grid_parity_mode(n, seed, leaf):
    """The reference's own downsample (pcl::ApproximateVoxelGrid, ref: src/auto_tracking.cpp:563-575) reproduced bit for
    bit, duplicates and flush order included: PassThrough(z) + the 512-entry cache walked in input order."""
    pts = _random_cloud(n, seed, span=0.6)
    want = oracle.approx_voxel_grid_pcl(oracle.passthrough(pts, 2, 0.0, 10.0), leaf)
    vg = pcl.ApproximateVoxelGrid()
    vg.setLeafSize(leaf); vg.setPassThrough("z", 0.0, 10.0); vg.setPclApproximateMode(True)
    vg.setInputCloud(pcl.PointCloud(pts))
    got = vg.filter().to_numpy()
    assert len(got) == len(want)
    assert np.array_equal(_bits(got), _bits(want))
    if n >= 20000:
        # upstream's cache emits some voxels more than once; the production path emits each voxel exactly once
        vg.setPclApproximateMode(False)
        assert len(vg.filter().to_numpy()) < len(want)


def test_pcl_approximate_voxel_grid_on_a_kinect_frame():
    pts, _ = synth.render(0, synth.default_objects(1))
    want = oracle.approx_voxel_grid_pcl(oracle.passthrough(pts, 2, 0.0, 10.0), 0.01)
    vg = pcl.ApproximateVoxelGrid()
    vg.setLeafSize(0.01); vg.setPassThrough("z", 0.0, 10.0); vg.setPclApproximateMode(True)
    vg.setInputCloud(pcl.PointCloud(pts))
    got = vg.filter().to_numpy()
    assert np.array_equal(_bits(got), _bits(want))
