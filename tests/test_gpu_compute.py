"""compute() end to end on the GPU: control flow of Tracker::compute / computeTracking,
CUDA-graph replay, empty inputs, multi-object batches, and tracking quality on synthetic motion."""
import numpy as np
import pytest

import oracle
from pcl_tracking_b200 import pcl, synth
from tests import util

pytestmark = pytest.mark.gpu


def _tracker_for(model, centre, kld=True, n=300, nmax=500, use_hsv=False, seed=1):
    t = pcl.KLDAdaptiveParticleFilterOMPTracker(16) if kld else pcl.ParticleFilterOMPTracker(16)
    pcl.configure_like_reference(t, particle_num=n, max_particle_num=nmax, use_hsv=use_hsv)
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = centre
    t.setTrans(m)
    t.seed(seed)
    t.setReferenceCloud(model)
    return t


def test_first_compute_matches_oracle_control_flow():
    """First compute(): initParticles, iteration 0 = weight+update only (changed_ is false), iteration 1 =
    resample+weight+update.  States agree with the oracle to 1e-4 (the chain amplifies ulp-level
    differences of the weights into the next iteration's selection only at CDF boundaries)."""
    scene, model, centre = util.small_case(8, n_scene=5000, n_model=300)
    g, o = util.make_pair(kld=True, particle_num=100, max_particle_num=220, use_hsv=True, oracle_nn=oracle.NN_EXACT_GRID)
    util.set_trans_both(g, o, centre)
    d = synth.draws(2, 220, seed=21)
    g.injectDraws(*d); o.inject_draws(*d)
    cloud = pcl.PointCloud(scene)
    g.setReferenceCloud(model); g.setInputCloud(cloud); g.compute()
    o.set_reference(model); o.set_input(scene); o.compute()
    gp, op = g.getParticles(), o.get_particles()
    assert len(gp) == len(op) and len(gp) != 100
    same = g.ancestors() == o.ancestors()
    assert same.mean() > 0.97
    for k in ("x", "y", "z", "roll", "pitch", "yaw"):
        np.testing.assert_allclose(gp[k][same], op[k][same], atol=1e-4)
    gr, orr = g.getResult(), o.get_result()
    for k in ("x", "y", "z"):
        assert abs(float(gr[k]) - float(orr[k])) < 2e-3


def _cdf_after_first_iteration(scene, model, centre, draws, n, nmax):
    """Cumulative selection table the resample of iteration 1 draws from, from the oracle run stage by stage
    (initParticles, weight, update: what iteration 0 of compute() does)."""
    _, o = util.make_pair(kld=True, particle_num=n, max_particle_num=nmax, use_hsv=True, oracle_nn=oracle.NN_EXACT_GRID)
    m = np.eye(4, dtype=np.float32); m[:3, 3] = centre
    o.set_trans(m[:3])
    o.inject_draws(*draws)
    o.set_reference(model); o.set_input(scene)
    o.init_particles(); o.weight(); o.update()
    w = o.get_particles()["weight"].astype(np.float64)
    return np.cumsum(w) / w.sum()


def test_first_compute_ancestor_mismatches_are_cdf_boundary_flips():
    """Whole compute() against the oracle: wherever the two disagree on an ancestor, the selection uniform of that
    candidate lies within 2e-6 of the boundary between the two (adjacent) ancestors in the cumulative table --
    i.e. the only source of divergence is an ulp-level difference of the weights moving a CDF boundary across a draw."""
    scene, model, centre = util.small_case(8, n_scene=5000, n_model=300)
    n, nmax = 100, 220
    d = synth.draws(2, nmax, seed=21)
    g, o = util.make_pair(kld=True, particle_num=n, max_particle_num=nmax, use_hsv=True, oracle_nn=oracle.NN_EXACT_GRID)
    util.set_trans_both(g, o, centre)
    g.injectDraws(*d); o.inject_draws(*d)
    g.setReferenceCloud(model); g.setInputCloud(pcl.PointCloud(scene)); g.compute()
    o.set_reference(model); o.set_input(scene); o.compute()
    cdf = _cdf_after_first_iteration(scene, model, centre, d, n, nmax)
    ga, oa = g.ancestors(), o.ancestors()
    assert len(ga) == len(oa)
    usel = d[0][1]                                   # iteration 1 draws from slot 1
    for i in np.nonzero(ga != oa)[0]:
        lo, hi = sorted((int(ga[i]), int(oa[i])))
        assert np.all(np.diff(cdf[lo:hi]) < 1e-9) or hi - lo == 1, "not neighbours in the table"
        assert abs(float(usel[i]) - cdf[lo]) < 2e-6, (i, usel[i], cdf[lo])


def test_first_compute_exact_when_draws_avoid_cdf_boundaries():
    """The same whole compute() with the selection uniforms nudged away from every boundary of the cumulative table
    (by 1e-4, far above the ulp-level differences of the weights): every ancestor agrees, and the particle set and
    the pose hold the north-star tolerance, 1e-4 m / 1e-4 rad."""
    scene, model, centre = util.small_case(8, n_scene=5000, n_model=300)
    n, nmax = 100, 220
    usel, normals, umot = synth.draws(2, nmax, seed=21)
    cdf = _cdf_after_first_iteration(scene, model, centre, (usel, normals, umot), n, nmax)
    u = usel[1].astype(np.float64)
    k = np.clip(np.searchsorted(cdf, u), 0, len(cdf) - 1)
    near_hi = np.abs(cdf[k] - u) < 1e-4
    near_lo = (k > 0) & (np.abs(cdf[np.maximum(k - 1, 0)] - u) < 1e-4)
    u[near_hi] -= 2e-4
    u[near_lo] += 2e-4
    usel = usel.copy()
    usel[1] = np.clip(u, 0.0, 1.0 - 2.0 ** -24).astype(np.float32)
    d = (usel, normals, umot)
    g, o = util.make_pair(kld=True, particle_num=n, max_particle_num=nmax, use_hsv=True, oracle_nn=oracle.NN_EXACT_GRID)
    util.set_trans_both(g, o, centre)
    g.injectDraws(*d); o.inject_draws(*d)
    g.setReferenceCloud(model); g.setInputCloud(pcl.PointCloud(scene)); g.compute()
    o.set_reference(model); o.set_input(scene); o.compute()
    assert np.array_equal(g.ancestors(), o.ancestors())
    gp, op = g.getParticles(), o.get_particles()
    assert len(gp) == len(op)
    for key in ("x", "y", "z", "roll", "pitch", "yaw"):
        np.testing.assert_allclose(gp[key], op[key], atol=1e-4)
    np.testing.assert_allclose(gp["weight"], op["weight"], rtol=2e-4, atol=1e-9)
    gr, orr = g.getResult(), o.get_result()
    for key in ("x", "y", "z", "roll", "pitch", "yaw"):
        assert abs(float(gr[key]) - float(orr[key])) < 1e-4, key


def test_tracking_follows_moving_object_and_graph_replays():
    objs = synth.default_objects(1)
    frames = [synth.render(f, objs)[0] for f in range(5)]
    pts0, oid0 = synth.render(0, objs)
    model, c = pcl.prepare_model(pcl.PointCloud(synth.model_points(pts0, oid0, 0)), 0.01)
    t = _tracker_for(model, c, kld=True, n=400, nmax=500, use_hsv=True)
    vg = pcl.ApproximateVoxelGrid()
    vg.setLeafSize(0.01, 0.01, 0.01)
    vg.setPassThrough("z", 0.0, 10.0)
    raw, ds = pcl.PointCloud(), pcl.PointCloud()
    errs = []
    for f in range(5):
        raw.upload(frames[f])
        vg.setInputCloud(raw)
        vg.filter(ds)
        t.setInputCloud(ds)
        t.compute()
        r = t.getResult()
        _, cf = synth.object_pose(objs[0], f)
        _, c0 = synth.object_pose(objs[0], 0)
        want = c + (cf - c0)  # the visible-surface centroid moves with the object (approximately)
        errs.append(float(np.linalg.norm([r["x"] - want[0], r["y"] - want[1], r["z"] - want[2]])))
        w = t.getParticles()["weight"].astype(np.float64)
        assert abs(w.sum() - 1.0) < 1e-4
    # object moves 2 cm/frame; the tracker stays on it (the CPU oracle shows the same ~4 cm lag on this sequence:
    # the visible-surface centroid is only an approximate ground truth)
    assert max(errs) < 0.06, errs
    assert t.graphReplays() >= 3           # steady-state frames are CUDA-graph replays


def test_empty_input_and_missing_reference_are_noops():
    scene, model, centre = util.small_case(9, n_scene=1000, n_model=100)
    t = _tracker_for(model, centre, kld=True, n=50, nmax=80)
    t.setInputCloud(pcl.PointCloud(np.zeros(0, dtype=pcl.POINT)))
    t.compute()                              # Tracker::initCompute fails silently on an empty cloud
    assert len(t.getParticles()) == 0
    cloud = pcl.PointCloud(scene)
    t.setInputCloud(cloud)
    t.compute(); t.compute()
    before = t.getParticles().copy()
    rep = t.getResult()
    # a frame whose points are all rejected by the PassThrough: the device-side count is 0 -> no-op
    vg = pcl.ApproximateVoxelGrid()
    vg.setPassThrough("z", 100.0, 200.0)
    vg.setInputCloud(cloud)
    empty = vg.filter()
    t.setInputCloud(empty)
    t.compute()
    after = t.getParticles()
    assert np.array_equal(before.view(np.uint32), after.view(np.uint32))
    assert rep.tobytes() == t.getResult().tobytes()
    # tracker without a reference cloud
    t2 = pcl.KLDAdaptiveParticleFilterOMPTracker(4)
    pcl.configure_like_reference(t2)
    t2.setInputCloud(cloud)
    t2.compute()
    assert len(t2.getParticles()) == 0


def test_fixed_tracker_keeps_particle_count_and_graph_equals_stream():
    scene, model, centre = util.small_case(10, n_scene=4000, n_model=250)
    cloud = pcl.PointCloud(scene)
    res = []
    for no_graph in (False, True):
        t = _tracker_for(model, centre, kld=False, n=256, use_hsv=True, seed=77)
        if no_graph:
            t.setDebugNN(1)  # any debug/timing mode takes the stream-launched path
        t.setInputCloud(cloud)
        for _ in range(4):
            t.compute()
        assert len(t.getParticles()) == 256
        res.append((t.getParticles().copy(), t.getResult().tobytes(), t.graphReplays()))
    assert res[0][2] >= 2 and res[1][2] == 0
    assert np.array_equal(res[0][0].view(np.uint32), res[1][0].view(np.uint32))  # graph replay == plain launches, bit for bit
    assert res[0][1] == res[1][1]


def test_multi_object_batch():
    """C5: 8 objects tracked in one scene; compute_batch == one compute() per tracker."""
    objs = synth.default_objects(8)
    pts, oid = synth.render(0, objs)
    vg = pcl.ApproximateVoxelGrid()
    vg.setLeafSize(0.01)
    vg.setPassThrough("z", 0.0, 10.0)
    vg.setInputCloud(pcl.PointCloud(pts))
    ds = vg.filter()
    trackers, singles = [], []
    for k in range(8):
        raw = synth.model_points(pts, oid, k)
        assert len(raw) > 200
        model, c = pcl.prepare_model(pcl.PointCloud(raw), 0.01)
        for lst in (trackers, singles):
            t = _tracker_for(model, c, kld=True, n=200, nmax=300, use_hsv=True, seed=100 + k)
            t.setInputCloud(ds)
            lst.append(t)
    for _ in range(3):
        pcl.compute_batch(trackers)
        for t in singles:
            t.compute()
    for k, (a, b) in enumerate(zip(trackers, singles)):
        assert np.array_equal(a.getParticles().view(np.uint32), b.getParticles().view(np.uint32))
        r = a.getResult()
        _, c = synth.object_pose(objs[k], 0)
        assert np.linalg.norm([r["x"] - c[0], r["y"] - c[1], r["z"] - c[2]]) < 0.15  # stays on its own object


def test_reset_tracking_redraws_particles():
    scene, model, centre = util.small_case(12, n_scene=2000, n_model=150)
    t = _tracker_for(model, centre, kld=True, n=64, nmax=100)
    t.setInputCloud(pcl.PointCloud(scene))
    t.compute(); t.compute()
    t.resetTracking()
    assert len(t.getParticles()) == 0
    t.compute()
    assert len(t.getParticles()) > 0


def test_device_philox_draws_are_standard():
    """Without injected draws the device generates its own: selection uniforms in [0,1), unit normals."""
    scene, model, centre = util.small_case(13, n_scene=1000, n_model=100)
    t = _tracker_for(model, centre, kld=False, n=4000)
    t.setInitialNoiseCovariance([1.0] * 6)
    t.setQuaternionSampling(False)
    t.setInputCloud(pcl.PointCloud(scene))
    t.initParticles()
    p = t.getParticles()
    z = np.stack([p["x"] - centre[0], p["y"] - centre[1], p["z"] - centre[2], p["roll"], p["pitch"], p["yaw"]], 1).astype(np.float64)
    assert np.all(np.abs(z.mean(0)) < 0.08) and np.all(np.abs(z.std(0) - 1.0) < 0.06)
    assert np.abs(np.corrcoef(z.T) - np.eye(6)).max() < 0.08


def test_result_box_matches_oracle():
    """centroid + PCA oriented bounding box of the model at the result pose (viz_cb, ref :432-466)."""
    scene, model, centre = util.small_case(31, n_scene=3000, n_model=500)
    t = _tracker_for(model, centre, kld=True, n=100, nmax=150)
    t.setInputCloud(pcl.PointCloud(scene))
    t.compute(); t.compute()
    r = t.getResult()
    for z_off in (-0.005, 0.0):
        got = t.getResultBox(z_off)
        want = oracle.result_box(model, [r["x"], r["y"], r["z"], r["roll"], r["pitch"], r["yaw"]], z_off)
        assert got["n"] == want["n"] == len(model)
        np.testing.assert_allclose(got["centroid"], want["centroid"], atol=1e-5)
        np.testing.assert_allclose(got["eigenvalues"], want["eigenvalues"], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(got["extent"], want["extent"], atol=1e-4)
        np.testing.assert_allclose(got["center"], want["center"], atol=1e-4)
        np.testing.assert_allclose(got["axes"], want["axes"], atol=2e-3)
        q = got["quat"].astype(np.float64)
        assert abs(np.linalg.norm(q) - 1.0) < 1e-5
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        np.testing.assert_allclose(R, got["axes"], atol=1e-4)   # qfinal = Quaternionf(eigDx)
    # the published object position is the centroid of the tracked cloud, near the object
    assert np.linalg.norm(t.getResultBox()["centroid"] - centre) < 0.05


def test_cluster_models_feed_the_trackers():
    """C5 end to end: the object models come from Euclidean clustering of the non-table points (the model builder,
    ref: src/create_model.cpp:148-230), one tracker per cluster, all tracked in the same scene."""
    objs = synth.default_objects(8, seed=3)              # (a layout in which no two objects touch)
    pts, oid = synth.render(0, objs)
    keep = (oid >= 0) & np.isfinite(pts["x"])            # what remains after the table plane is cut away
    obj_pts = pts[keep]
    ec = pcl.EuclideanClusterExtraction()
    ec.setClusterTolerance(0.02); ec.setMinClusterSize(200); ec.setMaxClusterSize(25000)
    cloud = pcl.PointCloud(obj_pts)
    ec.setInputCloud(cloud)
    clusters = ec.extract()
    assert len(clusters) == 8
    ids = oid[keep]
    seen = set()
    for k, idx in enumerate(clusters):
        owner = np.unique(ids[idx])
        assert len(owner) == 1                            # a cluster is one object ...
        assert len(idx) == int((ids == owner[0]).sum())   # ... and the whole of it
        seen.add(int(owner[0]))
    assert seen == set(range(8))
    vg = pcl.ApproximateVoxelGrid()
    vg.setLeafSize(0.01); vg.setPassThrough("z", 0.0, 10.0); vg.setInputCloud(pcl.PointCloud(pts))
    ds = vg.filter()
    trackers = []
    for k in range(8):
        model, c = pcl.prepare_model(ec.cluster_cloud(k), 0.01)
        t = _tracker_for(model, c, kld=False, n=200, nmax=200, use_hsv=True, seed=7 + k)
        t.setInputCloud(ds)
        trackers.append(t)
    for _ in range(2):
        pcl.compute_batch(trackers)
    for k, t in enumerate(trackers):
        box = t.getResultBox()
        want = obj_pts[clusters[k]]
        centre = np.array([want["x"].mean(), want["y"].mean(), want["z"].mean()])
        assert np.linalg.norm(box["centroid"] - centre) < 0.03   # every tracker stays on its own cluster


def test_multi_object_batch_one_object_vs_oracle():
    """C5 against the oracle: the first compute() of a batch of 8 trackers (one shared scene), with injected draws; one
    object of the batch is compared with the CPU oracle's compute() on the same inputs (control flow, particle count,
    ancestors up to CDF-boundary flips, matched states 1e-4)."""
    objs = synth.default_objects(8)
    pts, oid = synth.render(0, objs)
    vg = pcl.ApproximateVoxelGrid()
    vg.setLeafSize(0.01)
    vg.setPassThrough("z", 0.0, 10.0)
    vg.setInputCloud(pcl.PointCloud(pts))
    ds = vg.filter()
    scene = ds.to_numpy()
    n, nmax, pick = 120, 260, 3
    trackers, o_pick, model_pick = [], None, None
    for k in range(8):
        model_cloud, c = pcl.prepare_model(pcl.PointCloud(synth.model_points(pts, oid, k)), 0.01)
        g, o = util.make_pair(kld=True, particle_num=n, max_particle_num=nmax, use_hsv=True, oracle_nn=oracle.NN_EXACT_GRID)
        util.set_trans_both(g, o, c)
        d = synth.draws(2, nmax, seed=300 + k)
        g.injectDraws(*d)
        g.setReferenceCloud(model_cloud); g.setInputCloud(ds)
        trackers.append(g)
        if k == pick:
            o.inject_draws(*d)
            o.set_reference(model_cloud.to_numpy()); o.set_input(scene)
            o_pick = o
    pcl.compute_batch(trackers)
    o_pick.compute()
    g = trackers[pick]
    gp, op = g.getParticles(), o_pick.get_particles()
    assert len(gp) == len(op) and len(gp) != n
    same = g.ancestors() == o_pick.ancestors()
    assert same.mean() > 0.97
    for key in ("x", "y", "z", "roll", "pitch", "yaw"):
        np.testing.assert_allclose(gp[key][same], op[key][same], atol=1e-4)


def test_raising_maximum_particle_number_between_computes_keeps_the_kld_stop_rule():
    """setMaximumParticleNum(larger) after the first compute() re-sizes the particle buffers; the table of KLD bounds
    must follow (it is read up to the new capacity).  The resample after the change is compared with the oracle."""
    scene, model, centre = util.small_case(14, n_scene=3000, n_model=200)
    g, o = util.make_pair(kld=True, particle_num=100, max_particle_num=150, use_hsv=False, epsilon=0.02, bin_size=0.02, oracle_nn=oracle.NN_EXACT_GRID)
    util.set_trans_both(g, o, centre)
    cloud = pcl.PointCloud(scene)
    g.setReferenceCloud(model); g.setInputCloud(cloud)
    o.set_reference(model); o.set_input(scene)
    d = synth.draws(2, 600, seed=31)
    g.injectDraws(*d); o.inject_draws(*d)
    g.compute(); o.compute()
    assert len(g.getParticles()) == len(o.get_particles()) <= 150
    g.setMaximumParticleNum(600); o.set_i(oracle.MAX_PARTICLE_NUM, 600)
    # identical state on both sides, then one resample under the new cap
    o.set_particles(g.getParticles()); o.set_result(g.getResult()); o.set_motion(g.getMotion())
    g.resample(1); o.resample(1)
    assert np.array_equal(g.ancestors(), o.ancestors())
    assert len(g.getParticles()) == len(o.get_particles()) > 150


def test_reset_tracking_then_compute_resamples_in_the_same_frame():
    """resetTracking() only clears the particle set (upstream leaves changed_ as it is): the compute() that follows
    re-draws the particles AND resamples them in its first iteration, because an earlier weight() had set changed_."""
    scene, model, centre = util.small_case(15, n_scene=3000, n_model=200)
    g, o = util.make_pair(kld=True, particle_num=80, max_particle_num=200, use_hsv=False, oracle_nn=oracle.NN_EXACT_GRID)
    util.set_trans_both(g, o, centre)
    g.setReferenceCloud(model); g.setInputCloud(pcl.PointCloud(scene))
    o.set_reference(model); o.set_input(scene)
    d = synth.draws(2, 200, seed=33)
    g.injectDraws(*d); o.inject_draws(*d)
    g.compute(); o.compute()
    g.resetTracking(); o.reset()
    g.compute(); o.compute()
    gp, op = g.getParticles(), o.get_particles()
    assert len(gp) == len(op)
    same = g.ancestors() == o.ancestors()
    assert same.mean() > 0.97
    for key in ("x", "y", "z"):
        np.testing.assert_allclose(gp[key][same], op[key][same], atol=1e-4)
