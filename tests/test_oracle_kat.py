"""Pins the CPU oracle (oracle/pft_oracle.cpp) against the derived known-answer vectors of SURVEY.md
Appendix A.10 and against independent numpy restatements of the published formulas.

PARITY UNPINNED at the PCL boundary: the reference ships no golden vectors for the tracking path and
PCL 1.8.0 is not available; these KATs are the strongest pin that exists (see oracle/pft_oracle.cpp)."""
import math

import numpy as np
import pytest

import oracle
from pcl_tracking_b200 import synth


def test_normal_quantile_kat():
    assert abs(oracle.normal_quantile(0.99) - 0.8389129405830074) < 1e-12
    assert oracle.normal_quantile(0.0) == 0.5
    assert oracle.normal_quantile(7.0) == 1.0
    assert oracle.normal_quantile(-7.0) == 0.0
    # it is a normal CDF (upstream misnomer): cross-check against erf
    for u in (-2.5, -1.0, -0.3, 0.2, 0.99, 1.7, 3.3):
        assert abs(oracle.normal_quantile(u) - 0.5 * (1 + math.erf(u / math.sqrt(2)))) < 2e-6


def test_kl_bound_kat():
    expect = {2: 4.037, 3: 7.978, 5: 14.901, 10: 30.534, 20: 59.675, 50: 142.605, 100: 276.402, 150: 408.116, 200: 538.765}
    for k, v in expect.items():
        assert abs(oracle.kl_bound(k, 0.99, 0.2) - v) < 5e-3, (k, oracle.kl_bound(k, 0.99, 0.2))


def test_div_table_kat():
    expect = [1044480, 522240, 348160, 261120, 208896, 174080, 149211, 130560, 116053, 104448, 94953, 87040, 80345, 74606, 69632]
    assert [oracle.div_table(i) for i in range(1, 16)] == expect
    assert oracle.div_table(255) == 4096
    assert oracle.div_table(0) == 0


def test_rgb2hsv_kat():
    cases = {(255, 0, 0): (0, 255, 255), (0, 255, 0): (60, 255, 255), (0, 0, 255): (120, 255, 255), (200, 120, 40): (15, 203, 200),
             (10, 10, 10): (0, 0, 10), (255, 255, 255): (0, 0, 255), (0, 0, 0): (0, 0, 0)}
    for rgb, hsv in cases.items():
        assert oracle.rgb2hsv(*rgb) == hsv, rgb


def test_rgb2hsv_against_float_formula():
    rng = np.random.default_rng(0)
    for r, g, b in rng.integers(0, 256, (300, 3)):
        h, s, v = oracle.rgb2hsv(int(r), int(g), int(b))
        mx, mn = max(r, g, b), min(r, g, b)
        assert v == mx
        if mx > 0:
            assert abs(s - 255.0 * (mx - mn) / mx) <= 1.0
        if mx - mn > 8:
            d = float(mx - mn)
            if mx == r:
                hf = 60.0 * (float(g) - float(b)) / d
            elif mx == g:
                hf = 120.0 + 60.0 * (float(b) - float(r)) / d
            else:
                hf = 240.0 + 60.0 * (float(r) - float(g)) / d
            hf = (hf % 360.0) / 2.0
            dh = abs(h - hf)
            assert min(dh, 180 - dh) <= 1.0, (r, g, b, h, hf)


def _rot_zyx(roll, pitch, yaw):
    cr, sr, cp, sp, cy, sy = math.cos(roll), math.sin(roll), math.cos(pitch), math.sin(pitch), math.cos(yaw), math.sin(yaw)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def test_particle_to_matrix_is_Rz_Ry_Rx():
    rng = np.random.default_rng(1)
    for _ in range(50):
        s = rng.uniform(-1.2, 1.2, 6)
        m = oracle.particle_to_matrix(s)
        np.testing.assert_allclose(m[:, :3], _rot_zyx(s[3], s[4], s[5]), atol=3e-7)
        np.testing.assert_allclose(m[:, 3], s[:3].astype(np.float32), atol=0)
        back = oracle.matrix_to_particle(m)
        np.testing.assert_allclose([back["roll"], back["pitch"], back["yaw"]], s[3:], atol=2e-6)


def test_identity_matrix():
    m = oracle.particle_to_matrix([0, 0, 0, 0, 0, 0])
    np.testing.assert_array_equal(m, np.eye(4, dtype=np.float32)[:3])


def test_distance_and_hsv_coherence_formulas():
    a, b = (0.1, 0.2, 0.3), (0.13, 0.16, 0.3)
    d2 = np.float32(np.float32(np.float32(0.1) - np.float32(0.13)) ** 2 + np.float32(np.float32(0.2) - np.float32(0.16)) ** 2)
    d = float(np.sqrt(np.float32(d2)))
    assert abs(oracle.distance_coherence(a, b, 1.0) - 1.0 / (1.0 + d * d)) < 1e-12
    assert abs(oracle.distance_coherence(a, b, 3.0) - 1.0 / (1.0 + 3.0 * d * d)) < 1e-12
    # identical colours -> 1; G/B swap quirk: (R,G,B)=(255,0,0) vs (255,0,0) trivially 1
    c = (255 << 24) | (200 << 16) | (120 << 8) | 40
    assert oracle.hsv_coherence(c, c, 0.1) == 1.0
    # upstream passes (R, B, G): colour (r=200,g=120,b=40) is converted as RGB2HSV(200, 40, 120)
    h1, s1, _ = oracle.rgb2hsv(200, 40, 120)
    c2 = (255 << 24) | (10 << 16) | (200 << 8) | 30
    h2, s2, _ = oracle.rgb2hsv(10, 30, 200)
    fh1, fh2, fs1, fs2 = np.float32(h1) / np.float32(180), np.float32(h2) / np.float32(180), np.float32(s1) / np.float32(255), np.float32(s2) / np.float32(255)
    hd = abs(fh1 - fh2)
    hd2 = abs(np.float32(1.0) + min(fh1, fh2) - max(fh1, fh2))
    hdiff = np.float32(min(hd, hd2)) ** 2
    sdiff = np.float32(fs1 - fs2) ** 2
    want = 1.0 / (1.0 + 0.1 * float(np.float32(hdiff + sdiff)))
    assert abs(oracle.hsv_coherence(c, c2, 0.1) - want) < 1e-7


def test_passthrough_semantics():
    pts = oracle.make_points([[0, 0, 0.0], [0, 0, 10.0], [0, 0, 10.0001], [0, 0, -1e-6], [np.nan, 0, 1], [0, np.inf, 1], [1, 2, 3]])
    out = oracle.passthrough(pts, 2, 0.0, 10.0)
    assert [tuple(p)[:3] for p in out] == [(0, 0, 0), (0, 0, 10), (1, 2, 3)]  # inclusive limits, order kept, non-finite dropped


def test_approx_voxel_grid_pcl_small_case():
    # two points in voxel (0,0,0), one in voxel (5,0,0); no hash collision -> two centroids, flush order = slot order
    pts = oracle.make_points([[0.001, 0.002, 0.003], [0.003, 0.004, 0.005], [0.051, 0.001, 0.001]],
                             rgba=[(10 << 16) | (20 << 8) | 30, (30 << 16) | (40 << 8) | 50, (1 << 16) | (2 << 8) | 3])
    out = oracle.approx_voxel_grid_pcl(pts, 0.01)
    assert len(out) == 2
    # slot of voxel (0,0,0) is 0 -> emitted first
    np.testing.assert_allclose([out[0]["x"], out[0]["y"], out[0]["z"]], [0.002, 0.003, 0.004], rtol=1e-6)
    assert out[0]["rgba"] == (20 << 16) | (30 << 8) | 40
    np.testing.assert_allclose(out[1]["x"], 0.051, rtol=1e-6)


def test_approx_voxel_grid_pcl_collision_emits_duplicates():
    # voxels (0,0,0) and (512,0,0)... hash = ix*7171 & 511: ix=512 -> 0: same slot, different voxel -> eviction
    pts = oracle.make_points([[0.001, 0.001, 0.001], [5.121, 0.001, 0.001], [0.002, 0.002, 0.002]])
    out = oracle.approx_voxel_grid_pcl(pts, 0.01)
    assert len(out) == 3  # voxel (0,0,0) is emitted twice (partial centroids): "approximate"


def test_voxel_grid_pcl_order_and_centroids():
    rng = np.random.default_rng(2)
    xyz = rng.uniform(-0.05, 0.05, (500, 3))
    pts = oracle.make_points(xyz, rgba=rng.integers(0, 1 << 24, 500))
    out = oracle.voxel_grid_pcl(pts, 0.02)
    ex = oracle.voxel_grid_exact(pts, 0.02, -1, 0, 0)
    assert len(out) == len(ex)
    # same voxel set, same centroids (float vs double accumulation: 1e-6), different order
    key = lambda a: np.lexsort((np.floor(a["z"] * 50), np.floor(a["y"] * 50), np.floor(a["x"] * 50)))
    a, b = out[key(out)], ex[key(ex)]
    for k in "xyz":
        np.testing.assert_allclose(a[k], b[k], atol=1e-6)
    # PCL order: ascending voxel index, x fastest then y then z
    ix, iy, iz = (np.floor(out[k] * np.float32(50.0)).astype(int) for k in "xyz")
    lin = (ix - ix.min()) + (iy - iy.min()) * 100 + (iz - iz.min()) * 10000
    assert np.all(np.diff(lin) > 0)


def test_voxel_grid_exact_first_appearance_order_and_passthrough():
    pts = oracle.make_points([[0.051, 0, 1.0], [0.001, 0, 1.0], [0.052, 0, 1.0], [0.3, 0, 11.0], [np.nan, 0, 1.0]])
    out = oracle.voxel_grid_exact(pts, 0.01, 2, 0.0, 10.0)
    assert len(out) == 2
    np.testing.assert_allclose(out["x"], [0.0515, 0.001], rtol=1e-6)


def test_octree_approx_nearest_properties():
    rng = np.random.default_rng(3)
    pts = oracle.make_points(rng.uniform(0, 0.5, (2000, 3)))
    # querying exactly at data points returns a point at distance 0 in the same leaf
    idx, d2 = oracle.octree_approx_nearest(pts, 0.01, np.stack([pts["x"], pts["y"], pts["z"]], 1)[:200])
    assert np.all(d2 == 0)
    q = rng.uniform(0, 0.5, (500, 3)).astype(np.float32)
    idx, d2 = oracle.octree_approx_nearest(pts, 0.01, q)
    assert np.all(idx >= 0)
    xyz = np.stack([pts["x"], pts["y"], pts["z"]], 1)
    true_d2 = ((xyz[None, :, :] - q[:, None, :]) ** 2).sum(-1).min(1)
    got = ((xyz[idx] - q) ** 2).sum(-1)
    np.testing.assert_allclose(got, d2, rtol=1e-5)
    assert np.all(got >= true_d2 - 1e-9)        # never better than the true nearest
    assert np.mean(got <= true_d2 * 1.0001 + 1e-12) > 0.3  # and often equal to it


def _tracker(kld=True, **kw):
    t = oracle.Tracker(kld=kld)
    oracle.configure_like_reference(t, **kw)
    return t


def test_normalize_weight():
    t = _tracker()
    p = oracle.make_particles(np.zeros((5, 6)), [-10.0, -4.0, 0.0, -7.0, -10.0])
    t.set_particles(p)
    t.normalize()
    w = t.get_particles()["weight"].astype(np.float64)
    assert abs(w.sum() - 1.0) < 1e-6
    assert w[2] == 0.0                                   # zero raw weight stays zero
    e = np.exp(1.0 - 15.0 * (np.array([-10.0, -4.0, -7.0, -10.0]) + 10.0) / 6.0)
    np.testing.assert_allclose(w[[0, 1, 3, 4]], e / e.sum(), rtol=1e-6)
    assert t.fit_ratio() == -10.0
    # all equal -> uniform
    t.set_particles(oracle.make_particles(np.zeros((4, 6)), [-3.0] * 4))
    t.normalize()
    np.testing.assert_allclose(t.get_particles()["weight"], 0.25)


def test_update_weighted_mean_and_motion():
    t = _tracker()
    rng = np.random.default_rng(4)
    s = rng.normal(0, 1, (50, 6)).astype(np.float32)
    w = rng.random(50).astype(np.float32)
    w /= w.sum()
    t.set_particles(oracle.make_particles(s, w))
    rep0 = np.zeros(1, dtype=oracle.PARTICLE)
    rep0["x"], rep0["one"] = 0.5, 1.0
    t.set_result(rep0[0])
    t.update()
    r = t.get_result()
    want = (s.astype(np.float64) * w[:, None].astype(np.float64)).sum(0)
    np.testing.assert_allclose([r[k] for k in ("x", "y", "z", "roll", "pitch", "yaw")], want, atol=2e-6)
    assert abs(r["weight"] - 1 / 50) < 1e-9
    assert abs(t.get_motion()["x"] - (r["x"] - 0.5)) < 1e-7


def test_exact_grid_equals_brute_force():
    scene, model, centre = synth.uniform_surface_scene(6000, 400, seed=5)
    ws = []
    nns = []
    for mode in (oracle.NN_EXACT_BRUTE, oracle.NN_EXACT_GRID):
        t = _tracker(kld=False, particle_num=24, nn_mode=mode)
        rng = np.random.default_rng(6)
        s = np.zeros((24, 6), dtype=np.float32)
        s[:, :3] = centre + rng.normal(0, 0.02, (24, 3))
        s[:, 3:] = rng.normal(0, 0.1, (24, 3))
        t.set_reference(model); t.set_input(scene); t.set_particles(oracle.make_particles(s, np.full(24, 1 / 24)))
        t.weight(keep_nn=True)
        ws.append(t.raw_weights())
        nns.append([t.nn(p, len(model)) for p in range(24)])
    np.testing.assert_array_equal(ws[0], ws[1])
    for (bi, bd), (gi, gd) in zip(*nns):
        m = bd.astype(np.float64) < 0.1 * 0.1
        np.testing.assert_array_equal(bi[m], gi[m])
        np.testing.assert_array_equal(bd[m], gd[m])


def test_alias_table_preserves_weights():
    t = _tracker(kld=False, particle_num=200)
    t.set_i(oracle.SAMPLER, oracle.SAMPLER_ALIAS_PCL)
    rng = np.random.default_rng(8)
    w = rng.random(200).astype(np.float32) ** 3
    w /= w.sum()
    s = np.zeros((200, 6), dtype=np.float32)
    s[:, 0] = np.arange(200)
    t.set_particles(oracle.make_particles(s, w))
    t.set_vec6(oracle.STEP_COV, [0] * 6)
    t.set_i(oracle.QUAT_SAMPLE, 0)
    n_draw = 200
    counts = np.zeros(200)
    for rep in range(60):
        t.set_particles(oracle.make_particles(s, w))
        usel, normals, umot = synth.draws(1, n_draw, seed=100 + rep)
        t.inject_draws(usel, normals, umot)
        t.resample(0)
        anc = t.ancestors()
        a = anc[anc >= 0]
        np.add.at(counts, a, 1)
        got = t.get_particles()
        np.testing.assert_array_equal(got["x"][1:], a.astype(np.float32))  # slot 0 = representative
    freq = counts / counts.sum()
    assert np.abs(freq - w).max() < 0.02


def test_cdf_sampler_is_inverse_cdf():
    t = _tracker(kld=False, particle_num=50)
    t.set_i(oracle.SAMPLER, oracle.SAMPLER_CDF)
    w = np.zeros(50, dtype=np.float32)
    w[[3, 10, 40]] = [0.25, 0.5, 0.25]
    s = np.zeros((50, 6), dtype=np.float32)
    t.set_particles(oracle.make_particles(s, w))
    usel = np.array([[0.0, 0.1, 0.2499, 0.2501, 0.5, 0.7499, 0.7501, 0.99] + [0.5] * 42], dtype=np.float32)
    t.inject_draws(usel, np.zeros((1, 50, 6), np.float32), np.ones((1, 50), np.float32))
    t.resample(0)
    assert t.ancestors()[1:8].tolist() == [3, 3, 10, 10, 10, 40, 40]


def test_kld_resample_respects_bound():
    scene, model, centre = synth.uniform_surface_scene(3000, 200, seed=9)
    for eps, bins in ((0.2, 0.1), (0.05, 0.05), (0.02, 0.02)):
        t = _tracker(kld=True, particle_num=100, max_particle_num=5000)
        t.set_d(oracle.EPSILON, eps)
        t.set_vec6(oracle.BIN_SIZE, [bins] * 6)
        s = np.zeros((100, 6), dtype=np.float32)
        s[:, :3] = centre
        t.set_particles(oracle.make_particles(s, np.full(100, 0.01)))
        usel, normals, umot = synth.draws(1, 5000, seed=11)
        t.inject_draws(usel, normals, umot)
        t.resample(0)
        p = t.get_particles()
        n = len(p)
        st = np.stack([p[k] for k in ("x", "y", "z", "roll", "pitch", "yaw")], 1)
        k = len({tuple(r) for r in np.trunc(st / np.float32(bins)).astype(np.int64)})
        assert 1 <= n <= 5000
        if n < 5000:
            assert k >= 2 and n >= oracle.kl_bound(k, 0.99, eps)
            # one particle fewer would not have satisfied the stop rule
            km = len({tuple(r) for r in np.trunc(st[:-1] / np.float32(bins)).astype(np.int64)})
            assert km < 2 or (n - 1) < oracle.kl_bound(km, 0.99, eps)


def test_compute_first_iteration_skips_resample_and_tracks():
    scene, model, centre = synth.uniform_surface_scene(60000, 1500, seed=12)
    t = _tracker(kld=True, particle_num=400, max_particle_num=500, nn_mode=oracle.NN_EXACT_GRID, use_hsv=False)
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = centre + np.array([0.03, -0.02, 0.01], dtype=np.float32)
    t.set_trans(m[:3])
    t.set_reference(model)
    t.set_input(scene)
    for f in range(4):
        t.inject_draws(*synth.draws(2, 500, seed=13 + f))
        t.compute()
        if f == 0:
            assert len(t.get_particles()) != 400  # iteration 1 resampled (KLD changes the count), iteration 0 did not
    r = t.get_result()
    err = np.linalg.norm([r["x"] - centre[0], r["y"] - centre[1], r["z"] - centre[2]])
    assert err < 0.01, err  # starts 3.7 cm off, converges to a few mm
    assert abs(t.get_particles()["weight"].astype(np.float64).sum() - 1.0) < 1e-4


def test_empty_input_is_noop():
    t = _tracker(kld=True, particle_num=10, max_particle_num=20)
    t.set_reference(oracle.make_points([[0, 0, 0]]))
    t.set_input(oracle.make_points(np.zeros((0, 3))))
    t.compute()
    assert len(t.get_particles()) == 0


def test_particle_sample_modes():
    z = np.array([0.5, -1.0, 2.0, 0.3, -0.2, 0.1], dtype=np.float32)
    cov = [1e-4] * 3 + [4e-4] * 3
    p = oracle.particle_sample([1, 2, 3, 0.1, 0.2, 0.3], [0] * 6, cov, z, quat_mode=0)
    np.testing.assert_allclose([p["x"], p["y"], p["z"]], [1.005, 1.99, 3.02], atol=1e-6)
    np.testing.assert_allclose([p["roll"], p["pitch"], p["yaw"]], [0.106, 0.196, 0.302], atol=1e-6)
    q = oracle.particle_sample([1, 2, 3, 0.1, 0.2, 0.3], [0] * 6, cov, z, quat_mode=1)
    np.testing.assert_allclose([q["x"], q["y"], q["z"]], [1.005, 1.99, 3.02], atol=1e-6)
    # small-angle: quaternion noise (a,b,c)*sqrt(0.2862*cov) rotates by ~2*(a,b,c) about x,y,z
    R0 = _rot_zyx(0.1, 0.2, 0.3)
    ang = 2 * np.sqrt(0.2862 * 4e-4) * z[3:]
    Rs = _rot_zyx(ang[0], ang[1], ang[2])
    np.testing.assert_allclose(_rot_zyx(q["roll"], q["pitch"], q["yaw"]), Rs @ R0, atol=2e-4)
    # zero noise keeps the pose
    q0 = oracle.particle_sample([1, 2, 3, 0.1, 0.2, 0.3], [0] * 6, [0] * 6, z, quat_mode=1)
    np.testing.assert_allclose([q0["roll"], q0["pitch"], q0["yaw"]], [0.1, 0.2, 0.3], atol=1e-6)


def test_from_pointcloud2_known_layout():
    """fromPCLPointCloud2 restatement: the 32-byte PointXYZRGBA record of kinect2_bridge (x,y,z at 0,4,8; rgba at 16)."""
    import struct
    recs = [(1.0, -2.5, 3.25, 0x11223344), (float("nan"), 0.0, 7.0, 0xff000001)]
    blob = b"".join(struct.pack("<fff4xI12x", *r) for r in recs) + b"\xaa" * 16  # row padding
    out = oracle.from_pointcloud2(blob, 2, 1, 32, 80)
    assert out["x"][0] == 1.0 and out["y"][0] == -2.5 and out["z"][0] == 3.25 and out["rgba"][0] == 0x11223344
    assert np.isnan(out["x"][1]) and out["z"][1] == 7.0 and out["rgba"][1] == 0xff000001
    xyz = oracle.from_pointcloud2(blob, 2, 1, 32, 80, off_rgb=-1)
    assert not xyz["rgba"].any()


def test_result_box_known_answer():
    """PCA oriented bounding box of viz_cb (ref: src/auto_tracking.cpp:432-466) on a cloud whose box is known: an
    8 x 4 x 2 cm grid-filled cuboid, rotated and translated by the result pose."""
    g = np.stack(np.meshgrid(np.linspace(-0.04, 0.04, 17), np.linspace(-0.02, 0.02, 9), np.linspace(-0.01, 0.01, 5), indexing="ij"), -1).reshape(-1, 3)
    model = oracle.make_points(g.astype(np.float32))
    state = [0.3, -0.2, 1.1, 0.4, -0.3, 0.9]
    box = oracle.result_box(model, state, z_offset=-0.005)
    np.testing.assert_allclose(box["centroid"], [0.3, -0.2, 1.095], atol=2e-6)
    np.testing.assert_allclose(box["extent"], [0.02, 0.04, 0.08], atol=1e-5)       # ascending eigenvalues: thinnest axis first
    np.testing.assert_allclose(box["center"], box["centroid"], atol=1e-5)          # symmetric cloud: box centre = centroid
    m = oracle.particle_to_matrix(state)[:, :3]
    # the box axes are the rotated model axes (z, y, x in ascending order of variance), up to sign
    for col, axis in enumerate((2, 1, 0)):
        assert abs(abs(float(box["axes"][:, col] @ m[:, axis])) - 1.0) < 1e-4
    assert abs(np.linalg.det(box["axes"].astype(np.float64)) - 1.0) < 1e-5        # right-handed: col2 = col0 x col1


def test_euclidean_clusters_known_answer():
    """EuclideanClusterExtraction restatement (ref: src/create_model.cpp:169-179): two chains of points 1.5 cm apart
    (connected under a 2 cm tolerance), a third one too short for min_size, a NaN point and an isolated point."""
    def chain(x0, y, n):
        return [(x0 + 0.015 * k, y, 1.0) for k in range(n)]
    xyz = np.array(chain(0.0, 0.0, 30) + [(np.nan, 0.0, 1.0)] + chain(0.0, 0.5, 40) + chain(0.0, 1.0, 5) + [(5.0, 5.0, 5.0)], dtype=np.float32)
    labels, sizes = oracle.euclidean_clusters(oracle.make_points(xyz), 0.02, 10, 1000)
    assert sizes.tolist() == [40, 30]                       # largest first
    assert (labels[31:71] == 0).all() and (labels[:30] == 1).all()
    assert labels[30] == -1 and (labels[71:] == -1).all()   # NaN, too-small chain, isolated point
    # the tolerance is strict: points exactly 2 cm apart (in fp32) are not neighbours
    pair = oracle.make_points(np.array([(0.0, 0.0, 0.0), (0.02, 0.0, 0.0)], dtype=np.float32))
    assert (np.float32(0.02) * np.float32(0.02) < np.float32(0.02 * 0.02)) == (oracle.euclidean_clusters(pair, 0.02, 2, 10)[1].tolist() == [2])
    # max_size drops oversized components
    assert oracle.euclidean_clusters(oracle.make_points(xyz), 0.02, 10, 35)[1].tolist() == [30]


def test_segment_plane_known_answer():
    """SACSegmentation(PLANE, RANSAC) restatement (ref: src/create_model_planar_segmentation.cpp:157-163): a noisy plane
    z = 0.3 x - 0.2 y + 1 under a blob of object points."""
    rng = np.random.default_rng(0)
    n = 4000
    xy = rng.uniform(-1, 1, (n, 2))
    plane = np.c_[xy, 0.3 * xy[:, 0] - 0.2 * xy[:, 1] + 1.0 + rng.normal(0, 0.003, n)]
    blob = rng.uniform(-0.3, 0.3, (1200, 3)) + [0, 0, 1.5]
    pts = oracle.make_points(np.r_[plane, blob].astype(np.float32))
    samples = rng.integers(0, len(pts), (1001, 3))
    samples[0] = (5, 5, 9)                                      # a degenerate draw: skipped, not an iteration
    r = oracle.segment_plane(pts, samples, 0.015, 1000, 0.99, True)
    want = np.array([-0.3, 0.2, 1.0, -1.0]) / np.linalg.norm([0.3, 0.2, 1.0])
    c = r["coeff"].astype(np.float64)
    c = c if c[2] > 0 else -c
    np.testing.assert_allclose(c, want, atol=2e-3)
    assert r["inliers"][:n].mean() > 0.995 and r["inliers"][n:].mean() < 0.12
    # the adaptive stop: with ~77 % inliers k = log(0.01) / log(1 - 0.77^3) ~ 8 iterations, far below the 1000 allowed
    assert 3 <= r["iterations"] <= 40
    # without refinement the coefficients are those of the winning sample
    r0 = oracle.segment_plane(pts, samples, 0.015, 1000, 0.99, False)
    assert r0["best"] == r["best"] and np.array_equal(r0["coeff"], r["ransac_coeff"])


# ------------------------------------------------------------------ change detector
def _cd_cloud(xyz):
    a = np.zeros(len(xyz), dtype=oracle.POINT)
    xyz = np.asarray(xyz, dtype=np.float32).reshape(-1, 3)
    a["x"], a["y"], a["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    return a


def test_change_detector_known_answers():
    """OctreePointCloudChangeDetector as testChangeDetection drives it: everything is new for the first cloud, nothing
    for a repeated cloud, only the voxels the PREVIOUS cloud lacked afterwards (A, B, A reports A's own voxels again),
    and new voxels with fewer than min_points points are ignored."""
    # (the first point ever added ends up on the upper edge of its voxel: the initial box [p0 - res/2, p0 + res/2] is
    #  widened to two voxels per axis around its centre -- the other points are placed well inside their voxels)
    a = _cd_cloud([(0, 0, 0), (0.103, 0.003, 0.003), (0.104, 0.003, 0.003)])
    b = _cd_cloud([(0, 0, 0), (0.503, 0.303, 0.203), (0.504, 0.303, 0.203), (0.5045, 0.303, 0.203)])
    far = _cd_cloud([(-3.003, 2.003, 7.503), (0, 0, 0)])   # forces the bounding box to double several times
    assert oracle.change_detector_sequence([a, a, b, a, a], 0.01, 0).tolist() == [3, 0, 3, 2, 0]
    assert oracle.change_detector_sequence([a, b, far, a], 0.01, 0).tolist() == [3, 3, 1, 2]
    assert oracle.change_detector_sequence([a, b], 0.01, 2).tolist() == [2, 3]
    assert oracle.change_detector_sequence([a, b], 0.01, 4).tolist() == [0, 0]
    empty = _cd_cloud(np.zeros((0, 3)))
    assert oracle.change_detector_sequence([empty, a, empty, a], 0.01, 0).tolist() == [0, 3, 0, 3]
    nan = _cd_cloud([(np.nan, 0, 0), (0.003, 0.003, 0.003)])
    assert oracle.change_detector_sequence([nan], 0.01, 0).tolist() == [1]


@pytest.mark.parametrize("res,min_points", [(0.01, 0), (0.02, 2), (0.05, 3)])
def test_change_detector_equals_voxel_set_difference(res, min_points):
    """The double-buffered octree is a voxel set difference on the lattice the first point anchors (origin = first
    point - resolution + eps/2, growth shifts the box by whole voxels): checked against numpy on random clouds that
    overlap partly and make the box grow in every direction."""
    rng = np.random.default_rng(int(res * 1000) + min_points)
    clouds = []
    for k in range(6):
        centre = rng.uniform(-0.3, 0.3, 3) * (1 + k)
        n = int(rng.integers(200, 500))
        pts = centre + rng.uniform(-0.15, 0.15, (n, 3))
        if k:   # keep a part of the previous cloud
            prev = np.stack([clouds[-1]["x"], clouds[-1]["y"], clouds[-1]["z"]], axis=1).astype(np.float64)
            pts = np.concatenate([pts, prev[: len(prev) // 2]])
        clouds.append(_cd_cloud(pts))
    got = oracle.change_detector_sequence(clouds, res, min_points)
    p0 = np.array([clouds[0]["x"][0], clouds[0]["y"][0], clouds[0]["z"][0]], dtype=np.float64)
    eps = float(np.finfo(np.float32).eps)
    origin = p0 - res / 2 - ((2 * res - eps) - res) / 2
    prev_keys = set()
    for k, c in enumerate(clouds):
        xyz = np.stack([c["x"], c["y"], c["z"]], axis=1).astype(np.float64)
        keys = np.floor((xyz - origin) / res).astype(np.int64)
        uniq, inv, counts = np.unique(keys, axis=0, return_inverse=True, return_counts=True)
        want = 0
        for u, cnt in zip(map(tuple, uniq), counts):
            if u not in prev_keys and cnt >= min_points:
                want += int(cnt)
        assert got[k] == want, (k, got[k], want)
        prev_keys = set(map(tuple, uniq))


def test_tracker_with_change_detector_freezes_on_a_static_scene():
    """weight() with use_change_detector_: a test every `interval` calls; on a static scene the second test finds
    nothing new, changed_ drops to false, the weights are not recomputed and resample()/update() are skipped."""
    rng = np.random.default_rng(5)
    scene = oracle.make_points(rng.uniform(-0.2, 0.2, (3000, 3)).astype(np.float32) + np.float32([0, 0, 1.0]), rng.integers(0, 1 << 24, 3000).astype(np.uint32))
    model = oracle.make_points(rng.uniform(-0.05, 0.05, (200, 3)).astype(np.float32), rng.integers(0, 1 << 24, 200).astype(np.uint32))
    t = _tracker(kld=False, particle_num=40, nn_mode=oracle.NN_EXACT_GRID)
    t.set_change_detector(True, interval=1, min_points=1, resolution=0.02)
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = [0, 0, 1.0]
    t.set_trans(m[:3])
    t.set_reference(model)
    t.set_input(scene)
    t.set_i(oracle.ITERATION_NUM, 1)
    infos = []
    for f in range(5):
        t.compute()
        infos.append(t.change_detector_info())
    # call 1: counter 0 -> test (everything new) -> counter 1; call 2: counter 1 -> 0, weights computed; call 3: test
    assert [i["tests"] for i in infos[:2]] == [1, 1] and infos[0]["changed"] and infos[1]["changed"]
    assert infos[0]["last_found"] > 0
    # later tests only see what a slightly different crop box adds; once the particle set stops moving nothing is new
    later = [i for i in infos[2:] if not i["changed"]]
    for i in later:
        assert i["last_found"] == 0 and i["counter"] == 0


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_euclidean_clusters_equal_scipy_connected_components(seed):
    """Independent check of the clustering restatement: the clusters are the connected components of the graph that
    joins points closer than the tolerance (scipy cKDTree + csgraph), filtered by size, largest first."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    blobs = [rng.normal(c, s, (m, 3)) for c, s, m in (((0, 0, 1.0), 0.03, 400), ((0.5, 0.1, 1.1), 0.02, 250), ((-0.4, 0.3, 0.9), 0.015, 120),
                                                      ((0.1, -0.6, 1.3), 0.05, 90))]
    xyz = np.concatenate(blobs + [rng.uniform(-1, 1, (150, 3)) + np.array([0, 0, 1.0])]).astype(np.float32)
    xyz = xyz[rng.permutation(len(xyz))]
    tol, mn, mx = 0.02, 20, 380
    labels, sizes = oracle.euclidean_clusters(oracle.make_points(xyz), tol, mn, mx)
    pairs = cKDTree(xyz.astype(np.float64)).query_pairs(tol, output_type="ndarray")
    n = len(xyz)
    g = coo_matrix((np.ones(len(pairs)), (pairs[:, 0], pairs[:, 1])), shape=(n, n))
    _, comp = connected_components(g, directed=False)
    comp_sizes = np.bincount(comp)
    kept = [c for c in range(len(comp_sizes)) if mn <= comp_sizes[c] <= mx]
    assert sorted(sizes.tolist(), reverse=True) == sizes.tolist()
    assert sorted((int(comp_sizes[c]) for c in kept), reverse=True) == sizes.tolist()
    assert len(kept) >= 2
    # same partition: every kept component carries exactly one label, everything else is unlabelled
    for c in kept:
        lab = np.unique(labels[comp == c])
        assert len(lab) == 1 and lab[0] >= 0 and sizes[lab[0]] == comp_sizes[c]
    assert (labels[~np.isin(comp, kept)] == -1).all()


def _py_octree_approx_nearest(xyz, res, queries):
    """Second, flat restatement of OctreePointCloud::addPointsFromInputCloud + approxNearestSearch in
    numpy/Python: no tree, only the voxel keys -- the bounding box grows point by point as upstream's does (every
    growth re-bases the keys of the points already in), and the greedy descent walks the sets of occupied key prefixes."""
    eps = float(np.finfo(np.float32).eps)
    mn = np.zeros(3); mx = np.zeros(3)
    depth, defined = 0, False
    keys = np.zeros((len(xyz), 3), dtype=np.int64)
    n_in = 0
    for i, p in enumerate(xyz.astype(np.float64)):
        while True:
            lo = p < mn
            up = p >= mx
            if not (lo.any() or up.any() or not defined):
                break
            if defined:
                side = float(1 << depth) * res
                for d in range(3):
                    if not up[d]:
                        mn[d] -= side
                        keys[:n_in, d] += 1 << depth      # the old root becomes the upper child along this axis
                depth += 1
                mx = mn + (float(1 << depth) * res - eps)
            else:
                mn = p - res / 2
                mx = p + res / 2
                mk = np.ceil((mx - mn) / res).astype(np.int64)
                depth = int(max(min(30.0, np.ceil(np.log2(float(max(mk.max(), 2))) - eps)), 0.0))
                side = float(1 << depth) * res - eps
                over = (side - (mx - mn)) / 2.0
                mn = mn - over
                mx = mx + over
                defined = True
        keys[i] = ((p - mn) / res).astype(np.int64)   # truncation, as the unsigned cast
        n_in = i + 1
    occupied = [set() for _ in range(depth + 1)]          # level l: key prefixes of l bits
    leaves = {}
    for i, k in enumerate(map(tuple, keys)):
        for l in range(depth + 1):
            occupied[l].add(tuple(int(c) >> (depth - l) for c in k))
        leaves.setdefault(k, []).append(i)
    out_idx, out_d2 = [], []
    f32 = np.float32
    for q in queries.astype(np.float32):
        key = (0, 0, 0)
        for l in range(1, depth + 1):
            cell = res * float(1 << (depth - l))
            best, best_key = None, None
            for ci in range(8):
                nk = ((key[0] << 1) + ((ci >> 2) & 1), (key[1] << 1) + ((ci >> 1) & 1), (key[2] << 1) + (ci & 1))
                if nk not in occupied[l]:
                    continue
                c = [f32((float(nk[d]) + 0.5) * cell + mn[d]) for d in range(3)]
                dx, dy, dz = f32(c[0] - q[0]), f32(c[1] - q[1]), f32(c[2] - q[2])
                dist = float(f32(f32(f32(dx * dx) + f32(dy * dy)) + f32(dz * dz)))
                if best is None or dist < best:
                    best, best_key = dist, nk
            key = best_key
        smallest, idx, d2 = None, -1, f32(0)
        for i in leaves[key]:
            p = xyz[i]
            dx, dy, dz = f32(p[0] - q[0]), f32(p[1] - q[1]), f32(p[2] - q[2])
            sd = f32(f32(f32(dx * dx) + f32(dy * dy)) + f32(dz * dz))
            if smallest is None or float(sd) < smallest:
                smallest, idx, d2 = float(sd), i, sd
        out_idx.append(idx)
        out_d2.append(d2)
    return np.array(out_idx, dtype=np.int32), np.array(out_d2, dtype=np.float32)


@pytest.mark.parametrize("seed,res", [(11, 0.01), (12, 0.02), (13, 0.007)])
def test_octree_approx_nearest_equals_flat_restatement(seed, res):
    rng = np.random.default_rng(seed)
    # an object-sized cloud entered in an order that makes the box grow in every direction
    xyz = (rng.normal(0, 0.06, (700, 3)) + np.array([0.3, -0.2, 1.1])).astype(np.float32)
    q = (rng.normal(0, 0.07, (150, 3)) + np.array([0.3, -0.2, 1.1])).astype(np.float32)
    want_idx, want_d2 = _py_octree_approx_nearest(xyz, res, q)
    idx, d2 = oracle.octree_approx_nearest(oracle.make_points(xyz), res, q)
    np.testing.assert_array_equal(idx, want_idx)
    np.testing.assert_array_equal(d2, want_d2)


@pytest.mark.parametrize("use_hsv", [False, True])
def test_weight_equals_numpy_composition(use_hsv):
    """The oracle's weight() as a whole against a numpy composition of its parts: transform in the
    contract's operation order, crop to the union box (inclusive, finite only), brute-force nearest neighbour with
    ties to the lower index, coherence summed in model order in fp64, raw weight = -(float)sum, then normalizeWeight."""
    from tests import util
    f32 = np.float32
    scene, model, centre = util.small_case(seed=17, n_scene=1500, n_model=80)
    scene["x"][5] = np.nan
    n = 6
    parts = util.particles_around(centre, n, seed=18)
    t = _tracker(kld=False, particle_num=n, max_particle_num=n, use_hsv=use_hsv, nn_mode=oracle.NN_EXACT_BRUTE)
    t.set_d(oracle.MAX_DIST, 0.1)
    t.set_reference(model)
    t.set_input(scene)
    t.set_particles(parts)
    t.weight(keep_nn=True)
    # numpy
    transed = []
    for p in parts:
        m = oracle.particle_to_matrix([p["x"], p["y"], p["z"], p["roll"], p["pitch"], p["yaw"]]).astype(f32)
        x, y, z = model["x"], model["y"], model["z"]
        q = np.stack([((m[r, 0] * x + m[r, 1] * y) + m[r, 2] * z) + m[r, 3] for r in range(3)], axis=1).astype(f32)
        transed.append(q)
    allq = np.concatenate(transed)
    lo, hi = allq.min(0), allq.max(0)
    np.testing.assert_array_equal(t.aabb(), np.concatenate([lo, hi]))
    sxyz = np.stack([scene["x"], scene["y"], scene["z"]], axis=1)
    keep = np.isfinite(sxyz).all(1)
    with np.errstate(invalid="ignore"):
        for d in range(3):
            keep &= (sxyz[:, d] >= lo[d]) & (sxyz[:, d] <= hi[d])
    cidx, _ = t.cropped()
    np.testing.assert_array_equal(cidx, np.nonzero(keep)[0])
    cxyz = sxyz[keep]
    crgba = scene["rgba"][keep]
    raws = []
    for i, q in enumerate(transed):
        dx = q[:, None, 0] - cxyz[None, :, 0]
        dy = q[:, None, 1] - cxyz[None, :, 1]
        dz = q[:, None, 2] - cxyz[None, :, 2]
        d2 = ((dx * dx + dy * dy) + dz * dz).astype(f32)
        nn = d2.argmin(1)                      # first minimum = lowest index
        oi, od = t.nn(i, len(model))
        np.testing.assert_array_equal(oi, nn)
        np.testing.assert_array_equal(od, d2[np.arange(len(model)), nn])
        val = 0.0
        for j in range(len(model)):
            dd = d2[j, nn[j]]
            if float(dd) < 0.1 * 0.1:
                d = float(np.sqrt(dd))
                c = 1.0 / (1.0 + d * d * 1.0)
                if use_hsv:
                    c *= oracle.hsv_coherence(int(model["rgba"][j]), int(crgba[nn[j]]), 0.1)
                val += c
        raws.append(-f32(val))
    raws = np.array(raws, dtype=f32)
    np.testing.assert_allclose(t.raw_weights(), raws, rtol=2e-7)
    w = raws.astype(np.float64)
    e = np.where(w != 0, np.exp(1.0 - 15.0 * (w - w.min()) / (w[w != 0].max() - w.min())), 0.0).astype(f32)
    np.testing.assert_allclose(t.get_particles()["weight"], (e / f32(e.astype(np.float64).sum())).astype(f32), rtol=1e-6)


def _py_approx_voxel_grid(pts, leaf):
    """Python restatement of pcl::ApproximateVoxelGrid<PointXYZRGBA>::applyFilter: a 512-entry direct-mapped
    cache keyed by the voxel, fp32 running sums of x, y, z, r, g, b in input order, a partial centroid emitted whenever
    a different voxel claims the slot, everything left flushed in slot order."""
    f32 = np.float32
    inv = f32(1.0) / f32(leaf)
    slots = [None] * 512
    out = []

    def flush(s):
        key, cnt, acc = s
        c = [a / f32(cnt) for a in acc]
        rgb = (int(c[3]) << 16) | (int(c[4]) << 8) | int(c[5])
        out.append((c[0], c[1], c[2], rgb))

    for p in pts:
        x, y, z = f32(p["x"]), f32(p["y"]), f32(p["z"])
        key = (int(np.floor(x * inv)), int(np.floor(y * inv)), int(np.floor(z * inv)))
        h = (key[0] * 7171 + key[1] * 3079 + key[2] * 4231) & 511
        s = slots[h]
        if s is not None and s[0] != key:
            flush(s)
            s = None
        if s is None:
            s = [key, 0, [f32(0)] * 6]
        rgba = int(p["rgba"])
        vals = (x, y, z, f32((rgba >> 16) & 255), f32((rgba >> 8) & 255), f32(rgba & 255))
        s[1] += 1
        s[2] = [f32(a + v) for a, v in zip(s[2], vals)]
        slots[h] = s
    for s in slots:
        if s is not None:
            flush(s)
    return out


@pytest.mark.parametrize("seed,leaf,n", [(1, 0.01, 3000), (2, 0.02, 4000), (3, 0.005, 2500)])
def test_approx_voxel_grid_pcl_equals_python_restatement(seed, leaf, n):
    rng = np.random.default_rng(seed)
    # a sensor-like stream: neighbouring points are often in the same voxel, with jumps that cause evictions
    walk = np.cumsum(rng.normal(0, leaf * 0.4, (n, 3)), axis=0) + np.array([0.2, -0.1, 1.0])
    jumps = rng.random(n) < 0.05
    walk[jumps] += rng.uniform(-0.5, 0.5, (int(jumps.sum()), 3))
    pts = oracle.make_points(walk.astype(np.float32), rng.integers(0, 1 << 24, n).astype(np.uint32))
    got = oracle.approx_voxel_grid_pcl(pts, leaf)
    want = _py_approx_voxel_grid(pts, leaf)
    assert len(got) == len(want)
    assert len(got) > len(np.unique(np.floor(walk.astype(np.float32) / np.float32(leaf)), axis=0)) * 0.9
    w = np.array([(a, b, c) for a, b, c, _ in want], dtype=np.float32)
    np.testing.assert_array_equal(np.stack([got["x"], got["y"], got["z"]], axis=1), w)
    np.testing.assert_array_equal(got["rgba"] & 0xffffff, np.array([r for _, _, _, r in want], dtype=np.uint32))


def _py_alias_table(w):
    """genAliasTable (Walker): q_i = w_i * N as a float product, a two-ended stack of the indices with
    q >= 1 (front) and q < 1 (back)."""
    n = len(w)
    q = [float(np.float32(wi) * np.float32(n)) for wi in w]
    a = list(range(n))
    hl = [0] * n
    h, l = 0, n - 1
    for i in range(n):
        if q[i] >= 1.0:
            hl[h] = i; h += 1
        else:
            hl[l] = i; l -= 1
    while h != 0 and l != n - 1:
        j, k = hl[l + 1], hl[h - 1]
        a[j] = k
        q[k] += q[j] - 1.0
        l += 1
        if q[k] < 1.0:
            hl[l] = k; l -= 1
            h -= 1
    return a, q


def test_alias_sampler_equals_python_restatement():
    n = 300
    rng = np.random.default_rng(21)
    w = (rng.random(n).astype(np.float32) ** 4)
    w = (w / w.sum()).astype(np.float32)
    a, q = _py_alias_table(w)
    s = np.zeros((n, 6), dtype=np.float32)
    s[:, 0] = np.arange(n)
    t = _tracker(kld=False, particle_num=n)
    t.set_i(oracle.SAMPLER, oracle.SAMPLER_ALIAS_PCL)
    t.set_vec6(oracle.STEP_COV, [0] * 6)
    t.set_i(oracle.QUAT_SAMPLE, 0)
    t.set_particles(oracle.make_particles(s, w))
    usel, normals, umot = synth.draws(1, n, seed=22)
    t.inject_draws(usel, normals, np.ones_like(umot))   # no motion term
    t.resample(0)
    anc = t.ancestors()
    want = []
    for u in usel[0][1:n]:   # output particle i draws with entry i of the injected arrays (slot 0 is the representative)
        ru = float(u) * n
        k = int(ru)
        ru -= k
        want.append(k if ru < q[k] else a[k])
    got = anc[anc >= 0]
    assert len(got) == len(want)
    assert (got == np.array(want)).mean() > 0.995   # (the float/double type of U*N may differ at a table boundary)
    # the table itself reproduces the weights: column i keeps min(q_i, 1), the rest goes to its alias (columns left
    # on the "heavy" stack when the light one runs out keep q >= 1 up to rounding: they always return themselves)
    assert min(q) >= 0.0
    mass = np.zeros(n)
    for i in range(n):
        mass[i] += min(q[i], 1.0)
        mass[a[i]] += max(1.0 - q[i], 0.0)
    np.testing.assert_allclose(mass / n, w, atol=2e-6)


@pytest.mark.parametrize("eps,bin_size,motion_ratio", [(0.2, 0.1, 0.25), (0.05, 0.03, 0.0), (0.1, 0.05, 1.0)])
def test_kld_resample_equals_python_control_flow(eps, bin_size, motion_ratio):
    """KLDAdaptiveParticleFilterTracker::resample restated in Python around the (separately checked)
    building blocks: alias pick, sample(), motion with probability motion_ratio_, bins by C truncation, and the stop
    rule n >= n_max || (k >= 2 && n >= calcKLBound(k))."""
    n0, n_max = 120, 900
    rng = np.random.default_rng(int(eps * 1000))
    st = np.zeros((n0, 6), dtype=np.float32)
    st[:, :3] = np.array([0.1, -0.2, 1.0]) + rng.normal(0, 0.03, (n0, 3))
    st[:, 3:] = rng.normal(0, 0.1, (n0, 3))
    w = rng.random(n0).astype(np.float32) ** 2
    w = (w / w.sum()).astype(np.float32)
    parts = oracle.make_particles(st, w)
    t = _tracker(kld=True, particle_num=n0, max_particle_num=n_max)
    t.set_i(oracle.SAMPLER, oracle.SAMPLER_ALIAS_PCL)
    t.set_d(oracle.EPSILON, eps)
    t.set_d(oracle.MOTION_RATIO, motion_ratio)
    t.set_vec6(oracle.BIN_SIZE, [bin_size] * 6)
    step = [0.015 * 0.015] * 3 + [0.015 * 0.015 * 40.0] * 3
    t.set_vec6(oracle.STEP_COV, step)
    t.set_particles(parts)
    motion = oracle.make_particles([[0.01, -0.02, 0.005, 0.02, 0.0, -0.01]])[0]
    t.set_motion(motion)
    usel, normals, umot = synth.draws(1, n_max, seed=77)
    t.inject_draws(usel, normals, umot)
    t.resample(0)
    got = t.get_particles()
    anc = t.ancestors()
    # Python
    a, q = _py_alias_table(w)
    names = ("x", "y", "z", "roll", "pitch", "yaw")
    bins, k, n = set(), 0, 0
    want_anc, want = [], []
    while True:
        ru = float(usel[0][n]) * n0
        c = int(ru)
        ru -= c
        j = c if ru < q[c] else a[c]
        x = oracle.particle_sample([float(parts[j][m]) for m in names], [0] * 6, step, normals[0][n], quat_mode=1)
        xs = np.array([x[m] for m in names], dtype=np.float32)
        if float(umot[0][n]) < motion_ratio:
            xs = (xs + np.array([motion[m] for m in names], dtype=np.float32)).astype(np.float32)
        want_anc.append(j)
        want.append(xs)
        b = tuple(int(v) for v in np.trunc(xs / np.float32(bin_size)))
        if b not in bins:
            bins.add(b)
            k += 1
        n += 1
        if not (n < n_max and (k < 2 or n < oracle.kl_bound(k, 0.99, eps))):
            break
    assert len(got) == n
    np.testing.assert_array_equal(anc[:n], np.array(want_anc))
    np.testing.assert_array_equal(np.stack([got[m] for m in names], axis=1), np.stack(want))
