"""Per-stage parity of the particle-filter stages (K4: initParticles, resample, normalise, update)
with the CPU oracle on identical inputs and identical injected RNG draws (the reference
itself is not reproducible, so parity is per stage, never whole-trajectory).

Tolerances (north_star): ancestors / particle counts bit-exact; states 1e-4 m / 1e-4 rad (observed
~1e-6); weights 1e-5 relative."""
import numpy as np
import pytest

import oracle
from pcl_tracking_b200 import pcl, synth
from tests import util

pytestmark = pytest.mark.gpu


def _sync_oracle_from_gpu(g, o):
    o.set_particles(g.getParticles())
    o.set_result(g.getResult())
    o.set_motion(g.getMotion())


@pytest.mark.parametrize("quat", [1, 0])
def test_init_particles(quat):
    g, o = util.make_pair(kld=True, particle_num=200, max_particle_num=300, quat=quat)
    m = np.eye(4, dtype=np.float32)
    m[:3, :3] = oracle.particle_to_matrix([0, 0, 0, 0.3, -0.2, 0.5])[:, :3]
    m[:3, 3] = [0.1, 0.2, 1.3]
    g.setTrans(m); o.set_trans(m[:3])
    d = synth.draws(2, 300, seed=3)
    g.injectDraws(*d); o.inject_draws(*d)
    g.initParticles(); o.init_particles()
    util.assert_particles_close(g.getParticles(), o.get_particles(), 1e-6, 1e-6, 1e-7)
    gr, orr = g.getResult(), o.get_result()
    for k in ("x", "y", "z", "roll", "pitch", "yaw", "weight"):
        assert abs(float(gr[k]) - float(orr[k])) < 1e-6


@pytest.mark.parametrize("kld,sampler,quat", [(True, 1, 1), (True, 2, 1), (True, 1, 0), (False, 1, 1), (False, 2, 0),
                                              (True, 0, 1), (False, 0, 1)])  # sampler 0 = upstream's Walker alias table
def test_resample_matches_oracle(kld, sampler, quat):
    n, nmax = 150, 400
    g, o = util.make_pair(kld=kld, particle_num=n, max_particle_num=nmax, quat=quat, sampler=sampler, epsilon=0.1, bin_size=0.05)
    scene, model, centre = util.small_case(1, n_scene=500, n_model=50)
    rng = np.random.default_rng(5)
    parts = util.particles_around(centre, n, seed=6)
    w = rng.random(n).astype(np.float32) ** 4
    w[rng.random(n) < 0.1] = 0.0
    parts["weight"] = w / w.sum()
    cloud = pcl.PointCloud(scene)
    g.setReferenceCloud(model); g.setInputCloud(cloud)
    g.setParticles(parts); o.set_particles(parts)
    rep = parts[0].copy(); mot = parts[1].copy()
    mot["x"], mot["y"], mot["z"], mot["roll"], mot["pitch"], mot["yaw"] = 0.01, -0.02, 0.005, 0.03, -0.01, 0.02
    g.setResult(rep, mot); o.set_result(rep); o.set_motion(mot)
    d = synth.draws(2, nmax, seed=7)
    g.injectDraws(*d); o.inject_draws(*d)
    g.resample(1); o.resample(1)
    gp, op = g.getParticles(), o.get_particles()
    assert len(gp) == len(op)
    np.testing.assert_array_equal(g.ancestors(), o.ancestors())
    util.assert_particles_close(gp, op, 1e-6, 2e-6, None)
    if kld:
        assert 2 <= len(gp) <= nmax and len(gp) != n


def test_kld_particle_count_sweep():
    """C3: the KLD bound decides the particle count; epsilon / bin size sweep, cap 10 000."""
    scene, model, centre = util.small_case(2, n_scene=500, n_model=50)
    cloud = pcl.PointCloud(scene)
    counts = []
    for eps, bins in ((0.2, 0.1), (0.05, 0.05), (0.02, 0.02), (0.01, 0.005)):
        g, o = util.make_pair(kld=True, particle_num=300, max_particle_num=10000, epsilon=eps, bin_size=bins)
        parts = util.particles_around(centre, 300, seed=8)
        g.setReferenceCloud(model); g.setInputCloud(cloud)
        g.setParticles(parts); o.set_particles(parts)
        d = synth.draws(1, 10000, seed=9)
        g.injectDraws(*d); o.inject_draws(*d)
        g.resample(0); o.resample(0)
        assert len(g.getParticles()) == len(o.get_particles())
        np.testing.assert_array_equal(g.ancestors(), o.ancestors())
        counts.append(len(g.getParticles()))
    assert counts == sorted(counts) and counts[-1] == 10000 and counts[0] < 1000


def test_normalize_edge_cases():
    scene, model, centre = util.small_case(3, n_scene=300, n_model=40)
    cloud = pcl.PointCloud(scene)
    g, o = util.make_pair(kld=False, particle_num=6)
    g.setReferenceCloud(model); g.setInputCloud(cloud)
    # some particles far away: raw weight 0 stays 0, the others are min-max / exp normalised
    parts = util.particles_around(centre, 6, seed=1)
    parts["x"][4:] += 30.0
    g.setParticles(parts); o.set_reference(model); o.set_input(scene); o.set_particles(parts)
    g.weight(); o.weight()
    gw, ow = g.getParticles()["weight"], o.get_particles()["weight"]
    assert np.all(gw[4:] == 0) and np.all(ow[4:] == 0)
    np.testing.assert_allclose(gw, ow, rtol=1e-5, atol=1e-12)
    # identical particles: w_max == w_min -> uniform
    same = np.repeat(parts[:1], 6)
    g.setParticles(same); o.set_particles(same)
    g.weight(); o.weight()
    np.testing.assert_allclose(g.getParticles()["weight"], 1.0 / 6, rtol=1e-6)
    np.testing.assert_allclose(o.get_particles()["weight"], 1.0 / 6, rtol=1e-6)


def test_update_matches_oracle():
    scene, model, centre = util.small_case(4, n_scene=300, n_model=40)
    cloud = pcl.PointCloud(scene)
    for n in (1, 7, 1000, 5000):
        g, o = util.make_pair(kld=False, particle_num=n)
        g.setReferenceCloud(model); g.setInputCloud(cloud)
        rng = np.random.default_rng(n)
        parts = util.particles_around(centre, n, seed=n)
        w = rng.random(n).astype(np.float32)
        parts["weight"] = w / w.sum()
        g.setParticles(parts); o.set_particles(parts)
        rep = parts[0].copy()
        g.setResult(rep, None); o.set_result(rep)
        g.update(); o.update()
        gr, orr = g.getResult(), o.get_result()
        gm, om = g.getMotion(), o.get_motion()
        for k in ("x", "y", "z", "roll", "pitch", "yaw"):
            assert abs(float(gr[k]) - float(orr[k])) <= 1e-4, (n, k)   # oracle sums sequentially in fp32, the GPU in fp64
            assert abs(float(gm[k]) - float(om[k])) <= 1e-4, (n, k)
        assert gr["weight"] == np.float32(1.0) / np.float32(n)


def test_to_eigen_matrix():
    g, _ = util.make_pair(kld=False, particle_num=4)
    rng = np.random.default_rng(0)
    for _ in range(20):
        s = rng.uniform(-1.5, 1.5, 6).astype(np.float32)
        p = oracle.make_particles([s])[0]
        got = g.toEigenMatrix(p)
        want = oracle.particle_to_matrix(s)
        np.testing.assert_array_equal(got[:3], want)  # same double-evaluated sin/cos, same fp32 products
        np.testing.assert_array_equal(got[3], [0, 0, 0, 1])


def test_full_stage_sequence_with_resync():
    """One compute() worth of stages, re-synchronising the oracle's inputs from the GPU before each stage:
    every stage sees bit-identical inputs, so the per-stage tolerances apply end to end."""
    scene, model, centre = util.small_case(6, n_scene=5000, n_model=350)
    g, o = util.make_pair(kld=True, particle_num=120, max_particle_num=250, use_hsv=True, oracle_nn=oracle.NN_EXACT_GRID)
    util.set_trans_both(g, o, centre)
    d = synth.draws(2, 250, seed=11)
    g.injectDraws(*d); o.inject_draws(*d)
    cloud = pcl.PointCloud(scene)
    g.setReferenceCloud(model); g.setInputCloud(cloud)
    o.set_reference(model); o.set_input(scene)
    g.initParticles(); o.init_particles()
    util.assert_particles_close(g.getParticles(), o.get_particles(), 1e-6, 1e-6, 1e-7)
    for it in range(2):
        if it > 0:
            _sync_oracle_from_gpu(g, o)
            g.resample(it); o.resample(it)
            np.testing.assert_array_equal(g.ancestors(), o.ancestors())
            util.assert_particles_close(g.getParticles(), o.get_particles(), 1e-6, 2e-6, None)
        _sync_oracle_from_gpu(g, o)
        g.weight(); o.weight()
        np.testing.assert_array_equal(g.aabb(), o.aabb())
        np.testing.assert_allclose(g.rawWeights(), o.raw_weights(), rtol=1e-5)
        np.testing.assert_allclose(g.getParticles()["weight"], o.get_particles()["weight"], rtol=1e-5, atol=1e-12)
        _sync_oracle_from_gpu(g, o)
        g.update(); o.update()
        gr, orr = g.getResult(), o.get_result()
        for k in ("x", "y", "z", "roll", "pitch", "yaw"):
            assert abs(float(gr[k]) - float(orr[k])) <= 1e-4


@pytest.mark.parametrize("n_new", [230, 40])
def test_fixed_tracker_follows_set_particle_num(n_new):
    """ParticleFilterTracker::resample produces particle_num_ particles: setParticleNum between frames grows or shrinks
    the set at the next resample, drawing from the whole old set (same ancestors as the oracle)."""
    n = 120
    g, o = util.make_pair(kld=False, particle_num=n, use_hsv=False)
    scene, model, centre = util.small_case(4, n_scene=600, n_model=60)
    rng = np.random.default_rng(11)
    parts = util.particles_around(centre, n, seed=12)
    w = rng.random(n).astype(np.float32)
    parts["weight"] = w / w.sum()
    cloud = pcl.PointCloud(scene)
    g.setReferenceCloud(model); g.setInputCloud(cloud)
    g.setParticles(parts); o.set_particles(parts)
    g.setResult(parts[0].copy(), parts[1].copy()); o.set_result(parts[0].copy())
    g.setParticleNum(n_new); o.set_i(oracle.PARTICLE_NUM, n_new)
    d = synth.draws(1, max(n, n_new), seed=13)
    g.injectDraws(*d); o.inject_draws(*d)
    g.resample(0); o.resample(0)
    gp, op = g.getParticles(), o.get_particles()
    assert len(gp) == len(op) == n_new
    np.testing.assert_array_equal(g.ancestors(), o.ancestors())
    assert g.ancestors()[1:].max() < n and (n_new < n or g.ancestors()[1:].max() > 0)
    util.assert_particles_close(gp, op, 1e-6, 2e-6, None)
    # and a whole compute() keeps the new count
    g.injectDraws(None, None, None)
    g.compute()
    assert len(g.getParticles()) == n_new
    assert abs(float(g.getParticles()["weight"].astype(np.float64).sum()) - 1.0) < 1e-4
