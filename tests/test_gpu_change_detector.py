"""the change-detector path of pcl::tracking::ParticleFilterTracker (setUseChangeDetector; off in the
reference) -- GPU against the oracle's restatement of OctreePointCloudChangeDetector, through the C ABI."""
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


def _pair(interval, min_points, resolution, n=80):
    g, o = util.make_pair(kld=False, particle_num=n, use_hsv=True, oracle_nn=oracle.NN_EXACT_GRID)
    g.setUseChangeDetector(True)
    g.setIntervalOfChangeDetection(interval)
    g.setMinPointsOfChangeDetection(min_points)
    g.setResolutionOfChangeDetection(resolution)
    o.set_change_detector(True, interval=interval, min_points=min_points, resolution=resolution)
    return g, o


@pytest.mark.parametrize("interval,min_points,resolution,expect_skips", [
    (1, 1, 0.02, False),    # every test sees a few new voxels at the rim of the (moving) crop box: never frozen
    (1, 10, 0.04, True),    # sparse rim voxels are ignored: frozen on the static stretches, released by the jumps
    (0, 2, 0.03, True),     # a test at every weight(): frozen from the second frame on
    (3, 1, 0.01, False),
])
def test_change_detector_sequence_matches_oracle(interval, min_points, resolution, expect_skips):
    """A frame sequence with static stretches and jumps: the same tests run, report the same number of points in new
    voxels and take the same decisions as the oracle; skipped weight() calls leave the particle states alone and
    re-normalise the weights exactly as upstream does."""
    scene, model, centre = util.small_case(7, n_scene=5000, n_model=250)
    n = 80
    g, o = _pair(interval, min_points, resolution, n)
    util.set_trans_both(g, o, centre)
    g.setReferenceCloud(model); o.set_reference(model)
    shifts = [0.0, 0.0, 0.0, 0.0, 0.03, 0.03, 0.03, 0.0, 0.0, 0.06]
    skipped = computed = 0
    for f, dx in enumerate(shifts):
        sc = scene.copy()
        sc["x"] += np.float32(dx)
        d = synth.draws(2, n + 1, seed=100 + f)
        g.injectDraws(*d); o.inject_draws(*d)
        if f:
            _sync_oracle_from_gpu(g, o)
        g.setInputCloud(pcl.PointCloud(sc)); g.compute()
        o.set_input(sc); o.compute()
        gi, oi = g.changeDetectorInfo(), o.change_detector_info()
        assert gi == oi, (f, gi, oi)
        gp, op = g.getParticles(), o.get_particles()
        util.assert_particles_close(gp, op, 1e-4, 1e-4, 1e-5)
        gr, orr = g.getResult(), o.get_result()
        for k in ("x", "y", "z", "roll", "pitch", "yaw"):
            assert abs(float(gr[k]) - float(orr[k])) <= 1e-4
        if not gi["changed"]:
            skipped += 1
        else:
            computed += 1
    assert computed > 0
    assert (skipped > 0) == expect_skips


def test_change_detector_skips_coherence_and_keeps_states():
    """weight() right after a test that found nothing: the raw weights are the old weights, states untouched."""
    scene, model, centre = util.small_case(8, n_scene=4000, n_model=200)
    g, o = _pair(0, 1, 0.02, 60)
    parts = util.particles_around(centre, 60, seed=3)
    cloud = pcl.PointCloud(scene)
    g.setReferenceCloud(model); g.setInputCloud(cloud); g.setParticles(parts)
    o.set_reference(model); o.set_input(scene); o.set_particles(parts)
    g.weight(); o.weight()                      # first test: everything is new
    assert g.changeDetectorInfo() == o.change_detector_info()
    assert g.changeDetectorInfo()["changed"] and g.changeDetectorInfo()["last_found"] > 0
    np.testing.assert_allclose(g.getParticles()["weight"], o.get_particles()["weight"], rtol=1e-5, atol=1e-12)
    w1 = g.getParticles()["weight"].copy()
    o.set_particles(g.getParticles())
    g.weight(); o.weight()                      # same crop: nothing new
    gi = g.changeDetectorInfo()
    assert gi == o.change_detector_info()
    assert not gi["changed"] and gi["last_found"] == 0 and gi["tests"] == 2
    np.testing.assert_array_equal(g.rawWeights(), w1)
    gp, op = g.getParticles(), o.get_particles()
    for k in ("x", "y", "z", "roll", "pitch", "yaw"):
        np.testing.assert_array_equal(gp[k], parts[k])
    np.testing.assert_allclose(gp["weight"], op["weight"], rtol=1e-6, atol=1e-12)
    # off again: plain weight()
    g.setUseChangeDetector(False); o.set_change_detector(False)
    g.weight(); o.weight()
    assert g.changeDetectorInfo()["changed"]
    np.testing.assert_allclose(g.rawWeights(), o.raw_weights(), rtol=1e-5)


def test_change_detector_rejected_on_a_sharded_tracker():
    scene, model, centre = util.small_case(9, n_scene=1000, n_model=60)
    g, _ = _pair(1, 1, 0.02, 16)
    g.setShard(2, 0)
    g.setReferenceCloud(model); g.setInputCloud(pcl.PointCloud(scene)); g.setParticles(util.particles_around(centre, 16, seed=1))
    with pytest.raises(pcl.PftError):
        g.weight()
