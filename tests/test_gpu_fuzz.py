"""Randomised call sequences against the tracker's C ABI: reconfiguration between frames (particle counts, KLD knobs,
list modes, input clouds of changing size, resets, stage calls mixed with compute()) must never error, leak a stale CUDA
graph or break the invariants of the particle set."""
import numpy as np
import pytest

from pcl_tracking_b200 import pcl
from tests import util

pytestmark = pytest.mark.gpu


def _check(t, n_max, kld):
    p = t.getParticles()
    assert 1 <= len(p) <= n_max
    for k in ("x", "y", "z", "roll", "pitch", "yaw", "weight"):
        assert np.isfinite(p[k]).all(), k
    s = float(p["weight"].astype(np.float64).sum())
    assert abs(s - 1.0) < 1e-3, s
    r = t.getResult()
    assert all(np.isfinite(float(r[k])) for k in ("x", "y", "z", "roll", "pitch", "yaw"))


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_random_call_sequences(seed):
    rng = np.random.default_rng(seed)
    kld = bool(seed % 2)
    scenes = [util.small_case(100 + seed * 10 + k, n_scene=int(rng.integers(800, 6000)), n_model=int(rng.integers(60, 400))) for k in range(3)]
    clouds = [pcl.PointCloud(s[0]) for s in scenes]
    _, model, centre = scenes[0]
    n, n_max = int(rng.integers(20, 300)), 400
    t = (pcl.KLDAdaptiveParticleFilterOMPTracker if kld else pcl.ParticleFilterOMPTracker)(16)
    pcl.configure_like_reference(t, coherence_cls=pcl.NearestPairPointCloudCoherence, particle_num=n, max_particle_num=n_max, use_hsv=bool(rng.integers(0, 2)),
                                 iteration_num=int(rng.integers(1, 4)))
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = centre
    t.setTrans(m)
    t.seed(1000 + seed)
    t.setReferenceCloud(model)
    t.setInputCloud(clouds[0])
    t.compute()
    _check(t, n_max, kld)
    for step in range(40):
        op = int(rng.integers(0, 10))
        if op == 0:
            t.setInputCloud(clouds[int(rng.integers(0, 3))])
        elif op == 1:
            t.setCandidateLists(int(rng.integers(0, 3)))
        elif op == 2:
            t.setIterationNum(int(rng.integers(1, 5)))
        elif op == 3 and kld:
            t.setEpsilon(float(rng.choice([0.2, 0.05, 0.5])))
            t.setBinSize([float(rng.choice([0.1, 0.05, 0.2]))] * 6)
        elif op == 3 and not kld:
            t.setParticleNum(int(rng.integers(10, 400)))   # the next resample grows / shrinks the set
        elif op == 4:
            t.setReferenceCloud(scenes[int(rng.integers(0, 3))][1])
        elif op == 5:
            t.resetTracking()
        elif op == 6:
            t.weight()          # stage calls between compute() calls
            t.update()
        elif op == 7:
            t.setSampler(int(rng.integers(0, 3)))
        elif op == 8:
            t.setInputCloud(pcl.PointCloud(np.zeros(0, dtype=pcl.POINT)))   # an empty frame: compute() is a no-op
            before = t.getParticles().copy()
            t.compute()
            assert np.array_equal(t.getParticles().view(np.uint32), before.view(np.uint32))
            t.setInputCloud(clouds[int(rng.integers(0, 3))])
        t.compute()
        _check(t, n_max, kld)
    assert t.graphReplays() >= 1
