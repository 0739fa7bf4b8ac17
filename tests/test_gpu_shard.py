"""Particle sharding (multi-GPU row of SURVEY 8e) emulated on ONE GPU: R trackers with set_shard(R, r)
play the ranks; the two exchange points of weight() (crop-box all-reduce, raw-weight all-gather) are
done by hand through the phase API.  The sharded result must equal the unsharded one bit for bit."""
import numpy as np
import pytest

from pcl_tracking_b200 import pcl, synth
from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("R,kld", [(2, False), (3, True), (8, True)])
def test_sharded_weight_equals_unsharded(R, kld):
    scene, model, centre = util.small_case(20 + R, n_scene=4000, n_model=260)
    cloud = pcl.PointCloud(scene)
    n = 101
    parts = util.particles_around(centre, n, seed=4)

    def mk(shard=None):
        g, _ = util.make_pair(kld=kld, particle_num=n, max_particle_num=160, use_hsv=True)
        if shard:
            g.setShard(*shard)
        g.setReferenceCloud(model); g.setInputCloud(cloud); g.setParticles(parts)
        return g

    ref = mk()
    ref.weight()
    ranks = [mk((R, r)) for r in range(R)]
    for g in ranks:
        g.weightPhase(0)
    boxes = np.stack([g.cropBox() for g in ranks])
    box = np.concatenate([boxes[:, :3].min(0), boxes[:, 3:].max(0)])   # all-reduce(min), all-reduce(max)
    for g in ranks:
        g.setCropBox(box)
        g.weightPhase(1)
    slices = [g.rawSlice(r) for r, g in enumerate(ranks)]              # all-gather
    for g in ranks:
        for r in range(R):
            g.setRawSlice(r, slices[r])
        g.weightPhase(2)
    want_raw, want = ref.rawWeights(), ref.getParticles()
    np.testing.assert_array_equal(ref.aabb(), ranks[0].aabb())
    for g in ranks:
        assert np.array_equal(g.rawWeights().view(np.uint32), want_raw.view(np.uint32))
        assert np.array_equal(g.getParticles().view(np.uint32), want.view(np.uint32))
