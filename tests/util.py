"""Shared helpers of the parity tests: matched construction of the CUDA tracker (through the C ABI)
and of the CPU oracle, on the same seeded inputs and the same injected RNG draws."""
import numpy as np

import oracle
from pcl_tracking_b200 import pcl, synth


def small_case(seed=0, n_scene=4000, n_model=300, box=(0.30, 0.24, 0.20), spread=0.6, colour_noise=25):
    """A box-surface model (centred) and a scene = noisy box surface + clutter, object at `centre`."""
    rng = np.random.default_rng(seed)
    size = np.asarray(box)

    def box_surface(m):
        areas = np.array([size[1] * size[2], size[0] * size[2], size[0] * size[1]])
        faces = rng.choice(3, size=m, p=areas / areas.sum())
        q = rng.uniform(-0.5, 0.5, size=(m, 3)) * size
        sign = np.where(rng.random(m) < 0.5, -0.5, 0.5)
        q[np.arange(m), faces] = sign * size[faces]
        return q

    centre = np.array([0.11, -0.07, 1.23])
    model = np.zeros(n_model, dtype=pcl.POINT)
    mq = box_surface(n_model)
    model["x"], model["y"], model["z"] = mq[:, 0], mq[:, 1], mq[:, 2]
    mc = np.clip(np.array([200, 60, 50]) + rng.integers(-colour_noise, colour_noise + 1, (n_model, 3)), 0, 255).astype(np.uint32)
    model["rgba"] = (255 << 24) | (mc[:, 0] << 16) | (mc[:, 1] << 8) | mc[:, 2]
    n_obj = n_scene // 2
    obj = box_surface(n_obj) + centre + rng.normal(0, 0.002, (n_obj, 3))
    clutter = centre + rng.uniform(-spread, spread, (n_scene - n_obj, 3))
    xyz = np.concatenate([obj, clutter])
    scene = np.zeros(n_scene, dtype=pcl.POINT)
    scene["x"], scene["y"], scene["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    sc = rng.integers(0, 256, (n_scene, 3)).astype(np.uint32)
    sc[:n_obj] = np.clip(np.array([200, 60, 50]) + rng.integers(-colour_noise, colour_noise + 1, (n_obj, 3)), 0, 255)
    scene["rgba"] = (255 << 24) | (sc[:, 0] << 16) | (sc[:, 1] << 8) | sc[:, 2]
    scene = scene[rng.permutation(n_scene)]
    return scene, model, centre.astype(np.float32)


def particles_around(centre, n, seed=1, sigma_t=0.015, sigma_r=0.08):
    rng = np.random.default_rng(seed)
    s = np.zeros((n, 6), dtype=np.float32)
    s[:, :3] = centre + rng.normal(0, sigma_t, (n, 3))
    s[:, 3:] = rng.normal(0, sigma_r, (n, 3))
    return oracle.make_particles(s, np.full(n, 1.0 / n, dtype=np.float32))


STEP_COV = [0.015 * 0.015] * 3 + [0.015 * 0.015 * 40.0] * 3


def make_pair(kld=True, particle_num=64, max_particle_num=128, use_hsv=True, iteration_num=2, oracle_nn=oracle.NN_EXACT_BRUTE,
              epsilon=0.2, bin_size=0.1, max_dist=0.1, quat=1, sampler=1, search_res=0.01):
    """(gpu_tracker, oracle_tracker) configured identically with the reference's knob values."""
    g = pcl.KLDAdaptiveParticleFilterOMPTracker(16) if kld else pcl.ParticleFilterOMPTracker(16)
    pcl.configure_like_reference(g, coherence_cls=pcl.NearestPairPointCloudCoherence, particle_num=particle_num,
                                 max_particle_num=max_particle_num, use_hsv=use_hsv, iteration_num=iteration_num)
    if kld:
        g.setEpsilon(epsilon)
        g.setBinSize([bin_size] * 6)
    g._sd(pcl.capi.MAX_DIST, max_dist)
    g._sd(pcl.capi.SEARCH_RESOLUTION, search_res)
    g.setQuaternionSampling(quat)
    g.setSampler(sampler)
    o = oracle.Tracker(kld=kld)
    oracle.configure_like_reference(o, particle_num=particle_num, max_particle_num=max_particle_num, use_hsv=use_hsv, nn_mode=oracle_nn,
                                    iteration_num=iteration_num)
    o.set_d(oracle.EPSILON, epsilon)
    o.set_vec6(oracle.BIN_SIZE, [bin_size] * 6)
    o.set_d(oracle.MAX_DIST, max_dist)
    o.set_d(oracle.GRID_CELL, search_res)
    o.set_i(oracle.QUAT_SAMPLE, quat)
    o.set_i(oracle.SAMPLER, sampler)
    return g, o


def set_trans_both(g, o, centre):
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = centre
    g.setTrans(m)
    o.set_trans(m[:3])


def assert_particles_close(a, b, tol_pos=1e-4, tol_ang=1e-4, tol_w=1e-5):
    assert len(a) == len(b), (len(a), len(b))
    for k in ("x", "y", "z"):
        np.testing.assert_allclose(a[k], b[k], rtol=0, atol=tol_pos, err_msg=k)
    for k in ("roll", "pitch", "yaw"):
        np.testing.assert_allclose(a[k], b[k], rtol=0, atol=tol_ang, err_msg=k)
    if tol_w is not None:
        np.testing.assert_allclose(a["weight"], b["weight"], rtol=tol_w, atol=1e-12, err_msg="weight")
