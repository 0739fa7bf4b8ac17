"""Generates tests/golden/oracle_small_case.npz from the CPU ORACLE (oracle/pft_oracle.cpp).

The reference ships no golden vectors for the tracking path and PCL 1.8.0 cannot be run here (DESIGN.md section 2:
"parity unpinned"), so this fixture does NOT pin the oracle to the reference; it freezes the oracle's own answers on a
small seeded case so that (a) a later change of the oracle that alters results is caught on CPU and (b) the CUDA path
is compared against a committed artefact as well as against the live oracle.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from pcl_tracking_b200 import synth  # noqa: E402
from tests import util  # noqa: E402


def compute():
    scene, model, centre = util.small_case(seed=1234, n_scene=3000, n_model=200)
    n = 40
    parts = util.particles_around(centre, n, seed=77)
    t = oracle.Tracker(kld=True)
    oracle.configure_like_reference(t, particle_num=n, max_particle_num=96, use_hsv=True, nn_mode=oracle.NN_EXACT_BRUTE)
    t.set_reference(model)
    t.set_input(scene)
    t.set_particles(parts)
    t.weight(keep_nn=True)
    cidx, _ = t.cropped()
    nn_idx, nn_d2 = [], []
    for p in range(4):
        i, d = t.nn(p, len(model))
        nn_idx.append(np.where(i >= 0, cidx[np.maximum(i, 0)], -1))
        nn_d2.append(d)
    out = dict(scene=scene, model=model, particles=parts, aabb=t.aabb(), raw=t.raw_weights(), weights=t.get_particles()["weight"],
               nn_idx=np.stack(nn_idx), nn_d2=np.stack(nn_d2))
    # one KLD resample with fixed draws from the weighted set, then update
    usel, normals, umot = synth.draws(1, 96, seed=5)
    t.inject_draws(usel, normals, umot)
    t.update()
    out["result"] = np.array(t.get_result())
    t.resample(0)
    out["ancestors"] = t.ancestors()
    out["resampled"] = t.get_particles()
    out["downsampled"] = oracle.voxel_grid_exact(scene, 0.02, 2, 0.0, 10.0)
    return out


def compute_aux():
    """Second fixture (oracle_aux_case.npz): the steps either side of the path -- frame ingest, model acquisition,
    result post-processing."""
    rng = np.random.default_rng(4321)
    # a 24 x 5 PointCloud2 message with 24-byte records (rgb first, xyz at 8/12/16) and 8 bytes of row padding
    w, h, step, pad = 24, 5, 24, 8
    pts = oracle.make_points(rng.uniform(-1, 3, (w * h, 3)).astype(np.float32), rng.integers(0, 1 << 32, w * h, dtype=np.uint64).astype(np.uint32))
    pts["x"][7] = np.nan
    rec = rng.integers(0, 256, (w * h, step), dtype=np.uint8)
    for name, off in (("rgba", 0), ("x", 8), ("y", 12), ("z", 16)):
        rec[:, off:off + 4] = np.ascontiguousarray(pts[name]).view(np.uint8).reshape(-1, 4)
    raw = rng.integers(0, 256, (h, w * step + pad), dtype=np.uint8)
    raw[:, :w * step] = rec.reshape(h, w * step)
    out = dict(pc2_raw=raw, pc2_points=oracle.from_pointcloud2(raw.tobytes(), w, h, step, w * step + pad, 8, 12, 16, 0))
    # clustering: three blobs + clutter
    blobs = [np.array(c) + rng.uniform(-0.5, 0.5, (m, 3)) * np.array(sz) for c, m, sz in
             (((0.0, 0.0, 1.0), 260, (0.10, 0.06, 0.05)), ((0.4, 0.1, 1.1), 150, (0.05, 0.05, 0.08)), ((-0.3, 0.3, 0.9), 60, (0.04, 0.04, 0.04)))]
    xyz = np.concatenate(blobs + [rng.uniform(-1, 1, (80, 3)) + np.array([0, 0, 1.0])]).astype(np.float32)
    xyz = xyz[rng.permutation(len(xyz))]
    cl_pts = oracle.make_points(xyz, rng.integers(0, 1 << 32, len(xyz), dtype=np.uint64).astype(np.uint32))
    labels, sizes = oracle.euclidean_clusters(cl_pts, 0.02, 50, 25000)
    out.update(cluster_points=cl_pts, cluster_labels=labels, cluster_sizes=sizes)
    # result box of a model at a pose
    _, model, _ = util.small_case(seed=99, n_scene=10, n_model=300)
    state = np.array([0.21, -0.13, 1.07, 0.3, -0.2, 0.7], dtype=np.float32)
    box = oracle.result_box(model, state, -0.005)
    out.update(box_model=model, box_state=state, box_centroid=box["centroid"], box_axes=box["axes"], box_extent=box["extent"],
               box_center=box["center"], box_eigenvalues=box["eigenvalues"])
    return out


def compute_parity_modes():
    """Third fixture (oracle_parity_modes_case.npz): the modes that reproduce upstream behaviour the product path
    improves on -- the greedy octree approxNearestSearch of the Approx coherence, the 512-slot ApproximateVoxelGrid
    cache, and the change detector."""
    scene, model, centre = util.small_case(seed=4242, n_scene=3500, n_model=180)
    n = 24
    parts = util.particles_around(centre, n, seed=31)
    t = oracle.Tracker(kld=False)
    oracle.configure_like_reference(t, particle_num=n, max_particle_num=n, use_hsv=True, nn_mode=oracle.NN_PCL_APPROX)
    t.set_d(oracle.MAX_DIST, 0.1)
    t.set_d(oracle.OCTREE_RES, 0.01)
    t.set_reference(model)
    t.set_input(scene)
    t.set_particles(parts)
    t.weight(keep_nn=True)
    cidx, _ = t.cropped()
    nn_idx, nn_d2 = [], []
    for p in range(4):
        i, d = t.nn(p, len(model))
        nn_idx.append(np.where(i >= 0, cidx[np.maximum(i, 0)], -1))
        nn_d2.append(d)
    out = dict(scene=scene, model=model, particles=parts, approx_nn_idx=np.stack(nn_idx), approx_nn_d2=np.stack(nn_d2),
               approx_raw=t.raw_weights(), approx_weights=t.get_particles()["weight"])
    out["approx_grid"] = oracle.approx_voxel_grid_pcl(oracle.passthrough(scene, 2, 0.0, 10.0), 0.02)
    # change detector: the tracker's own sequence of tests on a scene that stands still, jumps and stands still again
    c = oracle.Tracker(kld=False)
    oracle.configure_like_reference(c, particle_num=n, max_particle_num=n, use_hsv=True, nn_mode=oracle.NN_EXACT_GRID)
    c.set_d(oracle.MAX_DIST, 0.1)
    c.set_i(oracle.SAMPLER, oracle.SAMPLER_CDF)
    c.set_change_detector(True, interval=0, min_points=2, resolution=0.03)
    c.set_reference(model)
    c.set_particles(parts)
    infos, weights = [], []
    for dx in (0.0, 0.0, 0.05, 0.05):
        sc = scene.copy()
        sc["x"] += np.float32(dx)
        c.set_input(sc)
        c.weight()
        i = c.change_detector_info()
        infos.append([i["counter"], i["tests"], i["last_found"], int(i["changed"])])
        weights.append(c.get_particles()["weight"].copy())
    out["cd_info"] = np.array(infos, dtype=np.int32)
    out["cd_weights"] = np.stack(weights)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_parity_modes_case.npz"), **compute_parity_modes())
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_aux_case.npz"), **compute_aux())
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_small_case.npz"), **compute())
    print("written")
