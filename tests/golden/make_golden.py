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


if __name__ == "__main__":
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_small_case.npz"), **compute())
    print("written")
