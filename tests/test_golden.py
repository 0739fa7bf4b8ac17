"""Committed fixture (tests/golden/oracle_small_case.npz, made by tests/golden/make_golden.py from the CPU oracle; the
reference itself ships no golden vectors for this path): the oracle must keep reproducing it bit for bit, and the CUDA
path must match it to the north_star tolerances."""
import os

import numpy as np
import pytest

from tests.golden import make_golden

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_small_case.npz")


def test_oracle_reproduces_committed_fixture():
    want = np.load(GOLDEN)
    got = make_golden.compute()
    for k in want.files:
        a, b = want[k], np.asarray(got[k])
        assert a.shape == b.shape, k
        assert a.tobytes() == b.tobytes(), k


@pytest.mark.gpu
def test_cuda_path_matches_committed_fixture():
    from pcl_tracking_b200 import pcl, synth
    from tests import util
    want = np.load(GOLDEN)
    scene, model, parts = want["scene"], want["model"], want["particles"]
    g, _ = util.make_pair(kld=True, particle_num=len(parts), max_particle_num=96, use_hsv=True)
    cloud = pcl.PointCloud(scene)
    g.setReferenceCloud(model); g.setInputCloud(cloud); g.setParticles(parts); g.setDebugNN(4)
    g.weight()
    np.testing.assert_array_equal(g.aabb(), want["aabb"])
    for p in range(4):
        gi, gd = g.nn(p, len(model))
        m = want["nn_d2"][p].astype(np.float64) < 0.1 * 0.1
        np.testing.assert_array_equal(gi[m], want["nn_idx"][p][m])
        np.testing.assert_array_equal(gd[m], want["nn_d2"][p][m])
    np.testing.assert_allclose(g.rawWeights(), want["raw"], rtol=1e-5)
    np.testing.assert_allclose(g.getParticles()["weight"], want["weights"], rtol=1e-5, atol=1e-12)
    g.injectDraws(*synth.draws(1, 96, seed=5))
    g.update()
    r = g.getResult()
    for k in ("x", "y", "z", "roll", "pitch", "yaw"):
        assert abs(float(r[k]) - float(want["result"][k])) <= 1e-4
    # resample from the fixture's weights (so that both sides select from identical tables)
    p2 = parts.copy()
    p2["weight"] = want["weights"]
    g.setParticles(p2)
    g.resample(0)
    np.testing.assert_array_equal(g.ancestors(), want["ancestors"])
    util.assert_particles_close(g.getParticles(), want["resampled"], 1e-4, 1e-4, None)
    vg = pcl.ApproximateVoxelGrid()
    vg.setLeafSize(0.02); vg.setPassThrough("z", 0.0, 10.0); vg.setInputCloud(cloud)
    ds = vg.filter().to_numpy()
    assert ds.tobytes() == want["downsampled"].tobytes()


AUX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_aux_case.npz")


def test_oracle_reproduces_committed_aux_fixture():
    want = np.load(AUX)
    got = make_golden.compute_aux()
    for k in want.files:
        a, b = want[k], np.asarray(got[k])
        assert a.shape == b.shape, k
        if k.startswith("box_") and k not in ("box_model", "box_state"):
            np.testing.assert_allclose(b, a, rtol=1e-6, atol=1e-7, err_msg=k)  # (numpy's eigh may differ in the last bit between builds)
        else:
            assert a.tobytes() == b.tobytes(), k


@pytest.mark.gpu
def test_cuda_path_matches_committed_aux_fixture():
    from pcl_tracking_b200 import pcl
    want = np.load(AUX)
    raw = want["pc2_raw"]
    got = pcl.PointCloud().fromPointCloud2(raw, 24, 5, 24, raw.shape[1], 8, 12, 16, 0).to_numpy()
    assert got.tobytes() == want["pc2_points"].tobytes()
    ec = pcl.EuclideanClusterExtraction()
    ec.setClusterTolerance(0.02); ec.setMinClusterSize(50); ec.setMaxClusterSize(25000)
    ec.setInputCloud(pcl.PointCloud(want["cluster_points"]))
    ec.extract()
    np.testing.assert_array_equal(ec.labels, want["cluster_labels"])
    np.testing.assert_array_equal(ec.sizes, want["cluster_sizes"])
    t = pcl.ParticleFilterOMPTracker(16)
    pcl.configure_like_reference(t, particle_num=8)
    t.setReferenceCloud(want["box_model"])
    st = want["box_state"]
    rep = np.zeros(1, dtype=pcl.PARTICLE)
    for k, v in zip(("x", "y", "z", "roll", "pitch", "yaw"), st):
        rep[k] = v
    rep["one"] = 1.0
    t.setResult(rep[0])
    box = t.getResultBox(-0.005)
    np.testing.assert_allclose(box["centroid"], want["box_centroid"], atol=1e-5)
    np.testing.assert_allclose(box["extent"], want["box_extent"], atol=1e-4)
    np.testing.assert_allclose(box["center"], want["box_center"], atol=1e-4)
    np.testing.assert_allclose(box["eigenvalues"], want["box_eigenvalues"], rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(box["axes"], want["box_axes"], atol=2e-3)


MODES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_parity_modes_case.npz")


def test_oracle_reproduces_committed_parity_modes_fixture():
    want = np.load(MODES)
    got = make_golden.compute_parity_modes()
    for k in want.files:
        a, b = want[k], np.asarray(got[k])
        assert a.shape == b.shape, k
        assert a.tobytes() == b.tobytes(), k


@pytest.mark.gpu
def test_cuda_parity_modes_match_committed_fixture():
    """PFT_NN_PCL_APPROX, pft_approx_voxel_grid_pcl and the change detector against the committed fixture."""
    from pcl_tracking_b200 import pcl
    from tests import util
    want = np.load(MODES)
    scene, model, parts = want["scene"], want["model"], want["particles"]
    n = len(parts)
    g, _ = util.make_pair(kld=False, particle_num=n, use_hsv=True)
    g._si(pcl.capi.NN_MODE, pcl.capi.NN_PCL_APPROX)
    cloud = pcl.PointCloud(scene)
    g.setReferenceCloud(model); g.setInputCloud(cloud); g.setParticles(parts); g.setDebugNN(4)
    g.weight()
    for p in range(4):
        gi, gd = g.nn(p, len(model))
        np.testing.assert_array_equal(gi, want["approx_nn_idx"][p])
        np.testing.assert_array_equal(gd, want["approx_nn_d2"][p])
    np.testing.assert_allclose(g.rawWeights(), want["approx_raw"], rtol=1e-5)
    np.testing.assert_allclose(g.getParticles()["weight"], want["approx_weights"], rtol=1e-5, atol=1e-12)
    vg = pcl.ApproximateVoxelGrid()
    vg.setLeafSize(0.02); vg.setPassThrough("z", 0.0, 10.0); vg.setPclApproximateMode(True); vg.setInputCloud(cloud)
    assert vg.filter().to_numpy().tobytes() == want["approx_grid"].tobytes()
    c, _ = util.make_pair(kld=False, particle_num=n, use_hsv=True)
    c.setUseChangeDetector(True); c.setIntervalOfChangeDetection(0); c.setMinPointsOfChangeDetection(2); c.setResolutionOfChangeDetection(0.03)
    c.setReferenceCloud(model); c.setParticles(parts)
    for k, dx in enumerate((0.0, 0.0, 0.05, 0.05)):
        sc = scene.copy()
        sc["x"] += np.float32(dx)
        c.setInputCloud(pcl.PointCloud(sc))
        c.weight()
        i = c.changeDetectorInfo()
        assert [i["counter"], i["tests"], i["last_found"], int(i["changed"])] == want["cd_info"][k].tolist()
        np.testing.assert_allclose(c.getParticles()["weight"], want["cd_weights"][k], rtol=1e-5, atol=1e-12)


PCL_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pcl_small_case.npz")


@pytest.mark.skipif(not os.path.exists(PCL_GOLDEN), reason="tests/golden/pcl_small_case.npz is produced by tests/golden/make_pcl_golden.cpp on a "
                                                           "machine with PCL 1.8.0 (not installable in this image): until it is committed the oracle is "
                                                           "unpinned against PCL binaries")
def test_oracle_matches_pcl_golden_when_present():
    """Pins the oracle to REAL PCL 1.8.0: what make_pcl_golden.cpp computed with pcl::tracking / pcl::filters / pcl::search on the
    inputs of oracle_small_case.npz against what the restatement computes on the same inputs."""
    import oracle
    pclg = np.load(PCL_GOLDEN)
    inp = np.load(GOLDEN)
    scene, model, parts = inp["scene"], inp["model"], inp["particles"]

    def tracker(nn_mode):
        t = oracle.Tracker(kld=True)
        oracle.configure_like_reference(t, particle_num=len(parts), max_particle_num=96, use_hsv=True, nn_mode=nn_mode)
        t.set_reference(model); t.set_input(scene); t.set_particles(parts)
        return t

    # exact coherence: crop box and count bit-exact, raw / normalised weights 1e-5, update() 1e-5
    t = tracker(oracle.NN_EXACT_BRUTE)
    t.weight(keep_nn=True)
    np.testing.assert_array_equal(t.aabb(), pclg["aabb"])
    cidx, _ = t.cropped()
    assert len(cidx) == int(pclg["cropped_count"][0])
    np.testing.assert_allclose(t.raw_weights(), pclg["raw"], rtol=1e-5)
    np.testing.assert_allclose(t.get_particles()["weight"], pclg["weights"], rtol=1e-5, atol=1e-12)
    t.update()
    r = t.get_result()
    for i, k in enumerate(("x", "y", "z", "one", "roll", "pitch", "yaw")):
        if k != "one":
            assert abs(float(r[k]) - float(pclg["result"][i])) < 1e-5, k
    # the reference's own (approximate) coherence: the greedy octree search, pair by pair, then the weights
    a = tracker(oracle.NN_PCL_APPROX)
    a.weight(keep_nn=True)
    for p in range(pclg["approx_nn_idx"].shape[0]):
        oi, od = a.nn(p, len(model))
        np.testing.assert_array_equal(oi, pclg["approx_nn_idx"][p])
        np.testing.assert_array_equal(od, pclg["approx_nn_d2"][p])
    np.testing.assert_allclose(a.raw_weights(), pclg["approx_raw"], rtol=1e-5)
    # filters: PassThrough order-preserving and bit-exact; ApproximateVoxelGrid bit-exact incl. duplicates and flush order;
    # VoxelGrid as a set of centroids (std::sort is unstable: the fp32 sums of a voxel may differ in the last bits)
    passed = oracle.passthrough(scene, 2, 0.0, 10.0)
    assert passed.tobytes() == pclg["passthrough"].tobytes()
    assert oracle.approx_voxel_grid_pcl(passed, 0.02).tobytes() == pclg["approx_voxel_grid"].tobytes()
    vg, want = oracle.voxel_grid_pcl(passed, 0.02), pclg["voxel_grid"]
    assert len(vg) == len(want)
    for k in ("x", "y", "z"):
        np.testing.assert_allclose(vg[k], want[k], atol=1e-6)
    # scalars: calcKLBound, Walker's alias table of the normalised weights
    np.testing.assert_allclose([oracle.kl_bound(k, 0.99, 0.2) for k in range(2, 201)], pclg["kl_bound"], rtol=1e-12)
    w = parts.copy()
    w["weight"] = pclg["weights"]
    t.set_particles(w)
    aa, qq = t.alias_table()
    np.testing.assert_array_equal(aa, pclg["alias_a"])
    np.testing.assert_allclose(qq, pclg["alias_q"], rtol=1e-12)


def test_pcl_golden_writer_emits_loadable_npy(tmp_path):
    """make_pcl_golden.cpp cannot be built here (no PCL), but its .npy writer can be checked: the same function compiled on
    its own writes files numpy loads with the declared dtype and shape."""
    import re
    import subprocess
    src = open(os.path.join(os.path.dirname(PCL_GOLDEN), "make_pcl_golden.cpp")).read()
    body = re.search(r"(// minimal \.npy \(v1\.0\) writer.*?\n}\n)", src, re.S).group(1)
    prog = tmp_path / "w.cpp"
    prog.write_text("#include <fstream>\n#include <string>\n#include <vector>\n" + body +
                    'int main(int, char** v) { float a[6] = {0, 1, 2, 3, 4, 5}; int b[3] = {7, 8, 9};\n'
                    ' write_npy(std::string(v[1]) + "/a.npy", "<f4", {2, 3}, a, 24); write_npy(std::string(v[1]) + "/b.npy", "<i4", {3}, b, 12); return 0; }\n')
    exe = tmp_path / "w"
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-o", str(exe), str(prog)])
    subprocess.check_call([str(exe), str(tmp_path)])
    a, b = np.load(tmp_path / "a.npy"), np.load(tmp_path / "b.npy")
    assert a.dtype == np.float32 and a.shape == (2, 3) and a.tolist() == [[0, 1, 2], [3, 4, 5]]
    assert b.dtype == np.int32 and b.tolist() == [7, 8, 9]
