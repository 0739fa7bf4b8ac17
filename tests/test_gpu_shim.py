"""The C++ shim drives the same library as the Python mirror: the offline driver (examples/, the reference's
initialize_trackers + cloud_cb without ROS) must print exactly the poses the Python path computes."""
import subprocess

import numpy as np
import pytest

from pcl_tracking_b200 import pcl, synth
from tests.test_shim import build_driver

pytestmark = pytest.mark.gpu


def test_offline_driver_matches_python_path(tmp_path):
    exe = build_driver()
    objs = synth.default_objects(1)
    frames = [synth.render(f, objs)[0] for f in range(3)]
    pts0, oid0 = synth.render(0, objs)
    raw_model = synth.model_points(pts0, oid0, 0)
    (tmp_path / "model.raw").write_bytes(synth.to_pcl32(raw_model).tobytes())
    names = []
    for k, f in enumerate(frames):
        p = tmp_path / ("frame%d.raw" % k)
        p.write_bytes(synth.to_pcl32(f).tobytes())
        names.append(str(p))
    out = subprocess.run([exe, str(tmp_path / "model.raw"), "4242"] + names, capture_output=True, text=True, check=True).stdout
    lines = [l.split() for l in out.strip().splitlines()]
    assert len(lines) == 3
    # single-process multi-device mode through the shim (PFT_DEVICES; the ranks share GPU 0 when there is only one):
    # the same one tracker object, the same poses
    import os
    import torch
    devs = "0,1" if torch.cuda.device_count() >= 2 else "0,0"
    out_md = subprocess.run([exe, str(tmp_path / "model.raw"), "4242"] + names, capture_output=True, text=True, check=True,
                            env=dict(os.environ, PFT_DEVICES=devs)).stdout
    assert out_md == out
    # the same pipeline through the Python mirror
    model, c = pcl.prepare_model(pcl.PointCloud(raw_model), 0.01)
    t = pcl.KLDAdaptiveParticleFilterOMPTracker(16)
    pcl.configure_like_reference(t)
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = c
    t.setTrans(m)
    t.seed(4242)
    t.setReferenceCloud(model)
    t.setMinIndices(model.size() // 2)
    for k, f in enumerate(frames):
        pt = pcl.PassThrough()
        pt.setFilterFieldName("z"); pt.setFilterLimits(0.0, 10.0); pt.setKeepOrganized(False)
        pt.setInputCloud(pcl.PointCloud(f))
        passed = pt.filter()
        vg = pcl.ApproximateVoxelGrid()
        vg.setLeafSize(0.01, 0.01, 0.01)
        vg.setInputCloud(passed)
        ds = vg.filter()
        t.setInputCloud(ds)
        t.compute()
        r = t.getResult()
        assert int(lines[k][1]) == len(t.getParticles())
        got = np.array([float(v) for v in lines[k][2:8]], dtype=np.float32)
        want = np.array([r[n] for n in ("x", "y", "z", "roll", "pitch", "yaw")], dtype=np.float32)
        assert np.array_equal(got, want), (k, got, want)


def test_offline_driver_reads_pcd_files(tmp_path):
    """BASELINE configs[0]: auto_tracking offline from PCD -- the driver fed with .pcd files (ASCII model as the model
    builder writes it, binary frames) prints the same poses as with raw records."""
    exe = build_driver()
    objs = synth.default_objects(1)
    frames = [synth.render(f, objs)[0] for f in range(2)]
    pts0, oid0 = synth.render(0, objs)
    raw_model = synth.model_points(pts0, oid0, 0)
    (tmp_path / "model.raw").write_bytes(synth.to_pcl32(raw_model).tobytes())
    pcl.savePCDFile(tmp_path / "model.pcd", raw_model, binary=False)      # ref: src/create_model.cpp:223 (ASCII)
    raw_names, pcd_names = [], []
    for k, f in enumerate(frames):
        (tmp_path / ("frame%d.raw" % k)).write_bytes(synth.to_pcl32(f).tobytes())
        pcl.savePCDFile(tmp_path / ("frame%d.pcd" % k), f, binary=True)
        raw_names.append(str(tmp_path / ("frame%d.raw" % k)))
        pcd_names.append(str(tmp_path / ("frame%d.pcd" % k)))
    a = subprocess.run([exe, str(tmp_path / "model.raw"), "77"] + raw_names, capture_output=True, text=True, check=True).stdout
    b = subprocess.run([exe, str(tmp_path / "model.pcd"), "77"] + pcd_names, capture_output=True, text=True, check=True).stdout
    assert a == b and len(a.strip().splitlines()) == 2
    # and the Python loader gives the device cloud back bit for bit
    back = pcl.loadPCDFile(tmp_path / "frame0.pcd").to_numpy()
    assert back.tobytes() == np.ascontiguousarray(frames[0]).tobytes()
