"""The C++ PCL-named shim (include/pft/pcl_shim.hpp) and the offline driver compile with plain g++ against
libpft.so; without a GPU the driver fails loudly (exit 3, "no CPU fallback")."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "auto_tracking_offline")


def build_driver():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples"), "-s", "CXX=/usr/bin/g++"])
    return EXE


def test_shim_compiles_and_fails_loudly_without_gpu(tmp_path):
    exe = build_driver()
    from pcl_tracking_b200 import _capi
    if _capi.load().pft_device_count() > 0:
        pytest.skip("a CUDA device is present")
    f = tmp_path / "x.raw"
    f.write_bytes(b"\0" * 64)
    r = subprocess.run([exe, str(f), "1", str(f)], capture_output=True, text=True)
    assert r.returncode == 3
    assert "no CPU fallback" in r.stderr


def test_shim_templates_instantiate(tmp_path):
    """Every PCL-named class of the shim, including the model-acquisition / ingest / read-out additions, compiles
    when instantiated the way the reference uses it (g++ -fsyntax-only: no GPU needed)."""
    src = tmp_path / "shim_check.cpp"
    src.write_text(r'''
#include <pft/pcl_shim.hpp>
int main() {
  typedef pcl::PointXYZRGBA P;
  pcl::PointCloud<P>::Ptr c(new pcl::PointCloud<P>());
  c->fromPointCloud2(nullptr, 0, 0, 32, 0, 0, 4, 8, 16);                       // ref: src/auto_tracking.cpp:619-622
  pcl::PassThrough<P> pass; pass.setFilterFieldName("z"); pass.setFilterLimits(0.0, 10.0); pass.setKeepOrganized(false);
  pass.setInputCloud(c);
  pcl::PointCloud<P> out; pass.filter(out);
  pcl::EuclideanClusterExtraction<P> ec;                                        // ref: src/create_model.cpp:169-179
  ec.setClusterTolerance(0.02); ec.setMinClusterSize(500); ec.setMaxClusterSize(25000);
  ec.setSearchMethod(0); ec.setInputCloud(c);
  std::vector<pcl::PointIndices> clusters; ec.extract(clusters);
  ec.clusterCloud(0, out);
  pcl::tracking::KLDAdaptiveParticleFilterOMPTracker<P, pcl::tracking::ParticleXYZRPY> t(8);
  pft_result_box b = t.getResultBox();                                          // ref: src/auto_tracking.cpp:432-466
  // overlapped ingest, the parity search of the Approx coherence, the change detector
  c->fromPointCloud2Async(nullptr, 0, 0, 32, 0, 0, 4, 8, 16);
  std::shared_ptr<pcl::tracking::ApproxNearestPairPointCloudCoherence<P>> coh(new pcl::tracking::ApproxNearestPairPointCloudCoherence<P>());
  std::shared_ptr<pcl::tracking::DistanceCoherence<P>> dc(new pcl::tracking::DistanceCoherence<P>());
  coh->addPointCoherence(dc);
  coh->setPclApproximateSearch(true);
  coh->setMaximumDistance(0.01);
  t.setCloudCoherence(coh);
  t.setUseChangeDetector(true); t.setIntervalOfChangeDetection(10); t.setMinPointsOfChangeDetection(10); t.setResolutionOfChangeDetection(0.01);
  return b.n;
}
''')
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])
