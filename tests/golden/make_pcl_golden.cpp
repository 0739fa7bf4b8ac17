// make_pcl_golden.cpp -- pins the CPU oracle (oracle/pft_oracle.cpp) to REAL PCL 1.8.0.
//
// This repository's oracle is a restatement of PCL 1.8.0 (the arithmetic of the reference's hot path lives in that
// un-vendored dependency: find_package(PCL 1.8.0 EXACT), ref: CMakeLists.txt:4); PCL is not installable in the build
// image, so every parity claim is "against the restatement" until this program has been run once on a machine that
// has PCL 1.8.0.  It calls the real pcl::tracking / pcl::ApproximateVoxelGrid / pcl::VoxelGrid / pcl::PassThrough /
// pcl::search::Octree code on the inputs of tests/golden/oracle_small_case.npz and oracle_parity_modes_case.npz and
// writes what PCL computes as .npy files; tests/golden/pack_pcl_golden.py zips them into tests/golden/pcl_small_case.npz,
// which tests/test_golden.py::test_oracle_matches_pcl_golden_when_present consumes.  One command pins the oracle:
//
//   python tests/golden/export_golden_inputs.py /tmp/pclgold            # scene / model / particles as raw binaries
//   g++ -O2 -std=c++11 tests/golden/make_pcl_golden.cpp -o /tmp/make_pcl_golden \
//       $(pkg-config --cflags --libs pcl_tracking-1.8 pcl_filters-1.8 pcl_search-1.8 pcl_octree-1.8 pcl_common-1.8)
//   /tmp/make_pcl_golden /tmp/pclgold && python tests/golden/pack_pcl_golden.py /tmp/pclgold
//
// What is pinned (the deterministic stages; resample() draws from mt19937(time(0)) upstream and cannot be):
//   weight():   transformPointCloud per particle, calcBoundingBox, cropInputPointCloud, NearestPairPointCloudCoherence with
//               DistanceCoherence + HSVColorCoherence(0.1) over pcl::search::Octree(0.01), maximum distance 0.1 (the
//               reference's configuration, ref: src/auto_tracking.cpp:235-253, exact-search variant :237-238) -> raw
//               weights; normalizeWeight() -> weights; update() -> representative state
//   approximate coherence: ApproxNearestPairPointCloudCoherence on the same inputs -> raw weights; and
//               OctreePointCloudSearch::approxNearestSearch index / distance per (particle, model point) for 4 particles
//   filters:    PassThrough(z in [0,10]), ApproximateVoxelGrid(0.02), VoxelGrid(0.02) of the scene
//   scalars:    calcKLBound(k) for k = 2..200 at delta 0.99, epsilon 0.2 (through a subclass), genAliasTable of the weights
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

#include <pcl/common/transforms.h>
#include <pcl/filters/approximate_voxel_grid.h>
#include <pcl/filters/passthrough.h>
#include <pcl/filters/voxel_grid.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/search/octree.h>
#include <pcl/tracking/approx_nearest_pair_point_cloud_coherence.h>
#include <pcl/tracking/distance_coherence.h>
#include <pcl/tracking/hsv_color_coherence.h>
#include <pcl/tracking/kld_adaptive_particle_filter.h>
#include <pcl/tracking/nearest_pair_point_cloud_coherence.h>
#include <pcl/tracking/tracking.h>

typedef pcl::PointXYZRGBA P;
typedef pcl::tracking::ParticleXYZRPY S;
typedef pcl::PointCloud<P> Cloud;

struct Packed16 { float x, y, z; uint32_t rgba; };                       // the repository's 16-byte point
struct Particle32 { float x, y, z, one, roll, pitch, yaw, weight; };     // = ParticleXYZRPY

template <typename T>
static std::vector<T> read_bin(const std::string& path) {
  std::ifstream f(path.c_str(), std::ios::binary);
  if (!f) { std::fprintf(stderr, "cannot read %s\n", path.c_str()); std::exit(2); }
  f.seekg(0, std::ios::end);
  const size_t n = (size_t)f.tellg() / sizeof(T);
  f.seekg(0);
  std::vector<T> v(n);
  f.read(reinterpret_cast<char*>(v.data()), n * sizeof(T));
  return v;
}

// minimal .npy (v1.0) writer: C-order array of `descr` elements
static void write_npy(const std::string& path, const char* descr, const std::vector<size_t>& shape, const void* data, size_t bytes) {
  std::string hdr = std::string("{'descr': '") + descr + "', 'fortran_order': False, 'shape': (";
  for (size_t i = 0; i < shape.size(); ++i) hdr += std::to_string(shape[i]) + (shape.size() == 1 || i + 1 < shape.size() ? "," : "");
  hdr += "), }";
  while ((10 + hdr.size() + 1) % 64) hdr += ' ';
  hdr += '\n';
  std::ofstream f(path.c_str(), std::ios::binary);
  const unsigned char magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
  f.write(reinterpret_cast<const char*>(magic), 8);
  const unsigned short hl = (unsigned short)hdr.size();
  f.write(reinterpret_cast<const char*>(&hl), 2);
  f.write(hdr.data(), hdr.size());
  f.write(reinterpret_cast<const char*>(data), bytes);
}

static Cloud::Ptr to_cloud(const std::vector<Packed16>& v) {
  Cloud::Ptr c(new Cloud());
  c->points.resize(v.size());
  for (size_t i = 0; i < v.size(); ++i) { c->points[i].x = v[i].x; c->points[i].y = v[i].y; c->points[i].z = v[i].z; c->points[i].rgba = v[i].rgba; }
  c->width = (uint32_t)v.size(); c->height = 1; c->is_dense = false;
  return c;
}
static std::vector<Packed16> from_cloud(const Cloud& c) {
  std::vector<Packed16> v(c.points.size());
  for (size_t i = 0; i < v.size(); ++i) { v[i].x = c.points[i].x; v[i].y = c.points[i].y; v[i].z = c.points[i].z; v[i].rgba = c.points[i].rgba; }
  return v;
}

// the protected stages of the tracker, made callable
struct Exposed : pcl::tracking::KLDAdaptiveParticleFilterTracker<P, S> {
  using pcl::tracking::KLDAdaptiveParticleFilterTracker<P, S>::calcKLBound;
  void setParticleSet(const std::vector<Particle32>& ps) {
    particles_.reset(new PointCloudState());
    for (size_t i = 0; i < ps.size(); ++i) {
      S p; p.x = ps[i].x; p.y = ps[i].y; p.z = ps[i].z; p.roll = ps[i].roll; p.pitch = ps[i].pitch; p.yaw = ps[i].yaw; p.weight = ps[i].weight;
      particles_->points.push_back(p);
    }
    particle_num_ = (int)ps.size();
  }
  bool init() { return initCompute(); }     // allocates transed_reference_vector_, hands the input to the coherence
  // weight() without normalizeWeight(): the same statements as upstream's weight() up to the call of normalizeWeight()
  void rawWeights(std::vector<float>& raw, double box[6], int& cropped) {
    for (size_t i = 0; i < particles_->points.size(); ++i)
      computeTransformedPointCloudWithoutNormal(particles_->points[i], *transed_reference_vector_[i]);
    PointCloudInPtr coherence_input(new PointCloudIn);
    cropInputPointCloud(input_, *coherence_input);
    calcBoundingBox(box[0], box[3], box[1], box[4], box[2], box[5]);
    cropped = (int)coherence_input->points.size();
    coherence_->setTargetCloud(coherence_input);
    coherence_->initCompute();
    raw.resize(particles_->points.size());
    for (size_t i = 0; i < particles_->points.size(); ++i) {
      pcl::IndicesPtr indices;
      coherence_->compute(transed_reference_vector_[i], indices, particles_->points[i].weight);
      raw[i] = particles_->points[i].weight;
    }
  }
  void normalize() { normalizeWeight(); }
  void runUpdate() { update(); }
  std::vector<float> weights() const { std::vector<float> w; for (size_t i = 0; i < particles_->points.size(); ++i) w.push_back(particles_->points[i].weight); return w; }
  void aliasTable(std::vector<int>& a, std::vector<double>& q) { genAliasTable(a, q, particles_); }
};

template <typename CoherenceT>
static void configure(Exposed& t, const Cloud::Ptr& model, const Cloud::Ptr& scene, const std::vector<Particle32>& parts) {
  t.setMaximumParticleNum(96);
  t.setDelta(0.99);
  t.setEpsilon(0.2);
  S bin; bin.x = bin.y = bin.z = bin.roll = bin.pitch = bin.yaw = 0.1f;
  t.setBinSize(bin);
  t.setTrans(Eigen::Affine3f::Identity());
  std::vector<double> step(6, 0.015 * 0.015); step[3] *= 40.0; step[4] *= 40.0; step[5] *= 40.0;
  t.setStepNoiseCovariance(step);
  t.setInitialNoiseCovariance(std::vector<double>(6, 0.00001));
  t.setInitialNoiseMean(std::vector<double>(6, 0.0));
  t.setIterationNum(2);
  t.setParticleNum((int)parts.size());
  t.setResampleLikelihoodThr(0.0);
  t.setUseNormal(false);
  typename CoherenceT::Ptr coherence(new CoherenceT());
  boost::shared_ptr<pcl::tracking::DistanceCoherence<P> > dc(new pcl::tracking::DistanceCoherence<P>());
  coherence->addPointCoherence(dc);
  boost::shared_ptr<pcl::tracking::HSVColorCoherence<P> > hc(new pcl::tracking::HSVColorCoherence<P>());
  hc->setWeight(0.1);
  coherence->addPointCoherence(hc);
  boost::shared_ptr<pcl::search::Octree<P> > search(new pcl::search::Octree<P>(0.01));
  coherence->setSearchMethod(search);
  coherence->setMaximumDistance(0.1);
  t.setCloudCoherence(coherence);
  t.setReferenceCloud(model);
  t.setInputCloud(scene);
  t.setParticleSet(parts);   // before init(): initCompute() only draws particles when the set is empty
  t.init();
}

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s DIR (holding scene.bin model.bin particles.bin)\n", argv[0]); return 2; }
  const std::string dir = std::string(argv[1]) + "/";
  const std::vector<Packed16> scene_v = read_bin<Packed16>(dir + "scene.bin"), model_v = read_bin<Packed16>(dir + "model.bin");
  const std::vector<Particle32> parts = read_bin<Particle32>(dir + "particles.bin");
  Cloud::Ptr scene = to_cloud(scene_v), model = to_cloud(model_v);
  const size_t N = parts.size(), M = model_v.size();

  // ---- exact coherence: raw weights, crop box, normalised weights, update()
  {
    Exposed t;
    configure<pcl::tracking::NearestPairPointCloudCoherence<P> >(t, model, scene, parts);
    std::vector<float> raw; double box[6]; int cropped = 0;
    t.rawWeights(raw, box, cropped);
    float boxf[6]; for (int d = 0; d < 6; ++d) boxf[d] = (float)box[d];
    write_npy(dir + "pcl_raw.npy", "<f4", {N}, raw.data(), N * 4);
    write_npy(dir + "pcl_aabb.npy", "<f4", {6}, boxf, 24);
    write_npy(dir + "pcl_cropped_count.npy", "<i4", {1}, &cropped, 4);
    std::vector<int> a; std::vector<double> q;
    t.normalize();
    std::vector<float> w = t.weights();
    write_npy(dir + "pcl_weights.npy", "<f4", {N}, w.data(), N * 4);
    t.aliasTable(a, q);
    write_npy(dir + "pcl_alias_a.npy", "<i4", {a.size()}, a.data(), a.size() * 4);
    write_npy(dir + "pcl_alias_q.npy", "<f8", {q.size()}, q.data(), q.size() * 8);
    t.runUpdate();
    const S r = t.getResult();
    const float res[8] = {r.x, r.y, r.z, 1.f, r.roll, r.pitch, r.yaw, r.weight};
    write_npy(dir + "pcl_result.npy", "<f4", {8}, res, 32);
    std::vector<double> kl;
    for (int k = 2; k <= 200; ++k) kl.push_back(t.calcKLBound(k));
    write_npy(dir + "pcl_kl_bound.npy", "<f8", {kl.size()}, kl.data(), kl.size() * 8);
  }
  // ---- the reference's own coherence (approximate search): raw weights + the searched pairs of 4 particles
  {
    Exposed t;
    configure<pcl::tracking::ApproxNearestPairPointCloudCoherence<P> >(t, model, scene, parts);
    std::vector<float> raw; double box[6]; int cropped = 0;
    t.rawWeights(raw, box, cropped);
    write_npy(dir + "pcl_approx_raw.npy", "<f4", {N}, raw.data(), N * 4);
    // approxNearestSearch over the cropped cloud, queried with the transformed model of the first particles
    pcl::PassThrough<P> px, py, pz;
    Cloud::Ptr cx(new Cloud()), cy(new Cloud()), cz(new Cloud());
    px.setFilterFieldName("x"); px.setFilterLimits((float)box[0], (float)box[3]); px.setKeepOrganized(false); px.setInputCloud(scene); px.filter(*cx);
    py.setFilterFieldName("y"); py.setFilterLimits((float)box[1], (float)box[4]); py.setKeepOrganized(false); py.setInputCloud(cx); py.filter(*cy);
    pz.setFilterFieldName("z"); pz.setFilterLimits((float)box[2], (float)box[5]); pz.setKeepOrganized(false); pz.setInputCloud(cy); pz.filter(*cz);
    pcl::search::Octree<P> oct(0.01);
    oct.setInputCloud(cz);
    const size_t K = std::min<size_t>(4, N);
    std::vector<int> idx(K * M); std::vector<float> d2(K * M);
    for (size_t i = 0; i < K; ++i) {
      S p; p.x = parts[i].x; p.y = parts[i].y; p.z = parts[i].z; p.roll = parts[i].roll; p.pitch = parts[i].pitch; p.yaw = parts[i].yaw;
      Cloud moved;
      pcl::transformPointCloud(*model, moved, p.toEigenMatrix());
      for (size_t j = 0; j < M; ++j) { int k = -1; float d = 0.f; oct.approxNearestSearch(moved.points[j], k, d); idx[i * M + j] = k; d2[i * M + j] = d; }
    }
    write_npy(dir + "pcl_approx_nn_idx.npy", "<i4", {K, M}, idx.data(), idx.size() * 4);   // indices into the CROPPED cloud
    write_npy(dir + "pcl_approx_nn_d2.npy", "<f4", {K, M}, d2.data(), d2.size() * 4);
    std::vector<Packed16> cropped_pts = from_cloud(*cz);
    write_npy(dir + "pcl_cropped.npy", "|V16", {cropped_pts.size()}, cropped_pts.data(), cropped_pts.size() * 16);
  }
  // ---- filters
  {
    pcl::PassThrough<P> pass;
    pass.setFilterFieldName("z"); pass.setFilterLimits(0, 10); pass.setKeepOrganized(false); pass.setInputCloud(scene);
    Cloud::Ptr passed(new Cloud());
    pass.filter(*passed);
    std::vector<Packed16> v = from_cloud(*passed);
    write_npy(dir + "pcl_passthrough.npy", "|V16", {v.size()}, v.data(), v.size() * 16);
    pcl::ApproximateVoxelGrid<P> ag; ag.setLeafSize(0.02f, 0.02f, 0.02f); ag.setInputCloud(passed);
    Cloud a; ag.filter(a);
    v = from_cloud(a);
    write_npy(dir + "pcl_approx_voxel_grid.npy", "|V16", {v.size()}, v.data(), v.size() * 16);
    pcl::VoxelGrid<P> vg; vg.setLeafSize(0.02f, 0.02f, 0.02f); vg.setInputCloud(passed);
    Cloud g; vg.filter(g);
    v = from_cloud(g);
    write_npy(dir + "pcl_voxel_grid.npy", "|V16", {v.size()}, v.data(), v.size() * 16);
  }
  std::printf("written: %s pcl_*.npy\n", dir.c_str());
  return 0;
}
