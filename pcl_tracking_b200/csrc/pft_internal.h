// pft_internal.h -- host-side objects behind the opaque C handles and the kernel launchers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <vector>

#include "../../include/pft/pft.h"
#include "pft_common.cuh"

namespace pft {

// Grow-only device buffer.
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int reserve(size_t need);
  void release();
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace pft

struct pft_context {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  // K1 scratch
  pft::DevBuf k1_keys, k1_first, k1_vid, k1_slot_of, k1_acc_xyz, k1_acc_rgbc, k1_blk, staging, tmp_cloud_pts, tmp_hdr, tmp_f;
  void* pinned = nullptr;  // small pinned buffer for scalar read-backs
  size_t pinned_bytes = 0;
  // NCCL communicator shared by the clouds and trackers of this context (one process per GPU)
  void* comm = nullptr;
  int nranks = 1, rank = 0;
  // captured downsample sequences (pft_filters.cu: run_voxel_grid), keyed by every pointer / size / parameter they use
  struct K1Key { const void* p[11]; size_t cap, H; float leaf, lo, hi; int field; };
  static constexpr int kK1Graphs = 8;
  K1Key k1_key[kK1Graphs];
  cudaGraphExec_t k1_exec[kK1Graphs] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int k1_next = 0;
  // Euclidean clustering scratch (pft_cluster.cu); the labels of the last call stay on the device for pft_cloud_select_cluster
  pft::DevBuf cl_grid, cl_cells, cl_work, cl_sel;
  size_t cl_n = 0;
  int cl_count = 0;
  const pft_cloud* cl_src = nullptr;
  cudaEvent_t batch_fork = nullptr;  // pft_compute_batch: the point of the context stream the trackers' streams fork from
  // frame ingest overlapped with compute (pft_cloud_upload_async): copies and unpack kernels run on a stream of their own
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copy_fence = nullptr;  // "everything enqueued on `stream` so far": the copy must not overtake earlier readers
};

struct pft_cloud {
  pft_context* ctx = nullptr;
  pft::DevBuf pts;              // float4[capacity]
  pft::DevBuf hdr;              // CloudHeader
  size_t capacity = 0;          // host-known upper bound on n
  long long host_n = -1;        // exact n if known on the host, else -1
  // pft_cloud_upload_async: `ready` fires on the copy stream once the points are in place; the first consumer on the
  // context stream waits for it (join_upload)
  pft::DevBuf raw_staging;      // raw PointCloud2 / 32-byte records of an asynchronous upload
  cudaEvent_t ready = nullptr;
  mutable bool upload_pending = false;
  int join_upload() const;
  // scene distribution by peer stores (pft_cloud_peer_*): once exported the storage never moves
  bool peer_exported = false, peer_attached = false;
  size_t peer_capacity = 0;
  int peer_nranks = 0, peer_rank = 0;
  pft::DevBuf peer_sync;         // {flag, blocks done}
  void* peer_pts[16] = {nullptr}; void* peer_hdr[16] = {nullptr}; void* peer_flag[16] = {nullptr};
  unsigned int* peer_error = nullptr;  // host-mapped
  unsigned int peer_epoch = 0;
  float4* d_pts() const { return pts.as<float4>(); }
  pft::CloudHeader* d_hdr() const { return hdr.as<pft::CloudHeader>(); }
  int ensure(size_t cap);
};

namespace pft {

// ---- filters (pft_filters.cu)
int launch_unpack_pcl32(cudaStream_t s, const void* src32, float4* dst, size_t n);
int launch_pack_pcl32(cudaStream_t s, const float4* src, void* dst32, size_t n);
int launch_unpack_pointcloud2(cudaStream_t s, const void* src, float4* dst, unsigned int width, unsigned int height, unsigned int point_step,
                              unsigned int row_step, int off_x, int off_y, int off_z, int off_rgb);
int launch_set_header(cudaStream_t s, CloudHeader* hdr, int n);
int run_passthrough(pft_context* ctx, const pft_cloud* in, pft_cloud* out, int field, float lo, float hi, int drop_zero);
int run_voxel_grid(pft_context* ctx, const pft_cloud* in, pft_cloud* out, float leaf, int field, float lo, float hi);
int run_approx_voxel_grid_pcl(pft_context* ctx, const pft_cloud* in, pft_cloud* out, float leaf, int field, float lo, float hi);
int run_centre_on_centroid(pft_context* ctx, pft_cloud* cloud, float* d_centroid3);

}  // namespace pft
