// pft_cluster.cu -- model acquisition: Euclidean cluster extraction on the GPU.
//
// Replaces pcl::EuclideanClusterExtraction as used by the model builder (ref: src/create_model.cpp:169-179:
// KdTree + setClusterTolerance(0.02) + setMinClusterSize(500) + setMaxClusterSize(25000) + extract), whose clusters
// become the reference clouds of the trackers (ref: src/create_model.cpp:209-230, src/auto_tracking.cpp:749-765).
// Upstream (PCL-1.8.0 segmentation/impl/extract_clusters.hpp, extractEuclideanClusters) grows every cluster
// breadth first with KdTree radius searches (FLANN: squared distance < (float)(tolerance^2), distance accumulated as
// (dx^2 + dy^2) + dz^2 in fp32): a cluster is a connected component of the graph "closer than the tolerance".
// Here: counting sort into a uniform grid (cell >= tolerance), lock-free union-find over the 27-cell neighbourhoods
// (the larger root is always hooked under the smaller one, so a component's root is its lowest point index = the seed
// upstream starts it from), component sizes, then the components inside [min, max] ordered by size descending
// (upstream: std::sort over reverse iterators; ties are unspecified there, here the lower seed index goes first).
#include <float.h>

#include <algorithm>

#include "pft_internal.h"

namespace pft {
namespace {

struct ClusterGrid {
  float origin[3];
  float inv_cell;
  int dim[3];
  int n_cells;
  int valid;
};

constexpr int kMaxClusterCells = 1 << 22;
constexpr int kMaxClusters = 4096;  // clusters that can pass the size filter (each holds >= min_size points)

__device__ __forceinline__ bool finite3(const float4& p) { return isfinite(p.x) && isfinite(p.y) && isfinite(p.z); }

// bounding box of the finite points -> grid header (one block)
__global__ void __launch_bounds__(1024) cl_grid_kernel(const float4* __restrict__ pts, const CloudHeader* __restrict__ hdr, float cell, ClusterGrid* g) {
  __shared__ float s_mn[32][3], s_mx[32][3];
  const int n = hdr->n;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float4 p = pts[i];
    if (!finite3(p)) continue;
    mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
    mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
  }
  for (int d = 0; d < 3; ++d) { mn[d] = warp_min(mn[d]); mx[d] = warp_max(mx[d]); }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) for (int d = 0; d < 3; ++d) { s_mn[wid][d] = mn[d]; s_mx[wid][d] = mx[d]; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) for (int d = 0; d < 3; ++d) { mn[d] = fminf(mn[d], s_mn[w][d]); mx[d] = fmaxf(mx[d], s_mx[w][d]); }
    ClusterGrid h;
    h.valid = mn[0] <= mx[0] ? 1 : 0;
    // cell edge: the tolerance, doubled until the grid fits (a larger cell only means more candidates per neighbourhood)
    float c = cell;
    for (int it = 0; it < 40; ++it) {
      long long cells = 1;
      for (int d = 0; d < 3; ++d) {
        h.dim[d] = h.valid ? (int)fminf(floorf((mx[d] - mn[d]) / c), 2.0e9f) + 1 : 1;
        cells *= (long long)h.dim[d];
        if (cells > (1ll << 40)) cells = 1ll << 40;
      }
      if (cells <= (long long)kMaxClusterCells) { h.n_cells = (int)cells; break; }
      c *= 2.0f;
    }
    for (int d = 0; d < 3; ++d) h.origin[d] = h.valid ? mn[d] : 0.f;
    h.inv_cell = 1.0f / c;
    *g = h;
  }
}

__device__ __forceinline__ void cl_cell_of(const ClusterGrid& g, const float4& p, int& cx, int& cy, int& cz) {
  cx = min(max((int)floorf((p.x - g.origin[0]) * g.inv_cell), 0), g.dim[0] - 1);
  cy = min(max((int)floorf((p.y - g.origin[1]) * g.inv_cell), 0), g.dim[1] - 1);
  cz = min(max((int)floorf((p.z - g.origin[2]) * g.inv_cell), 0), g.dim[2] - 1);
}

__global__ void cl_count_kernel(const float4* __restrict__ pts, const CloudHeader* __restrict__ hdr, const ClusterGrid* __restrict__ gp, int* cell_count,
                                int* __restrict__ parent, int* __restrict__ comp_size) {
  const ClusterGrid g = *gp;
  const int n = hdr->n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    comp_size[i] = 0;
    if (!finite3(p)) { parent[i] = -1; continue; }  // never part of a cluster (the model builder drops NaNs first, ref: src/create_model.cpp:59-77)
    parent[i] = i;
    int cx, cy, cz;
    cl_cell_of(g, p, cx, cy, cz);
    atomicAdd(&cell_count[(cz * g.dim[1] + cy) * g.dim[0] + cx], 1);
  }
}

__global__ void __launch_bounds__(1024) cl_scan_kernel(const ClusterGrid* __restrict__ gp, int* cell_count, int* __restrict__ cell_start) {
  __shared__ int smem[34];
  const int nc = gp->n_cells;
  const int total = block_exclusive_scan<int>(
      nc, [&](int i) { return cell_count[i]; }, [&](int i, int ex) { cell_start[i] = ex; cell_count[i] = 0; }, smem);
  if (threadIdx.x == 0) cell_start[nc] = total;
}

__global__ void cl_scatter_kernel(const float4* __restrict__ pts, const CloudHeader* __restrict__ hdr, const ClusterGrid* __restrict__ gp,
                                  const int* __restrict__ cell_start, int* cell_count, int* __restrict__ order) {
  const ClusterGrid g = *gp;
  const int n = hdr->n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    if (!finite3(p)) continue;
    int cx, cy, cz;
    cl_cell_of(g, p, cx, cy, cz);
    const int c = (cz * g.dim[1] + cy) * g.dim[0] + cx;
    order[cell_start[c] + atomicAdd(&cell_count[c], 1)] = i;
  }
}

__device__ __forceinline__ int cl_find(int* parent, int i) {
  // path halving; parent links only ever point to lower indices, so concurrent hooks keep this loop finite
  int p = parent[i];
  while (p != i) {
    const int gp = parent[p];
    if (gp != p) parent[i] = gp;
    i = p; p = gp;
  }
  return i;
}
__device__ __forceinline__ void cl_unite(int* parent, int a, int b) {
  while (true) {
    a = cl_find(parent, a); b = cl_find(parent, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }   // hook the larger root under the smaller one
    const int old = atomicCAS(&parent[a], a, b);
    if (old == a) return;
    a = old;
  }
}

// every point against the points of its 27 neighbouring cells with a lower index (each pair once)
__global__ void cl_union_kernel(const float4* __restrict__ pts, const CloudHeader* __restrict__ hdr, const ClusterGrid* __restrict__ gp,
                                const int* __restrict__ cell_start, const int* __restrict__ order, int* parent, float tol2) {
  const ClusterGrid g = *gp;
  const int n = hdr->n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    if (!finite3(p)) continue;
    int cx, cy, cz;
    cl_cell_of(g, p, cx, cy, cz);
    for (int z = max(cz - 1, 0); z <= min(cz + 1, g.dim[2] - 1); ++z)
      for (int y = max(cy - 1, 0); y <= min(cy + 1, g.dim[1] - 1); ++y) {
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dim[0] - 1);
        const int base = (z * g.dim[1] + y) * g.dim[0];
        const int s1 = cell_start[base + x1 + 1];
        for (int s = cell_start[base + x0]; s < s1; ++s) {  // cells are x-fastest: the three cells of a row are one slot range
          const int j = order[s];
          if (j >= i) continue;
          const float4 q = pts[j];
          const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
          const float d2 = (dx * dx + dy * dy) + dz * dz;
          if (d2 < tol2) cl_unite(parent, i, j);
        }
      }
  }
}

__global__ void cl_flatten_kernel(const CloudHeader* __restrict__ hdr, int* parent, int* comp_size) {
  const int n = hdr->n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (parent[i] < 0) continue;
    int r = i;
    while (parent[r] != r) r = parent[r];
    atomicAdd(&comp_size[r], 1);
    // (roots are final: all hooks happened in the previous kernel; writing the root back here would race with the
    //  walks of other threads only benignly, but labels are taken from a fresh walk in cl_label_kernel anyway)
  }
}

// the components inside [min, max], ordered by size descending then seed ascending (one block)
__global__ void __launch_bounds__(1024) cl_select_kernel(const CloudHeader* __restrict__ hdr, const int* __restrict__ parent, const int* __restrict__ comp_size,
                                                         int min_size, int max_size, unsigned long long* keys /* [kMaxClusters] */, int* __restrict__ rank_of_root,
                                                         int* __restrict__ out_sizes, int* out_count /* [0] clusters kept, [1] clusters that passed (may exceed the cap) */) {
  __shared__ int smem[34];
  __shared__ unsigned long long s_key[kMaxClusters];
  const int n = hdr->n;
  auto is_kept = [&](int i) -> int { return (parent[i] == i && comp_size[i] >= min_size && comp_size[i] <= max_size) ? 1 : 0; };
  for (int k = threadIdx.x; k < kMaxClusters; k += blockDim.x) s_key[k] = ~0ull;
  __syncthreads();
  const int total = block_exclusive_scan<int>(
      n, is_kept,
      [&](int i, int ex) {
        if (is_kept(i) && ex < kMaxClusters) s_key[ex] = ((unsigned long long)(0xffffffffu - (unsigned int)comp_size[i]) << 32) | (unsigned int)i;
      },
      smem);
  __syncthreads();
  const int kept = min(total, kMaxClusters);
  // bitonic sort of the (padded) key array in shared memory
  for (int k = 2; k <= kMaxClusters; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < kMaxClusters; t += blockDim.x) {
        const int ixj = t ^ j;
        if (ixj > t) {
          const unsigned long long a = s_key[t], b = s_key[ixj];
          const bool up = (t & k) == 0;
          if ((a > b) == up) { s_key[t] = b; s_key[ixj] = a; }
        }
      }
      __syncthreads();
    }
  for (int k = threadIdx.x; k < kept; k += blockDim.x) {
    const unsigned long long key = s_key[k];
    const int root = (int)(key & 0xffffffffu);
    keys[k] = key;
    rank_of_root[root] = k;
    out_sizes[k] = (int)(0xffffffffu - (unsigned int)(key >> 32));
  }
  if (threadIdx.x == 0) { out_count[0] = kept; out_count[1] = total; }
}

__global__ void cl_label_kernel(const CloudHeader* __restrict__ hdr, const int* __restrict__ parent, const int* __restrict__ comp_size, int min_size,
                                int max_size, const int* __restrict__ rank_of_root, int* __restrict__ labels) {
  const int n = hdr->n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int lab = -1;
    if (parent[i] >= 0) {
      int r = i;
      while (parent[r] != r) r = parent[r];
      if (comp_size[r] >= min_size && comp_size[r] <= max_size) lab = rank_of_root[r];
    }
    labels[i] = lab;
  }
}

// the points carrying one label, in input order (= the cluster's sorted indices upstream): one block
__global__ void __launch_bounds__(1024) cl_extract_kernel(const float4* __restrict__ in, const CloudHeader* __restrict__ in_hdr, const int* __restrict__ labels,
                                                          int label, float4* __restrict__ out, CloudHeader* out_hdr) {
  __shared__ int smem[34];
  const int n = in_hdr->n;
  const int total = block_exclusive_scan<int>(
      n, [&](int i) { return labels[i] == label ? 1 : 0; }, [&](int i, int ex) { if (labels[i] == label) out[ex] = in[i]; }, smem);
  if (threadIdx.x == 0) out_hdr->n = total;
}

inline int cl_grid_for(size_t n, int sm) {
  size_t g = (n + 255) / 256;
  if (g > (size_t)sm * 8) g = (size_t)sm * 8;
  return (int)std::max<size_t>(g, 1);
}

}  // namespace
}  // namespace pft

using namespace pft;

extern "C" {

int pft_euclidean_clusters(pft_context* ctx, const pft_cloud* in, double tolerance, int min_size, int max_size, int32_t* labels, size_t labels_capacity,
                           int32_t* sizes, size_t sizes_capacity, size_t* n_clusters) {
  if (!ctx || !in || !n_clusters) { set_last_error("pft_euclidean_clusters: null argument"); return PFT_ERR_INVALID; }
  if (in->ctx != ctx) { set_last_error("pft_euclidean_clusters: cloud belongs to another context"); return PFT_ERR_INVALID; }
  if (!(tolerance > 0.0) || min_size < 1 || max_size < min_size) { set_last_error("pft_euclidean_clusters: need tolerance > 0 and 1 <= min_size <= max_size"); return PFT_ERR_INVALID; }
  PFT_CUDA_TRY(cudaSetDevice(ctx->device));
  size_t n = 0;
  int rc = pft_cloud_size(const_cast<pft_cloud*>(in), &n);
  if (rc) return rc;
  *n_clusters = 0;
  ctx->cl_n = n;
  ctx->cl_src = in;
  if (n == 0) return PFT_OK;
  if (labels && labels_capacity < n) { set_last_error("pft_euclidean_clusters: labels capacity %zu < %zu points", labels_capacity, n); return PFT_ERR_CAPACITY; }
  cudaStream_t s = ctx->stream;
  if ((rc = ctx->cl_grid.reserve(sizeof(ClusterGrid)))) return rc;
  if ((rc = ctx->cl_cells.reserve(((size_t)kMaxClusterCells + 16) * sizeof(int) * 2))) return rc;
  if ((rc = ctx->cl_work.reserve(n * sizeof(int) * 5 + 64))) return rc;
  if ((rc = ctx->cl_sel.reserve((size_t)kMaxClusters * (sizeof(unsigned long long) + sizeof(int)) + 64))) return rc;
  int* cell_count = ctx->cl_cells.as<int>();
  int* cell_start = cell_count + kMaxClusterCells + 16;
  int* parent = ctx->cl_work.as<int>();
  int* comp_size = parent + n;
  int* order = comp_size + n;
  int* rank_of_root = order + n;
  int* dlabels = rank_of_root + n;
  unsigned long long* keys = ctx->cl_sel.as<unsigned long long>();
  int* out_sizes = reinterpret_cast<int*>(keys + kMaxClusters);
  int* out_count = out_sizes + kMaxClusters;
  ClusterGrid* g = ctx->cl_grid.as<ClusterGrid>();
  const float tol2 = (float)(tolerance * tolerance);  // KdTreeFLANN::radiusSearch passes (float)(radius * radius)
  const int grid = cl_grid_for(n, ctx->sm_count);
  cl_grid_kernel<<<1, 1024, 0, s>>>(in->d_pts(), in->d_hdr(), (float)tolerance * 1.0001f, g);
  PFT_LAUNCH_CHECK();
  PFT_CUDA_TRY(cudaMemsetAsync(cell_count, 0, ((size_t)kMaxClusterCells + 16) * sizeof(int), s));
  cl_count_kernel<<<grid, 256, 0, s>>>(in->d_pts(), in->d_hdr(), g, cell_count, parent, comp_size);
  PFT_LAUNCH_CHECK();
  cl_scan_kernel<<<1, 1024, 0, s>>>(g, cell_count, cell_start);
  PFT_LAUNCH_CHECK();
  cl_scatter_kernel<<<grid, 256, 0, s>>>(in->d_pts(), in->d_hdr(), g, cell_start, cell_count, order);
  PFT_LAUNCH_CHECK();
  cl_union_kernel<<<grid, 256, 0, s>>>(in->d_pts(), in->d_hdr(), g, cell_start, order, parent, tol2);
  PFT_LAUNCH_CHECK();
  cl_flatten_kernel<<<grid, 256, 0, s>>>(in->d_hdr(), parent, comp_size);
  PFT_LAUNCH_CHECK();
  cl_select_kernel<<<1, 1024, 0, s>>>(in->d_hdr(), parent, comp_size, min_size, max_size, keys, rank_of_root, out_sizes, out_count);
  PFT_LAUNCH_CHECK();
  cl_label_kernel<<<grid, 256, 0, s>>>(in->d_hdr(), parent, comp_size, min_size, max_size, rank_of_root, dlabels);
  PFT_LAUNCH_CHECK();
  int counts[2] = {0, 0};
  PFT_CUDA_TRY(cudaMemcpyAsync(counts, out_count, sizeof(counts), cudaMemcpyDeviceToHost, s));
  PFT_CUDA_TRY(cudaStreamSynchronize(s));
  if (counts[1] > kMaxClusters) { set_last_error("pft_euclidean_clusters: %d clusters pass the size filter, at most %d are supported", counts[1], kMaxClusters); return PFT_ERR_CAPACITY; }
  *n_clusters = (size_t)counts[0];
  ctx->cl_count = counts[0];
  if (sizes) {
    if (sizes_capacity < (size_t)counts[0]) { set_last_error("pft_euclidean_clusters: sizes capacity %zu < %d clusters", sizes_capacity, counts[0]); return PFT_ERR_CAPACITY; }
    if (counts[0]) PFT_CUDA_TRY(cudaMemcpy(sizes, out_sizes, (size_t)counts[0] * sizeof(int), cudaMemcpyDeviceToHost));
  }
  if (labels) PFT_CUDA_TRY(cudaMemcpy(labels, dlabels, n * sizeof(int), cudaMemcpyDeviceToHost));
  return PFT_OK;
}

int pft_cloud_select_cluster(pft_context* ctx, const pft_cloud* in, int k, pft_cloud* out) {
  if (!ctx || !in || !out) { set_last_error("pft_cloud_select_cluster: null argument"); return PFT_ERR_INVALID; }
  if (in->ctx != ctx || out->ctx != ctx || in == out) { set_last_error("pft_cloud_select_cluster: clouds must be distinct and belong to the context"); return PFT_ERR_INVALID; }
  if (ctx->cl_src != in || !ctx->cl_work.p) { set_last_error("pft_cloud_select_cluster: call pft_euclidean_clusters on this cloud first"); return PFT_ERR_STATE; }
  if (k < 0 || k >= ctx->cl_count) { set_last_error("pft_cloud_select_cluster: cluster %d of %d", k, ctx->cl_count); return PFT_ERR_INVALID; }
  PFT_CUDA_TRY(cudaSetDevice(ctx->device));
  int rc = out->ensure(in->capacity);
  if (rc) return rc;
  const int* dlabels = ctx->cl_work.as<int>() + 4 * ctx->cl_n;
  cl_extract_kernel<<<1, 1024, 0, ctx->stream>>>(in->d_pts(), in->d_hdr(), dlabels, k, out->d_pts(), out->d_hdr());
  PFT_LAUNCH_CHECK();
  out->host_n = -1;
  return PFT_OK;
}

}  // extern "C"
