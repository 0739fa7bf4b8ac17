// pft_filters.cu -- kernel K1: PassThrough + voxel-grid downsampling as a sort-free voxel-hash
// reduction, plus the order-preserving PassThrough compaction and the model preparation helpers.
//
// Replaces pcl::PassThrough::filter and pcl::(Approximate)VoxelGrid::filter as called from
// ref: src/auto_tracking.cpp:536-575.
//
// Data layout: clouds are float4 {x,y,z,rgba} arrays in HBM with a 16-byte device header holding the
// point count (no host round trip between stages).  The voxel hash is an open-addressing table of
// 64-bit packed lattice keys (21 bits per axis); per-voxel sums are kept in double (coordinates) and
// uint32 (colour bytes), which makes the centroid independent of the order the atomics land in:
// sums of fp32 sensor-range values are exact in fp64.  Output order = order of first appearance.
#include <stdlib.h>
#include <string.h>

#include "pft_internal.h"

namespace pft {

namespace {

constexpr unsigned long long kEmptyKey = ~0ull;
constexpr int kFirstInit = 0x7f7f7f7f;

__device__ __forceinline__ bool finite3(const float4& p) { return isfinite(p.x) && isfinite(p.y) && isfinite(p.z); }

__device__ __forceinline__ bool passes(const float4& p, int field, float lo, float hi) {
  if (!finite3(p)) return false;
  if (field < 0) return true;
  const float v = field == 0 ? p.x : (field == 1 ? p.y : p.z);
  return !(v < lo || v > hi);
}

__device__ __forceinline__ unsigned long long voxel_key(const float4& p, float inv) {
  // lattice of upstream (Approximate)VoxelGrid: floor(coord * inverse_leaf_size)
  const int ix = (int)floorf(p.x * inv), iy = (int)floorf(p.y * inv), iz = (int)floorf(p.z * inv);
  const unsigned long long bx = (unsigned long long)((ix + (1 << 20)) & 0x1fffff);
  const unsigned long long by = (unsigned long long)((iy + (1 << 20)) & 0x1fffff);
  const unsigned long long bz = (unsigned long long)((iz + (1 << 20)) & 0x1fffff);
  return (bx << 42) | (by << 21) | bz;
}
__device__ __forceinline__ unsigned int hash_key(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (unsigned int)k;
}

__global__ void unpack_pcl32_kernel(const uint4* __restrict__ src, float4* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 a = src[2 * i];      // x y z w
    const uint4 b = src[2 * i + 1];  // rgba pad pad pad
    dst[i] = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(b.x));
  }
}
__global__ void pack_pcl32_kernel(const float4* __restrict__ src, uint4* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float4 p = src[i];
    dst[2 * i] = make_uint4(__float_as_uint(p.x), __float_as_uint(p.y), __float_as_uint(p.z), __float_as_uint(1.0f));
    dst[2 * i + 1] = make_uint4(__float_as_uint(p.w), 0u, 0u, 0u);
  }
}
// sensor_msgs/PointCloud2 -> float4 {x,y,z,rgba}: replaces pcl_conversions::toPCL + pcl::fromPCLPointCloud2
// (ref: src/auto_tracking.cpp:619-622).  Rows of `row_step` bytes hold `width` records of `point_step` bytes; the
// float32 fields x, y, z and the 4 colour bytes sit at arbitrary (4-byte aligned) offsets inside a record.  Organised
// clouds keep their row-major order, as fromPCLPointCloud2 does.  A warp reads whole 4-byte words of consecutive
// records, so the loads of a row are as coalesced as the record stride allows.
__global__ void unpack_pointcloud2_kernel(const unsigned int* __restrict__ src, float4* __restrict__ dst, unsigned int width, unsigned int height,
                                          unsigned int point_words, unsigned int row_words, int wx, int wy, int wz, int wrgb) {
  const size_t n = (size_t)width * height;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / width, col = i - row * width;
    const unsigned int* rec = src + row * row_words + col * point_words;
    const unsigned int rgba = wrgb >= 0 ? rec[wrgb] : 0u;
    dst[i] = make_float4(__uint_as_float(rec[wx]), __uint_as_float(rec[wy]), __uint_as_float(rec[wz]), __uint_as_float(rgba));
  }
}
__global__ void set_header_kernel(CloudHeader* h, int n) { h->n = n; }

// ---- PassThrough (order-preserving compaction), one thread block: flags -> exclusive scan -> scatter.
__global__ void __launch_bounds__(1024) passthrough_kernel(const float4* __restrict__ in, const CloudHeader* __restrict__ in_hdr,
                                                           float4* __restrict__ out, CloudHeader* out_hdr, int field, float lo,
                                                           float hi, int drop_zero) {
  __shared__ int smem[34];
  const int n = in_hdr->n;
  auto keep = [&](int i) -> int {
    const float4 p = in[i];
    if (!passes(p, field, lo, hi)) return 0;
    // removeZeroPoints (ref: src/auto_tracking.cpp:577-595): the comparison is made in double upstream
    if (drop_zero && fabs((double)p.x) < 0.01 && fabs((double)p.y) < 0.01 && fabs((double)p.z) < 0.01) return 0;
    return 1;
  };
  const int total = block_exclusive_scan<int>(
      n, keep, [&](int i, int ex) { const float4 p = in[i]; if (keep(i)) out[ex] = p; }, smem);
  if (threadIdx.x == 0) out_hdr->n = total;
}

// ---- K1a: insert every surviving point's voxel key; remember the lowest point index per voxel.
__global__ void k1_insert_kernel(const float4* __restrict__ in, const CloudHeader* __restrict__ in_hdr, unsigned long long* keys,
                                 int* first, int* __restrict__ slot_of, unsigned int mask, float inv, int field, float lo, float hi) {
  const int n = in_hdr->n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = in[i];
    int slot = -1;
    if (passes(p, field, lo, hi)) {
      const unsigned long long key = voxel_key(p, inv);
      unsigned int s = hash_key(key) & mask;
      while (true) {
        const unsigned long long prev = atomicCAS(&keys[s], kEmptyKey, key);
        if (prev == kEmptyKey || prev == key) break;
        s = (s + 1) & mask;
      }
      slot = (int)s;
      atomicMin(&first[s], i);
    }
    slot_of[i] = slot;
  }
}
// ---- K1b: voxel id = rank of the voxel's first point among all first points (output order = order of first
// appearance).  Three small kernels: per-tile counts, one-block scan of the tile counts, per-tile assignment.
constexpr int kTile = 4096;  // points per block (1024 threads x 4)

__device__ __forceinline__ int k1_is_first(int i, int n, const int* __restrict__ slot_of, const int* __restrict__ first) {
  if (i >= n) return 0;
  const int s = slot_of[i];
  return (s >= 0 && first[s] == i) ? 1 : 0;
}
__global__ void __launch_bounds__(1024) k1_tile_count_kernel(const CloudHeader* __restrict__ in_hdr, const int* __restrict__ slot_of,
                                                             const int* __restrict__ first, int* __restrict__ tile_count) {
  __shared__ int red[32];
  const int n = in_hdr->n;
  const int base = blockIdx.x * kTile + threadIdx.x * 4;
  int c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) c += k1_is_first(base + k, n, slot_of, first);
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = red[threadIdx.x];
    v = warp_sum(v);
    if (threadIdx.x == 0) tile_count[blockIdx.x] = v;
  }
}
__global__ void __launch_bounds__(1024) k1_tile_scan_kernel(int n_tiles, int* tile_count, CloudHeader* out_hdr) {
  __shared__ int smem[34];
  const int total = block_exclusive_scan<int>(
      n_tiles, [&](int i) { return tile_count[i]; }, [&](int i, int ex) { tile_count[i] = ex; }, smem);
  if (threadIdx.x == 0) out_hdr->n = total;
}
__global__ void __launch_bounds__(1024) k1_tile_assign_kernel(const CloudHeader* __restrict__ in_hdr, const int* __restrict__ slot_of,
                                                              const int* __restrict__ first, const int* __restrict__ tile_offset,
                                                              int* __restrict__ vid, double* acc_xyz, unsigned int* acc_rgbc) {
  __shared__ int wsum[32];
  const int n = in_hdr->n;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int base = blockIdx.x * kTile + threadIdx.x * 4;
  int f[4], sum = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) { f[k] = k1_is_first(base + k, n, slot_of, first); sum += f[k]; }
  int inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) wsum[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    const int w = wsum[lane];
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(kFull, winc, o); if (lane >= o) winc += t; }
    wsum[lane] = winc - w;
  }
  __syncthreads();
  int ex = tile_offset[blockIdx.x] + wsum[wid] + (inc - sum);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (f[k]) {
      vid[slot_of[base + k]] = ex;  // (the accumulators were cleared by two contiguous memsets before the pass)
      ++ex;
    }
  }
}
// ---- K1c: accumulate.  fp64 sums of fp32 sensor coordinates are exact => order independent.
__global__ void k1_accum_kernel(const float4* __restrict__ in, const CloudHeader* __restrict__ in_hdr, const int* __restrict__ slot_of,
                                const int* __restrict__ vid, double* acc_xyz, unsigned int* acc_rgbc) {
  const int n = in_hdr->n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int s = slot_of[i];
    if (s < 0) continue;
    const int v = vid[s];
    const float4 p = in[i];
    const unsigned int rgba = __float_as_uint(p.w);
    atomicAdd(&acc_xyz[3 * v], (double)p.x);
    atomicAdd(&acc_xyz[3 * v + 1], (double)p.y);
    atomicAdd(&acc_xyz[3 * v + 2], (double)p.z);
    atomicAdd(&acc_rgbc[4 * v], (rgba >> 16) & 0xffu);
    atomicAdd(&acc_rgbc[4 * v + 1], (rgba >> 8) & 0xffu);
    atomicAdd(&acc_rgbc[4 * v + 2], rgba & 0xffu);
    atomicAdd(&acc_rgbc[4 * v + 3], 1u);
  }
}
// ---- K1d: centroid = fp64 quotient rounded to fp32; colour = (int)(float_sum / float_count), alpha 0
__global__ void k1_final_kernel(const CloudHeader* __restrict__ out_hdr, const double* __restrict__ acc_xyz,
                                const unsigned int* __restrict__ acc_rgbc, float4* __restrict__ out) {
  const int n = out_hdr->n;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
    const uint4 c = reinterpret_cast<const uint4*>(acc_rgbc)[v];
    const double cnt = (double)c.w;
    const float fc = (float)c.w;
    const unsigned int r = (unsigned int)(int)((float)c.x / fc), g = (unsigned int)(int)((float)c.y / fc), b = (unsigned int)(int)((float)c.z / fc);
    out[v] = make_float4((float)(acc_xyz[3 * v] / cnt), (float)(acc_xyz[3 * v + 1] / cnt), (float)(acc_xyz[3 * v + 2] / cnt),
                         __uint_as_float((r << 16) | (g << 8) | b));
  }
}

// ---- pcl::ApproximateVoxelGrid<PointXYZRGBA>::applyFilter reproduced exactly (PCL-1.8.0 filters/impl/approximate_voxel_grid.hpp,
// SURVEY A.1; ref: src/auto_tracking.cpp:563-575): a 512-entry direct-mapped cache of voxel accumulators walked in input
// order, an entry flushed (a partial centroid emitted) whenever another voxel hashes onto it, the rest flushed at the end
// in slot order.  Output order, duplicate partial centroids and the sequential fp32 sums all depend on the input order:
// the algorithm is a sequential scan, run here by ONE thread.  Parity mode (pft_approx_voxel_grid_pcl); the product
// path is the exact one-centroid-per-voxel reduction above.
__global__ void approx_voxel_grid_pcl_kernel(const float4* __restrict__ in, const CloudHeader* __restrict__ in_hdr, float inv, int field, float lo,
                                             float hi, float4* __restrict__ out, CloudHeader* out_hdr) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  constexpr int kHist = 512;
  __shared__ int s_key[kHist][4];   // ix, iy, iz, count
  __shared__ float s_acc[kHist][6]; // x, y, z, r, g, b  (upstream's 7-vector also carries the constant 1 of the padding lane)
  for (int k = 0; k < kHist; ++k) { s_key[k][3] = 0; for (int q = 0; q < 6; ++q) s_acc[k][q] = 0.f; }
  const int n = in_hdr->n;
  int op = 0;
  auto flush = [&](int k) {
    const float cnt = (float)s_key[k][3];
    const unsigned int r = (unsigned int)(int)(s_acc[k][3] / cnt), g = (unsigned int)(int)(s_acc[k][4] / cnt), b = (unsigned int)(int)(s_acc[k][5] / cnt);
    out[op++] = make_float4(s_acc[k][0] / cnt, s_acc[k][1] / cnt, s_acc[k][2] / cnt, __uint_as_float((r << 16) | (g << 8) | b));
  };
  for (int i = 0; i < n; ++i) {
    const float4 p = in[i];
    if (!passes(p, field, lo, hi)) continue;  // the PassThrough that precedes the grid (ref :637), order preserving
    const int ix = (int)floorf(p.x * inv), iy = (int)floorf(p.y * inv), iz = (int)floorf(p.z * inv);
    const int k = (int)((unsigned int)(ix * 7171 + iy * 3079 + iz * 4231) & (unsigned int)(kHist - 1));
    if (s_key[k][3] && (ix != s_key[k][0] || iy != s_key[k][1] || iz != s_key[k][2])) {
      flush(k);
      s_key[k][3] = 0;
      for (int q = 0; q < 6; ++q) s_acc[k][q] = 0.f;
    }
    s_key[k][0] = ix; s_key[k][1] = iy; s_key[k][2] = iz; s_key[k][3]++;
    const unsigned int rgba = __float_as_uint(p.w);
    s_acc[k][0] += p.x; s_acc[k][1] += p.y; s_acc[k][2] += p.z;
    s_acc[k][3] += (float)((rgba >> 16) & 0xffu); s_acc[k][4] += (float)((rgba >> 8) & 0xffu); s_acc[k][5] += (float)(rgba & 0xffu);
  }
  for (int k = 0; k < kHist; ++k) if (s_key[k][3]) flush(k);
  out_hdr->n = op;
}

// ---- centroid (fp64 sums, exact) and in-place translation by -centroid; one block.
__global__ void __launch_bounds__(1024) centre_kernel(float4* pts, const CloudHeader* __restrict__ hdr, float* centroid3) {
  __shared__ double red[32];
  __shared__ float c[3];
  const int n = hdr->n;
  double sx = 0, sy = 0, sz = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { const float4 p = pts[i]; sx += (double)p.x; sy += (double)p.y; sz += (double)p.z; }
  sx = block_sum(sx, red); sy = block_sum(sy, red); sz = block_sum(sz, red);
  if (threadIdx.x == 0) {
    const double dn = n > 0 ? (double)n : 1.0;
    c[0] = (float)(sx / dn); c[1] = (float)(sy / dn); c[2] = (float)(sz / dn);
    centroid3[0] = c[0]; centroid3[1] = c[1]; centroid3[2] = c[2];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) { float4 p = pts[i]; p.x = p.x - c[0]; p.y = p.y - c[1]; p.z = p.z - c[2]; pts[i] = p; }
}

inline int grid_for(size_t n, int block, int sm_count) {
  size_t g = (n + block - 1) / block;
  size_t cap = (size_t)sm_count * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

int launch_unpack_pcl32(cudaStream_t s, const void* src32, float4* dst, size_t n) {
  if (n == 0) return PFT_OK;
  unpack_pcl32_kernel<<<grid_for(n, 256, 148), 256, 0, s>>>(reinterpret_cast<const uint4*>(src32), dst, n);
  PFT_LAUNCH_CHECK();
  return PFT_OK;
}
int launch_unpack_pointcloud2(cudaStream_t s, const void* src, float4* dst, unsigned int width, unsigned int height, unsigned int point_step,
                              unsigned int row_step, int off_x, int off_y, int off_z, int off_rgb) {
  const size_t n = (size_t)width * height;
  if (n == 0) return PFT_OK;
  unpack_pointcloud2_kernel<<<grid_for(n, 256, 148), 256, 0, s>>>(reinterpret_cast<const unsigned int*>(src), dst, width, height, point_step / 4, row_step / 4,
                                                                 off_x / 4, off_y / 4, off_z / 4, off_rgb >= 0 ? off_rgb / 4 : -1);
  PFT_LAUNCH_CHECK();
  return PFT_OK;
}
int launch_pack_pcl32(cudaStream_t s, const float4* src, void* dst32, size_t n) {
  if (n == 0) return PFT_OK;
  pack_pcl32_kernel<<<grid_for(n, 256, 148), 256, 0, s>>>(src, reinterpret_cast<uint4*>(dst32), n);
  PFT_LAUNCH_CHECK();
  return PFT_OK;
}
int launch_set_header(cudaStream_t s, CloudHeader* hdr, int n) {
  set_header_kernel<<<1, 1, 0, s>>>(hdr, n);
  PFT_LAUNCH_CHECK();
  return PFT_OK;
}

int run_passthrough(pft_context* ctx, const pft_cloud* in, pft_cloud* out, int field, float lo, float hi, int drop_zero) {
  int rc = out->ensure(in->capacity);
  if (rc) return rc;
  passthrough_kernel<<<1, 1024, 0, ctx->stream>>>(in->d_pts(), in->d_hdr(), out->d_pts(), out->d_hdr(), field, lo, hi, drop_zero);
  PFT_LAUNCH_CHECK();
  out->host_n = -1;
  return PFT_OK;
}

// The ten stream operations of one downsample (4 memsets + 6 kernels).
static int enqueue_voxel_grid(pft_context* ctx, const pft_cloud* in, pft_cloud* out, float leaf, int field, float lo, float hi, size_t cap, size_t H,
                              int n_tiles) {
  cudaStream_t s = ctx->stream;
  PFT_CUDA_TRY(cudaMemsetAsync(ctx->k1_keys.p, 0xff, H * sizeof(unsigned long long), s));
  PFT_CUDA_TRY(cudaMemsetAsync(ctx->k1_first.p, 0x7f, H * sizeof(int), s));
  PFT_CUDA_TRY(cudaMemsetAsync(ctx->k1_acc_xyz.p, 0, cap * 3 * sizeof(double), s));
  PFT_CUDA_TRY(cudaMemsetAsync(ctx->k1_acc_rgbc.p, 0, cap * 4 * sizeof(unsigned int), s));
  const float inv = 1.0f / leaf;
  const int grid = grid_for(cap, 256, ctx->sm_count);
  k1_insert_kernel<<<grid, 256, 0, s>>>(in->d_pts(), in->d_hdr(), ctx->k1_keys.as<unsigned long long>(), ctx->k1_first.as<int>(),
                                        ctx->k1_slot_of.as<int>(), (unsigned int)(H - 1), inv, field, lo, hi);
  PFT_LAUNCH_CHECK();
  k1_tile_count_kernel<<<n_tiles, 1024, 0, s>>>(in->d_hdr(), ctx->k1_slot_of.as<int>(), ctx->k1_first.as<int>(), ctx->k1_blk.as<int>());
  PFT_LAUNCH_CHECK();
  k1_tile_scan_kernel<<<1, 1024, 0, s>>>(n_tiles, ctx->k1_blk.as<int>(), out->d_hdr());
  PFT_LAUNCH_CHECK();
  k1_tile_assign_kernel<<<n_tiles, 1024, 0, s>>>(in->d_hdr(), ctx->k1_slot_of.as<int>(), ctx->k1_first.as<int>(), ctx->k1_blk.as<int>(),
                                                 ctx->k1_vid.as<int>(), ctx->k1_acc_xyz.as<double>(), ctx->k1_acc_rgbc.as<unsigned int>());
  PFT_LAUNCH_CHECK();
  k1_accum_kernel<<<grid, 256, 0, s>>>(in->d_pts(), in->d_hdr(), ctx->k1_slot_of.as<int>(), ctx->k1_vid.as<int>(),
                                       ctx->k1_acc_xyz.as<double>(), ctx->k1_acc_rgbc.as<unsigned int>());
  PFT_LAUNCH_CHECK();
  k1_final_kernel<<<grid, 256, 0, s>>>(out->d_hdr(), ctx->k1_acc_xyz.as<double>(), ctx->k1_acc_rgbc.as<unsigned int>(), out->d_pts());
  PFT_LAUNCH_CHECK();
  return PFT_OK;
}

int run_voxel_grid(pft_context* ctx, const pft_cloud* in, pft_cloud* out, float leaf, int field, float lo, float hi) {
  if (!(leaf > 0.f)) { set_last_error("leaf size must be positive"); return PFT_ERR_INVALID; }
  const size_t cap = in->capacity;
  int rc = out->ensure(cap);
  if (rc) return rc;
  if (cap == 0) return launch_set_header(ctx->stream, out->d_hdr(), 0);
  size_t H = 1024;
  while (H < 2 * cap) H <<= 1;
  if ((rc = ctx->k1_keys.reserve(H * sizeof(unsigned long long)))) return rc;
  if ((rc = ctx->k1_first.reserve(H * sizeof(int)))) return rc;
  if ((rc = ctx->k1_vid.reserve(H * sizeof(int)))) return rc;
  if ((rc = ctx->k1_slot_of.reserve(cap * sizeof(int)))) return rc;
  if ((rc = ctx->k1_acc_xyz.reserve(cap * 3 * sizeof(double)))) return rc;
  if ((rc = ctx->k1_acc_rgbc.reserve(cap * 4 * sizeof(unsigned int)))) return rc;
  const int n_tiles = (int)((cap + kTile - 1) / kTile);
  if ((rc = ctx->k1_blk.reserve((size_t)n_tiles * sizeof(int)))) return rc;
  out->host_n = -1;
  (void)kFirstInit;
  // A camera stream downsamples frame after frame between the same buffers: the ten operations are captured once
  // into a CUDA graph and replayed (one launch per frame instead of ten).  Everything the sequence touches is in the key.
  static const bool use_graph = [] { const char* e = getenv("PFT_NO_GRAPH"); return !(e && e[0] == '1'); }();
  if (!use_graph) return enqueue_voxel_grid(ctx, in, out, leaf, field, lo, hi, cap, H, n_tiles);
  pft_context::K1Key key;
  memset(&key, 0, sizeof(key));
  key.p[0] = in->d_pts(); key.p[1] = in->d_hdr(); key.p[2] = out->d_pts(); key.p[3] = out->d_hdr();
  key.p[4] = ctx->k1_keys.p; key.p[5] = ctx->k1_first.p; key.p[6] = ctx->k1_vid.p; key.p[7] = ctx->k1_slot_of.p;
  key.p[8] = ctx->k1_acc_xyz.p; key.p[9] = ctx->k1_acc_rgbc.p; key.p[10] = ctx->k1_blk.p;
  key.cap = cap; key.H = H; key.leaf = leaf; key.lo = lo; key.hi = hi; key.field = field;
  // a few alternating (in, out) pairs are common (double-buffered frames): keep a small set of graphs
  int slot = -1;
  for (int k = 0; k < pft_context::kK1Graphs; ++k)
    if (ctx->k1_exec[k] && memcmp(&ctx->k1_key[k], &key, sizeof(key)) == 0) { slot = k; break; }
  if (slot < 0) {
    slot = ctx->k1_next;
    ctx->k1_next = (ctx->k1_next + 1) % pft_context::kK1Graphs;
    if (ctx->k1_exec[slot]) { cudaGraphExecDestroy(ctx->k1_exec[slot]); ctx->k1_exec[slot] = nullptr; }
    cudaGraph_t graph = nullptr;
    const unsigned long long before = g_launch_count.load();
    PFT_CUDA_TRY(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    rc = enqueue_voxel_grid(ctx, in, out, leaf, field, lo, hi, cap, H, n_tiles);
    cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
    g_launch_count -= g_launch_count.load() - before;  // captured, not launched
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) { set_last_error("downsample graph capture failed: %s", cudaGetErrorString(ce)); cudaGetLastError(); return PFT_ERR_CUDA; }
    cudaError_t ie = cudaGraphInstantiate(&ctx->k1_exec[slot], graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { set_last_error("downsample graph instantiate failed: %s", cudaGetErrorString(ie)); cudaGetLastError(); ctx->k1_exec[slot] = nullptr; return PFT_ERR_CUDA; }
    ctx->k1_key[slot] = key;
  }
  PFT_CUDA_TRY(cudaGraphLaunch(ctx->k1_exec[slot], ctx->stream));
  g_launch_count += 6;  // the six kernels of the sequence
  return PFT_OK;
}

int run_approx_voxel_grid_pcl(pft_context* ctx, const pft_cloud* in, pft_cloud* out, float leaf, int field, float lo, float hi) {
  if (!(leaf > 0.f)) { set_last_error("leaf size must be positive"); return PFT_ERR_INVALID; }
  int rc = out->ensure(in->capacity);
  if (rc) return rc;
  approx_voxel_grid_pcl_kernel<<<1, 32, 0, ctx->stream>>>(in->d_pts(), in->d_hdr(), 1.0f / leaf, field, lo, hi, out->d_pts(), out->d_hdr());
  PFT_LAUNCH_CHECK();
  out->host_n = -1;
  return PFT_OK;
}

int run_centre_on_centroid(pft_context* ctx, pft_cloud* cloud, float* d_centroid3) {
  centre_kernel<<<1, 1024, 0, ctx->stream>>>(cloud->d_pts(), cloud->d_hdr(), d_centroid3);
  PFT_LAUNCH_CHECK();
  return PFT_OK;
}

}  // namespace pft
