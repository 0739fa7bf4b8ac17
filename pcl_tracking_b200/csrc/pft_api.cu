// pft_api.cu -- C ABI: context, device clouds and the filter entry points (include/pft/pft.h).
//
// Host-side plumbing only; the kernels live in pft_filters.cu / pft_tracker_kernels.cuh.  Everything
// here replaces glue the reference does on the CPU around PCL containers:
//   pft_cloud_upload        pcl::fromPCLPointCloud2 into PointCloud<PointXYZRGBA>  (ref: src/auto_tracking.cpp:619-622)
//   pft_passthrough         filterPassThrough                                       (ref: src/auto_tracking.cpp:536-547)
//   pft_passthrough_voxel_grid  filterPassThrough + gridSampleApprox / gridSample   (ref: src/auto_tracking.cpp:637, :683, :641)
//   pft_prepare_model       removeZeroPoints + centroid + translate + gridSample    (ref: src/auto_tracking.cpp:656-674)
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "pft_internal.h"

namespace pft {

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launch_count{0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int DevBuf::reserve(size_t need) {
  if (need <= bytes) return PFT_OK;
  // grow geometrically so that a slowly growing frame does not reallocate every call
  size_t want = need + need / 4 + 256;
  void* np = nullptr;
  PFT_CUDA_TRY(cudaMalloc(&np, want));
  if (p) {
    // callers never rely on the old contents across a reserve() except clouds, which copy explicitly
    cudaFree(p);
  }
  p = np;
  bytes = want;
  return PFT_OK;
}
void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  bytes = 0;
}

}  // namespace pft

int pft_cloud::ensure(size_t cap) {
  int rc = hdr.reserve(sizeof(pft::CloudHeader));
  if (rc) return rc;
  if (peer_exported && cap * sizeof(float4) > pts.bytes) {
    pft::set_last_error("the cloud is mapped by its peers with room for %zu points; %zu do not fit (export it with a larger capacity)", peer_capacity, cap);
    return PFT_ERR_CAPACITY;
  }
  if (cap > capacity || !pts.p) {
    rc = pts.reserve((cap ? cap : 1) * sizeof(float4));
    if (rc) return rc;
  }
  capacity = cap;
  return PFT_OK;
}

// Orders the context stream after an asynchronous upload into this cloud (no host wait).
int pft_cloud::join_upload() const {
  if (!upload_pending) return PFT_OK;
  PFT_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ready, 0));
  upload_pending = false;
  return PFT_OK;
}

using namespace pft;

namespace {

// Common front of the asynchronous uploads: the copy stream, ordered after everything the context stream holds so far
// (earlier readers of the cloud that is about to be overwritten).
int begin_async_upload(pft_cloud* c, size_t n, cudaStream_t* cs) {
  pft_context* ctx = c->ctx;
  PFT_CUDA_TRY(cudaSetDevice(ctx->device));
  if (!ctx->copy_stream) {
    PFT_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    PFT_CUDA_TRY(cudaEventCreateWithFlags(&ctx->copy_fence, cudaEventDisableTiming));
  }
  if (!c->ready) PFT_CUDA_TRY(cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming));
  int rc = c->ensure(n);
  if (rc) return rc;
  PFT_CUDA_TRY(cudaEventRecord(ctx->copy_fence, ctx->stream));
  PFT_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_fence, 0));
  *cs = ctx->copy_stream;
  return PFT_OK;
}

int end_async_upload(pft_cloud* c, size_t n, cudaStream_t cs) {
  int rc = launch_set_header(cs, c->d_hdr(), (int)n);
  if (rc) return rc;
  PFT_CUDA_TRY(cudaEventRecord(c->ready, cs));
  c->upload_pending = true;
  c->host_n = (long long)n;
  return PFT_OK;
}

int check_pointcloud2_layout(size_t n, uint32_t width, uint32_t point_step, uint32_t row_step, int32_t off_x, int32_t off_y, int32_t off_z, int32_t off_rgb,
                             int is_bigendian) {
  if (n > 0x7fffffffull) { set_last_error("cloud too large"); return PFT_ERR_INVALID; }
  if (is_bigendian) { set_last_error("big-endian PointCloud2 data is not supported"); return PFT_ERR_INVALID; }
  if (n) {
    if (point_step < 12 || (point_step & 3) || (row_step & 3) || (size_t)row_step < (size_t)width * point_step) {
      set_last_error("bad PointCloud2 strides: point_step %u, row_step %u, width %u (4-byte multiples, row_step >= width * point_step)", point_step, row_step, width);
      return PFT_ERR_INVALID;
    }
    const int32_t offs[4] = {off_x, off_y, off_z, off_rgb};
    for (int k = 0; k < 4; ++k) {
      const bool optional = k == 3 && offs[k] < 0;  // no colour field: rgba = 0
      if (!optional && (offs[k] < 0 || (offs[k] & 3) || (uint32_t)offs[k] + 4 > point_step)) {
        set_last_error("bad PointCloud2 field offset %d (float32 x, y, z and the packed rgb(a) word, 4-byte aligned inside point_step %u)", offs[k], point_step);
        return PFT_ERR_INVALID;
      }
    }
  }
  return PFT_OK;
}

}  // namespace

extern "C" {

const char* pft_last_error(void) { return g_err; }
const char* pft_version(void) { return "pft 0.1 (sm_100a)"; }

int pft_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int pft_context_create(int device, pft_context** out) {
  if (!out) { set_last_error("pft_context_create: null out"); return PFT_ERR_INVALID; }
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_last_error("no CUDA device available (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
    return PFT_ERR_CUDA;
  }
  if (device < 0 || device >= n) { set_last_error("device %d out of range [0,%d)", device, n); return PFT_ERR_INVALID; }
  PFT_CUDA_TRY(cudaSetDevice(device));
  pft_context* c = new pft_context();
  c->device = device;
  cudaDeviceProp prop;
  PFT_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  PFT_CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->pinned_bytes = 4096;
  PFT_CUDA_TRY(cudaHostAlloc(&c->pinned, c->pinned_bytes, cudaHostAllocDefault));
  *out = c;
  return PFT_OK;
}

void pft_context_destroy(pft_context* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  pft_context_comm_destroy(c);
  pft::DevBuf* bufs[] = {&c->k1_keys, &c->k1_first, &c->k1_vid, &c->k1_slot_of, &c->k1_acc_xyz, &c->k1_acc_rgbc, &c->k1_blk, &c->staging, &c->tmp_cloud_pts, &c->tmp_hdr, &c->tmp_f,
                         &c->cl_grid, &c->cl_cells, &c->cl_work, &c->cl_sel};
  for (auto* b : bufs) b->release();
  if (c->pinned) cudaFreeHost(c->pinned);
  if (c->batch_fork) cudaEventDestroy(c->batch_fork);
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  if (c->copy_fence) cudaEventDestroy(c->copy_fence);
  for (int k = 0; k < pft_context::kK1Graphs; ++k) if (c->k1_exec[k]) cudaGraphExecDestroy(c->k1_exec[k]);
  cudaStreamDestroy(c->stream);
  delete c;
}

int pft_context_synchronize(pft_context* c) {
  if (!c) { set_last_error("null context"); return PFT_ERR_INVALID; }
  PFT_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return PFT_OK;
}

void* pft_context_stream(pft_context* c) { return c ? (void*)c->stream : nullptr; }

int pft_host_alloc(void** ptr, size_t bytes) {
  if (!ptr) { set_last_error("null ptr"); return PFT_ERR_INVALID; }
  PFT_CUDA_TRY(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return PFT_OK;
}
int pft_host_free(void* ptr) {
  if (ptr) PFT_CUDA_TRY(cudaFreeHost(ptr));
  return PFT_OK;
}

uint64_t pft_kernel_launch_count(void) { return (uint64_t)g_launch_count.load(); }

// ------------------------------------------------------------------ clouds
int pft_cloud_create(pft_context* ctx, pft_cloud** out) {
  if (!ctx || !out) { set_last_error("pft_cloud_create: null argument"); return PFT_ERR_INVALID; }
  PFT_CUDA_TRY(cudaSetDevice(ctx->device));
  pft_cloud* c = new pft_cloud();
  c->ctx = ctx;
  int rc = c->ensure(0);
  if (rc) { delete c; return rc; }
  rc = launch_set_header(ctx->stream, c->d_hdr(), 0);
  if (rc) { delete c; return rc; }
  c->host_n = 0;
  *out = c;
  return PFT_OK;
}

void pft_cloud_destroy(pft_cloud* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  cudaStreamSynchronize(c->ctx->stream);
  if (c->ready) { cudaEventSynchronize(c->ready); cudaEventDestroy(c->ready); }
  pft_cloud_peer_detach(c);
  c->pts.release();
  c->hdr.release();
  c->raw_staging.release();
  delete c;
}

int pft_cloud_upload(pft_cloud* c, const void* host_points, size_t n, int layout) {
  if (!c || (n && !host_points)) { set_last_error("pft_cloud_upload: null argument"); return PFT_ERR_INVALID; }
  if (layout != PFT_LAYOUT_PACKED16 && layout != PFT_LAYOUT_PCL32) { set_last_error("unknown layout %d", layout); return PFT_ERR_INVALID; }
  if (n > 0x7fffffffull) { set_last_error("cloud too large"); return PFT_ERR_INVALID; }
  pft_context* ctx = c->ctx;
  PFT_CUDA_TRY(cudaSetDevice(ctx->device));
  int rc = c->join_upload();
  if (rc) return rc;
  if ((rc = c->ensure(n))) return rc;
  cudaStream_t s = ctx->stream;
  if (n) {
    if (layout == PFT_LAYOUT_PACKED16) {
      PFT_CUDA_TRY(cudaMemcpyAsync(c->d_pts(), host_points, n * sizeof(float4), cudaMemcpyHostToDevice, s));
    } else {
      // 32-byte pcl::PointXYZRGBA records: copy raw, repack to float4 {x,y,z,rgba} on the device
      if ((rc = ctx->staging.reserve(n * 32))) return rc;
      PFT_CUDA_TRY(cudaMemcpyAsync(ctx->staging.p, host_points, n * 32, cudaMemcpyHostToDevice, s));
      if ((rc = launch_unpack_pcl32(s, ctx->staging.p, c->d_pts(), n))) return rc;
    }
  }
  if ((rc = launch_set_header(s, c->d_hdr(), (int)n))) return rc;
  c->host_n = (long long)n;
  return PFT_OK;
}

int pft_cloud_upload_pointcloud2(pft_cloud* c, const void* data, uint32_t width, uint32_t height, uint32_t point_step, uint32_t row_step,
                                 int32_t off_x, int32_t off_y, int32_t off_z, int32_t off_rgb, int is_bigendian) {
  if (!c) { set_last_error("pft_cloud_upload_pointcloud2: null cloud"); return PFT_ERR_INVALID; }
  const size_t n = (size_t)width * height;
  if (n && !data) { set_last_error("pft_cloud_upload_pointcloud2: null data"); return PFT_ERR_INVALID; }
  int rc = check_pointcloud2_layout(n, width, point_step, row_step, off_x, off_y, off_z, off_rgb, is_bigendian);
  if (rc) return rc;
  pft_context* ctx = c->ctx;
  PFT_CUDA_TRY(cudaSetDevice(ctx->device));
  if ((rc = c->join_upload())) return rc;
  if ((rc = c->ensure(n))) return rc;
  cudaStream_t s = ctx->stream;
  if (n) {
    const size_t bytes = (size_t)row_step * height;
    if ((rc = ctx->staging.reserve(bytes))) return rc;
    PFT_CUDA_TRY(cudaMemcpyAsync(ctx->staging.p, data, bytes, cudaMemcpyHostToDevice, s));
    if ((rc = launch_unpack_pointcloud2(s, ctx->staging.p, c->d_pts(), width, height, point_step, row_step, off_x, off_y, off_z, off_rgb))) return rc;
  }
  if ((rc = launch_set_header(s, c->d_hdr(), (int)n))) return rc;
  c->host_n = (long long)n;
  return PFT_OK;
}

int pft_cloud_upload_async(pft_cloud* c, const void* host_points, size_t n, int layout) {
  if (!c || (n && !host_points)) { set_last_error("pft_cloud_upload_async: null argument"); return PFT_ERR_INVALID; }
  if (layout != PFT_LAYOUT_PACKED16 && layout != PFT_LAYOUT_PCL32) { set_last_error("unknown layout %d", layout); return PFT_ERR_INVALID; }
  if (n > 0x7fffffffull) { set_last_error("cloud too large"); return PFT_ERR_INVALID; }
  cudaStream_t cs = nullptr;
  int rc = begin_async_upload(c, n, &cs);
  if (rc) return rc;
  if (n) {
    if (layout == PFT_LAYOUT_PACKED16) {
      PFT_CUDA_TRY(cudaMemcpyAsync(c->d_pts(), host_points, n * sizeof(float4), cudaMemcpyHostToDevice, cs));
    } else {
      if ((rc = c->raw_staging.reserve(n * 32))) return rc;
      PFT_CUDA_TRY(cudaMemcpyAsync(c->raw_staging.p, host_points, n * 32, cudaMemcpyHostToDevice, cs));
      if ((rc = launch_unpack_pcl32(cs, c->raw_staging.p, c->d_pts(), n))) return rc;
    }
  }
  return end_async_upload(c, n, cs);
}

int pft_cloud_upload_pointcloud2_async(pft_cloud* c, const void* data, uint32_t width, uint32_t height, uint32_t point_step, uint32_t row_step,
                                       int32_t off_x, int32_t off_y, int32_t off_z, int32_t off_rgb, int is_bigendian) {
  if (!c) { set_last_error("pft_cloud_upload_pointcloud2_async: null cloud"); return PFT_ERR_INVALID; }
  const size_t n = (size_t)width * height;
  if (n && !data) { set_last_error("pft_cloud_upload_pointcloud2_async: null data"); return PFT_ERR_INVALID; }
  int rc = check_pointcloud2_layout(n, width, point_step, row_step, off_x, off_y, off_z, off_rgb, is_bigendian);
  if (rc) return rc;
  cudaStream_t cs = nullptr;
  if ((rc = begin_async_upload(c, n, &cs))) return rc;
  if (n) {
    const size_t bytes = (size_t)row_step * height;
    if ((rc = c->raw_staging.reserve(bytes))) return rc;
    PFT_CUDA_TRY(cudaMemcpyAsync(c->raw_staging.p, data, bytes, cudaMemcpyHostToDevice, cs));
    if ((rc = launch_unpack_pointcloud2(cs, c->raw_staging.p, c->d_pts(), width, height, point_step, row_step, off_x, off_y, off_z, off_rgb))) return rc;
  }
  return end_async_upload(c, n, cs);
}

int pft_cloud_wait_upload(pft_cloud* c) {
  if (!c) { set_last_error("pft_cloud_wait_upload: null cloud"); return PFT_ERR_INVALID; }
  PFT_CUDA_TRY(cudaSetDevice(c->ctx->device));
  return c->join_upload();
}

int pft_cloud_size(pft_cloud* c, size_t* n) {
  if (!c || !n) { set_last_error("pft_cloud_size: null argument"); return PFT_ERR_INVALID; }
  { int jrc = c->join_upload(); if (jrc) return jrc; }  // every consumer that sizes its work from the cloud passes here
  if (c->host_n < 0) {
    pft_context* ctx = c->ctx;
    PFT_CUDA_TRY(cudaSetDevice(ctx->device));
    PFT_CUDA_TRY(cudaMemcpyAsync(ctx->pinned, c->d_hdr(), sizeof(CloudHeader), cudaMemcpyDeviceToHost, ctx->stream));
    PFT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    c->host_n = reinterpret_cast<CloudHeader*>(ctx->pinned)->n;
  }
  *n = (size_t)c->host_n;
  return PFT_OK;
}

int pft_cloud_download(pft_cloud* c, void* host_points, size_t capacity, int layout, size_t* n_out) {
  if (!c) { set_last_error("pft_cloud_download: null cloud"); return PFT_ERR_INVALID; }
  if (layout != PFT_LAYOUT_PACKED16 && layout != PFT_LAYOUT_PCL32) { set_last_error("unknown layout %d", layout); return PFT_ERR_INVALID; }
  size_t n = 0;
  int rc = pft_cloud_size(c, &n);
  if (rc) return rc;
  if (n_out) *n_out = n;
  if (n > capacity) { set_last_error("pft_cloud_download: %zu points, capacity %zu", n, capacity); return PFT_ERR_CAPACITY; }
  if (n == 0) return PFT_OK;
  if (!host_points) { set_last_error("pft_cloud_download: null buffer"); return PFT_ERR_INVALID; }
  pft_context* ctx = c->ctx;
  cudaStream_t s = ctx->stream;
  if ((rc = c->join_upload())) return rc;
  if (layout == PFT_LAYOUT_PACKED16) {
    PFT_CUDA_TRY(cudaMemcpyAsync(host_points, c->d_pts(), n * sizeof(float4), cudaMemcpyDeviceToHost, s));
  } else {
    if ((rc = ctx->staging.reserve(n * 32))) return rc;
    if ((rc = launch_pack_pcl32(s, c->d_pts(), ctx->staging.p, n))) return rc;
    PFT_CUDA_TRY(cudaMemcpyAsync(host_points, ctx->staging.p, n * 32, cudaMemcpyDeviceToHost, s));
  }
  PFT_CUDA_TRY(cudaStreamSynchronize(s));
  return PFT_OK;
}

// ------------------------------------------------------------------ filters
static int check_filter_args(pft_context* ctx, const pft_cloud* in, pft_cloud* out, const char* who) {
  if (!ctx || !in || !out) { set_last_error("%s: null argument", who); return PFT_ERR_INVALID; }
  if (in == out) { set_last_error("%s: in-place filtering is not supported", who); return PFT_ERR_INVALID; }
  if (in->ctx != ctx || out->ctx != ctx) { set_last_error("%s: clouds belong to another context", who); return PFT_ERR_INVALID; }
  PFT_CUDA_TRY(cudaSetDevice(ctx->device));
  int rc = in->join_upload();  // an asynchronous upload into the input / output must land before the filter touches it
  if (rc) return rc;
  return out->join_upload();
}

int pft_passthrough(pft_context* ctx, const pft_cloud* in, pft_cloud* out, int field, float lo, float hi) {
  int rc = check_filter_args(ctx, in, out, "pft_passthrough");
  if (rc) return rc;
  if (field < 0 || field > 2) { set_last_error("pft_passthrough: field must be 0 (x), 1 (y) or 2 (z)"); return PFT_ERR_INVALID; }
  return run_passthrough(ctx, in, out, field, lo, hi, 0);
}

int pft_passthrough_voxel_grid(pft_context* ctx, const pft_cloud* in, pft_cloud* out, float leaf, int field, float lo, float hi) {
  int rc = check_filter_args(ctx, in, out, "pft_passthrough_voxel_grid");
  if (rc) return rc;
  if (field > 2) { set_last_error("pft_passthrough_voxel_grid: field must be <0 (off), 0, 1 or 2"); return PFT_ERR_INVALID; }
  return run_voxel_grid(ctx, in, out, leaf, field, lo, hi);
}

int pft_approx_voxel_grid_pcl(pft_context* ctx, const pft_cloud* in, pft_cloud* out, float leaf, int field, float lo, float hi) {
  int rc = check_filter_args(ctx, in, out, "pft_approx_voxel_grid_pcl");
  if (rc) return rc;
  if (field > 2) { set_last_error("pft_approx_voxel_grid_pcl: field must be <0 (off), 0, 1 or 2"); return PFT_ERR_INVALID; }
  return run_approx_voxel_grid_pcl(ctx, in, out, leaf, field, lo, hi);
}

int pft_prepare_model(pft_context* ctx, const pft_cloud* raw, pft_cloud* out, float leaf, float* centroid3) {
  int rc = check_filter_args(ctx, raw, out, "pft_prepare_model");
  if (rc) return rc;
  // removeZeroPoints -> centroid -> translate by -centroid -> VoxelGrid(leaf)
  pft_cloud tmp;
  tmp.ctx = ctx;
  // scratch cloud backed by context buffers (kept across calls)
  tmp.pts = ctx->tmp_cloud_pts;
  tmp.hdr = ctx->tmp_hdr;
  tmp.capacity = tmp.pts.bytes / sizeof(float4);
  auto stash = [&]() { ctx->tmp_cloud_pts = tmp.pts; ctx->tmp_hdr = tmp.hdr; tmp.pts = DevBuf(); tmp.hdr = DevBuf(); };
  if ((rc = tmp.ensure(raw->capacity))) { stash(); return rc; }
  if ((rc = ctx->tmp_f.reserve(16 * sizeof(float)))) { stash(); return rc; }
  if ((rc = run_passthrough(ctx, raw, &tmp, -1, 0.f, 0.f, 1))) { stash(); return rc; }
  if ((rc = run_centre_on_centroid(ctx, &tmp, ctx->tmp_f.as<float>()))) { stash(); return rc; }
  rc = run_voxel_grid(ctx, &tmp, out, leaf, -1, 0.f, 0.f);
  stash();
  if (rc) return rc;
  if (centroid3) {
    PFT_CUDA_TRY(cudaMemcpyAsync(ctx->pinned, ctx->tmp_f.p, 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PFT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    memcpy(centroid3, ctx->pinned, 3 * sizeof(float));
  }
  return PFT_OK;
}

}  // extern "C"
