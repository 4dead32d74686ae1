// LiDAR/radar hard voxelizer, VFE mean and pillar scatter (sm_100a).
//
// Reference contract being replaced (third-party, reached through models/bev_depth.py:181-183 of
// the reference; semantics restated in SURVEY.md Appendix A from mmcv-full 1.7.0 /
// mmdet3d 1.0.0rc4): `hard_voxelize` is *serial in point order* -- a voxel is created at the
// first in-range point of its cell, at most max_voxels voxels exist (later cells are dropped
// entirely), and the first max_points points of a voxel are kept in order.
//
// Parallel formulation with bit-identical results:
//   1. warp-cooperative hash insert of (sample, cell) keys; per slot atomicMin of the point
//      index  -> first-occurrence point of every cell (order independent)
//   2. flag "point is the first of its cell" -> exclusive scan in point order = voxel id
//      (mmcv's voxel numbering), capped at max_voxels
//   3. per voxel, the max_points smallest point indices via a chain of atomicMin on a sorted
//      slot list (each slot ends up holding the r-th smallest index whatever the interleaving)
//   4. gather: every element of voxels / num_points is written exactly once, zero padding
//      included; optional fused VFE mean (HardSimpleVFE)
// mmcv's own deterministic CUDA path needs an O(N^2) predecessor scan and a single-thread
// numbering kernel for the same result.
#include "common.cuh"
#include "scan.cuh"

namespace bevpool {

constexpr unsigned long long kEmptyKey = ~0ull;
constexpr int32_t kEmptyIdx = 0x7f7f7f7f;   // what cudaMemset(0x7f) writes

struct VoxGeom {
  float vx, vy, vz, xmin, ymin, zmin;
  int gx, gy, gz;
};

// IEEE float32 subtract, true division, floor -- exactly mmcv's `floor((p - min) / voxel_size)`.
// NaN and out-of-int-range values compare false / out of range and are dropped.
__device__ __forceinline__ bool point_to_cell(const float *p, const VoxGeom &g, int &x, int &y, int &z) {
  const float fx = floorf(__fdiv_rn(__fsub_rn(p[0], g.xmin), g.vx));
  const float fy = floorf(__fdiv_rn(__fsub_rn(p[1], g.ymin), g.vy));
  const float fz = floorf(__fdiv_rn(__fsub_rn(p[2], g.zmin), g.vz));
  const bool ok = fx >= 0.f && fx < (float)g.gx && fy >= 0.f && fy < (float)g.gy && fz >= 0.f && fz < (float)g.gz;
  x = ok ? (int)fx : -1;
  y = ok ? (int)fy : -1;
  z = ok ? (int)fz : -1;
  return ok;
}

__device__ __forceinline__ uint32_t hash_u64(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  return (uint32_t)k;
}

// ---- 1. hash insert --------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vox_insert_kernel(const float *__restrict__ points, const int32_t *__restrict__ offsets, int F, VoxGeom g,
                  unsigned long long *__restrict__ keys, int32_t *__restrict__ first, uint32_t hmask,
                  int32_t *__restrict__ point_slot) {
  const int b = blockIdx.y;
  const int begin = offsets[b], end = offsets[b + 1];
  const int i = begin + blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = i < end;
  int x, y, z;
  unsigned long long key = kEmptyKey;
  if (in && point_to_cell(points + (int64_t)i * F, g, x, y, z))
    key = (unsigned long long)b * ((unsigned long long)g.gx * g.gy * g.gz) +
          ((unsigned long long)z * g.gy + y) * g.gx + x;
  // one probe sequence per distinct key in the warp; the lowest lane holds the lowest index
  const unsigned peers = __match_any_sync(0xffffffffu, key);
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(peers) - 1;
  int slot = -1;
  if (key != kEmptyKey && lane == leader) {
    uint32_t s = hash_u64(key) & hmask;
    while (true) {
      const unsigned long long prev = atomicCAS(keys + s, kEmptyKey, key);
      if (prev == kEmptyKey || prev == key) break;
      s = (s + 1) & hmask;
    }
    atomicMin(first + s, i);
    slot = (int)s;
  }
  slot = __shfl_sync(0xffffffffu, slot, leader);
  if (in) point_slot[i] = key != kEmptyKey ? slot : -1;
}

// ---- 2. first-occurrence flags (scanned afterwards) -----------------------------------------------
__global__ void vox_flag_kernel(const int32_t *__restrict__ point_slot, const int32_t *__restrict__ first,
                                int64_t n, uint32_t *__restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  uint32_t f = 0;
  if (i < n) {
    const int s = point_slot[i];
    f = (s >= 0 && first[s] == (int32_t)i) ? 1u : 0u;
  }
  flags[i] = f;   // element n stays 0: the scan leaves the grand total there
}

// per-sample voxel counts (capped) and output row bases; B is small, one thread is enough
__global__ void vox_base_kernel(const uint32_t *__restrict__ prefix, const int32_t *__restrict__ offsets,
                                int batch, int max_voxels, int32_t *__restrict__ voxel_base) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int run = 0;
    for (int b = 0; b < batch; ++b) {
      voxel_base[b] = run;
      const int distinct = (int)(prefix[offsets[b + 1]] - prefix[offsets[b]]);
      run += distinct < max_voxels ? distinct : max_voxels;
    }
    voxel_base[batch] = run;
  }
}

// ---- 3. assign points to voxel slots (first max_points indices, ascending) ---------------------------
__global__ void __launch_bounds__(256)
vox_assign_kernel(const float *__restrict__ points, const int32_t *__restrict__ offsets, int F, VoxGeom g,
                  const int32_t *__restrict__ point_slot, const int32_t *__restrict__ first,
                  const uint32_t *__restrict__ prefix, const int32_t *__restrict__ voxel_base, int max_voxels,
                  int max_points, int32_t *__restrict__ lists, int32_t *__restrict__ counts,
                  int32_t *__restrict__ coors) {
  const int b = blockIdx.y;
  const int begin = offsets[b], end = offsets[b + 1];
  const int i = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= end) return;
  const int s = point_slot[i];
  if (s < 0) return;
  const int f = first[s];
  const int vid = (int)(prefix[f] - prefix[begin]);
  if (vid >= max_voxels) return;                       // voxel cap: the whole cell is dropped
  const int row = voxel_base[b] + vid;
  if (i == f) {                                        // the creating point records the coordinates
    int x, y, z;
    point_to_cell(points + (int64_t)i * F, g, x, y, z);
    reinterpret_cast<int4 *>(coors)[row] = make_int4(b, z, y, x);
  }
  atomicAdd(counts + row, 1);
  int32_t *list = lists + (int64_t)row * max_points;
  // hint: once the last slot holds a smaller index this point can never enter the list
  if (*reinterpret_cast<volatile int32_t *>(list + max_points - 1) < i) return;
  int v = i;
  for (int r = 0; r < max_points; ++r) {
    const int old = atomicMin(list + r, v);
    if (old == kEmptyIdx) break;
    v = old > v ? old : v;                              // carry the larger one to the next slot
  }
}

// ---- 4. gather ---------------------------------------------------------------------------------
// one warp per voxel row: copies max_points x F floats (zeros for unused slots), writes num_points
// and, optionally, the mean of the first `mean_features` columns (HardSimpleVFE).
__global__ void __launch_bounds__(256)
vox_gather_kernel(const float *__restrict__ points, int F, const int32_t *__restrict__ lists,
                  const int32_t *__restrict__ counts, const int32_t *__restrict__ voxel_base, int batch,
                  int max_points, float *__restrict__ voxels, int32_t *__restrict__ num_points,
                  float *__restrict__ voxel_mean, int mean_features) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= voxel_base[batch]) return;
  const int32_t *list = lists + (int64_t)row * max_points;
  const int cnt = min(counts[row], max_points);
  float *vout = voxels + (int64_t)row * max_points * F;
  const int total = max_points * F;
  for (int e = lane; e < total; e += 32) {
    const int t = e / F, c = e - t * F;
    float v = 0.f;
    if (t < cnt) v = __ldg(points + (int64_t)list[t] * F + c);
    vout[e] = v;
  }
  if (lane == 0) num_points[row] = cnt;
  if (voxel_mean && lane < mean_features) {
    float s = 0.f;
    for (int t = 0; t < cnt; ++t) s += __ldg(points + (int64_t)list[t] * F + lane);   // sequential, slot order
    voxel_mean[(int64_t)row * mean_features + lane] = s / (float)cnt;
  }
}

// ---- dynamic voxelization: per-point (z, y, x) or -1 -----------------------------------------------
__global__ void vox_dynamic_kernel(const float *__restrict__ points, int64_t n, int F, VoxGeom g,
                                   int32_t *__restrict__ coors) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int x, y, z;
  point_to_cell(points + i * F, g, x, y, z);
  coors[i * 3 + 0] = z;
  coors[i * 3 + 1] = y;
  coors[i * 3 + 2] = x;
}

// ---- pillar scatter --------------------------------------------------------------------------------
__global__ void scatter_index_kernel(const int32_t *__restrict__ coors, int64_t M, int batch, int nz, int ny,
                                     int nx, int32_t *__restrict__ index_map) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int4 c = reinterpret_cast<const int4 *>(coors)[m];   // (b, z, y, x)
  if (c.x < 0 || c.x >= batch || c.y < 0 || c.y >= nz || c.z < 0 || c.z >= ny || c.w < 0 || c.w >= nx) return;
  // duplicates cannot come from hard voxelization; if a caller has them, the highest row wins (deterministic)
  atomicMax(index_map + (((int64_t)c.x * nz + c.y) * ny + c.z) * nx + c.w, (int)m);
}

// canvas (B, C, nz, ny, nx): every element written exactly once (feature or zero), coalesced along x
template <typename T>
__global__ void __launch_bounds__(256)
scatter_canvas_kernel(const T *__restrict__ feats, const int32_t *__restrict__ index_map, int C,
                      int64_t cells_per_sample, T *__restrict__ canvas) {
  const int b = blockIdx.y;
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= cells_per_sample) return;
  const int m = index_map[(int64_t)b * cells_per_sample + cell];
  T *out = canvas + (int64_t)b * C * cells_per_sample + cell;
  if (m < 0) {
    const T z = T(0.f);
    for (int c = 0; c < C; ++c) out[(int64_t)c * cells_per_sample] = z;
  } else {
    const T *row = feats + (int64_t)m * C;
    for (int c = 0; c < C; ++c) out[(int64_t)c * cells_per_sample] = row[c];
  }
}

template <typename T>
__global__ void scatter_backward_kernel(const T *__restrict__ grad_canvas, const int32_t *__restrict__ coors,
                                        int64_t M, int C, int batch, int nz, int ny, int nx,
                                        T *__restrict__ grad_feats) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M * C) return;
  const int64_t m = e / C;
  const int c = (int)(e - m * C);
  const int4 co = reinterpret_cast<const int4 *>(coors)[m];
  T v = T(0.f);
  if (co.x >= 0 && co.x < batch && co.y >= 0 && co.y < nz && co.z >= 0 && co.z < ny && co.w >= 0 && co.w < nx) {
    const int64_t cells = (int64_t)nz * ny * nx;
    v = grad_canvas[((int64_t)co.x * C + c) * cells + ((int64_t)co.y * ny + co.z) * nx + co.w];
  }
  grad_feats[e] = v;
}

struct VoxTemp {
  size_t off_scan, off_counts, zero_bytes;   // [0, zero_bytes) memset 0
  size_t off_keys;                           // memset 0xff
  size_t off_first, off_lists, fill7f_bytes; // [off_first, off_first + fill7f_bytes) memset 0x7f
  size_t off_slot, off_flags, bytes;
  uint32_t hash_slots;
};

static VoxTemp vox_temp_layout(int batch, int64_t total_points, int max_voxels, int max_points) {
  VoxTemp L{};
  uint32_t h = 1024;
  while ((double)h < 1.25 * (double)total_points) h <<= 1;
  L.hash_slots = h;
  const size_t rows = (size_t)batch * max_voxels;
  size_t o = 0;
  L.off_scan = o;   o += scan_workspace_bytes(total_points + 1);
  L.off_counts = o; o = align_up(o + rows * 4, 256);
  L.zero_bytes = o;
  L.off_keys = o;   o = align_up(o + (size_t)h * 8, 256);
  L.off_first = o;  o = align_up(o + (size_t)h * 4, 256);
  L.off_lists = o;  o = align_up(o + rows * max_points * 4, 256);
  L.fill7f_bytes = o - L.off_first;
  L.off_slot = o;   o = align_up(o + (size_t)total_points * 4, 256);
  L.off_flags = o;  o = align_up(o + (size_t)(total_points + 1) * 4, 256);
  L.bytes = o;
  return L;
}

static int check_vox_args(int batch, int64_t total_points, int F, int max_voxels, int max_points) {
  if (batch <= 0 || batch > 65535 || total_points < 0 || F < 3 || max_voxels <= 0 || max_points <= 0) return BEVPOOL_E_ARG;
  if (total_points >= 0x7f7f7f7f || (int64_t)batch * max_voxels * max_points >= INT32_MAX) return BEVPOOL_E_RANGE;
  return BEVPOOL_OK;
}

static VoxGeom make_geom(const float *vs, const float *range, const int *grid) {
  VoxGeom g;
  g.vx = vs[0]; g.vy = vs[1]; g.vz = vs[2];
  g.xmin = range[0]; g.ymin = range[1]; g.zmin = range[2];
  g.gx = grid[0]; g.gy = grid[1]; g.gz = grid[2];
  return g;
}

}  // namespace bevpool

using namespace bevpool;

extern "C" int bevvox_temp_bytes(int batch, int64_t total_points, int max_voxels, int max_points,
                                 size_t *temp_bytes) {
  int rc = check_vox_args(batch, total_points, 3, max_voxels, max_points);
  if (rc) return rc;
  if (!temp_bytes) return BEVPOOL_E_ARG;
  *temp_bytes = vox_temp_layout(batch, total_points, max_voxels, max_points).bytes;
  return BEVPOOL_OK;
}

extern "C" int bevvox_hard_voxelize(const float *points, const int32_t *sample_offsets, int batch,
                                    int64_t total_points, int64_t max_sample_points, int num_features,
                                    const float *voxel_size_host, const float *range_host,
                                    const int *grid_host, int max_points, int max_voxels, float *voxels,
                                    int32_t *coors, int32_t *num_points, int32_t *voxel_base,
                                    float *voxel_mean, int mean_features, void *temp, void *stream_) {
  int rc = check_vox_args(batch, total_points, num_features, max_voxels, max_points);
  if (rc) return rc;
  if (!sample_offsets || !voxel_size_host || !range_host || !grid_host || !voxels || !coors || !num_points ||
      !voxel_base || !temp)
    return BEVPOOL_E_ARG;
  if (total_points > 0 && !points) return BEVPOOL_E_ARG;
  if (voxel_mean && (mean_features <= 0 || mean_features > num_features || mean_features > 32)) return BEVPOOL_E_ARG;
  if (!aligned16(temp) || !aligned16(coors)) return BEVPOOL_E_ALIGN;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const VoxGeom g = make_geom(voxel_size_host, range_host, grid_host);
  if (g.gx <= 0 || g.gy <= 0 || g.gz <= 0) return BEVPOOL_E_ARG;
  const VoxTemp L = vox_temp_layout(batch, total_points, max_voxels, max_points);
  char *tb = static_cast<char *>(temp);
  int32_t *counts = reinterpret_cast<int32_t *>(tb + L.off_counts);
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(tb + L.off_keys);
  int32_t *first = reinterpret_cast<int32_t *>(tb + L.off_first);
  int32_t *lists = reinterpret_cast<int32_t *>(tb + L.off_lists);
  int32_t *point_slot = reinterpret_cast<int32_t *>(tb + L.off_slot);
  uint32_t *flags = reinterpret_cast<uint32_t *>(tb + L.off_flags);

  BEVPOOL_RETURN_IF_CUDA(cudaMemsetAsync(tb, 0, L.zero_bytes, stream));
  BEVPOOL_RETURN_IF_CUDA(cudaMemsetAsync(keys, 0xff, (size_t)L.hash_slots * 8, stream));
  BEVPOOL_RETURN_IF_CUDA(cudaMemsetAsync(first, 0x7f, L.fill7f_bytes, stream));

  const dim3 pgrid((unsigned)ceil_div64(max_sample_points > 0 ? max_sample_points : 1, 256), (unsigned)batch);
  if (total_points > 0) {
    vox_insert_kernel<<<pgrid, 256, 0, stream>>>(points, sample_offsets, num_features, g, keys, first,
                                                L.hash_slots - 1, point_slot);
    BEVPOOL_LAUNCH_CHECK();
  }
  vox_flag_kernel<<<(unsigned)ceil_div64(total_points + 1, 256), 256, 0, stream>>>(point_slot, first, total_points, flags);
  BEVPOOL_LAUNCH_CHECK();
  rc = launch_scan_exclusive(flags, flags, total_points + 1, tb + L.off_scan, stream);
  if (rc) return rc;
  vox_base_kernel<<<1, 32, 0, stream>>>(flags, sample_offsets, batch, max_voxels, voxel_base);
  BEVPOOL_LAUNCH_CHECK();
  if (total_points > 0) {
    vox_assign_kernel<<<pgrid, 256, 0, stream>>>(points, sample_offsets, num_features, g, point_slot, first, flags,
                                                voxel_base, max_voxels, max_points, lists, counts, coors);
    BEVPOOL_LAUNCH_CHECK();
  }
  const int64_t rows = (int64_t)batch * max_voxels;
  vox_gather_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, stream>>>(points, num_features, lists, counts, voxel_base,
                                                                      batch, max_points, voxels, num_points,
                                                                      voxel_mean, mean_features);
  BEVPOOL_LAUNCH_CHECK();
  return BEVPOOL_OK;
}

extern "C" int bevvox_dynamic_voxelize(const float *points, int64_t num_points, int num_features,
                                       const float *voxel_size_host, const float *range_host,
                                       const int *grid_host, int32_t *coors, void *stream_) {
  if (num_points < 0 || num_features < 3 || !voxel_size_host || !range_host || !grid_host) return BEVPOOL_E_ARG;
  if (num_points == 0) return BEVPOOL_OK;
  if (!points || !coors) return BEVPOOL_E_ARG;
  const VoxGeom g = make_geom(voxel_size_host, range_host, grid_host);
  vox_dynamic_kernel<<<(unsigned)ceil_div64(num_points, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      points, num_points, num_features, g, coors);
  BEVPOOL_LAUNCH_CHECK();
  return BEVPOOL_OK;
}

extern "C" int pillar_scatter_forward(const void *voxel_features, const int32_t *coors, int64_t num_voxels,
                                      int channels, int dtype, int batch, int nz, int ny, int nx,
                                      void *canvas, int32_t *index_map, void *stream_) {
  if (num_voxels < 0 || channels <= 0 || batch <= 0 || batch > 65535 || nz <= 0 || ny <= 0 || nx <= 0) return BEVPOOL_E_ARG;
  if (!canvas || !index_map || (num_voxels > 0 && (!voxel_features || !coors))) return BEVPOOL_E_ARG;
  if (num_voxels > 0 && !aligned16(coors)) return BEVPOOL_E_ALIGN;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t cells = (int64_t)nz * ny * nx;
  BEVPOOL_RETURN_IF_CUDA(cudaMemsetAsync(index_map, 0xff, (size_t)batch * cells * 4, stream));
  if (num_voxels > 0) {
    scatter_index_kernel<<<(unsigned)ceil_div64(num_voxels, 256), 256, 0, stream>>>(coors, num_voxels, batch, nz, ny, nx, index_map);
    BEVPOOL_LAUNCH_CHECK();
  }
  const dim3 grid((unsigned)ceil_div64(cells, 256), (unsigned)batch);
  switch (dtype) {
    case BEVPOOL_F32:
      scatter_canvas_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float *>(voxel_features), index_map, channels, cells, static_cast<float *>(canvas));
      break;
    case BEVPOOL_F16:
      scatter_canvas_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half *>(voxel_features), index_map, channels, cells, static_cast<__half *>(canvas));
      break;
    case BEVPOOL_BF16:
      scatter_canvas_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(voxel_features), index_map, channels, cells, static_cast<__nv_bfloat16 *>(canvas));
      break;
    default: return BEVPOOL_E_DTYPE;
  }
  BEVPOOL_LAUNCH_CHECK();
  return BEVPOOL_OK;
}

extern "C" int pillar_scatter_backward(const void *grad_canvas, const int32_t *coors, int64_t num_voxels,
                                       int channels, int dtype, int batch, int nz, int ny, int nx,
                                       void *grad_voxel_features, void *stream_) {
  if (num_voxels < 0 || channels <= 0 || batch <= 0 || nz <= 0 || ny <= 0 || nx <= 0) return BEVPOOL_E_ARG;
  if (num_voxels == 0) return BEVPOOL_OK;
  if (!grad_canvas || !coors || !grad_voxel_features) return BEVPOOL_E_ARG;
  if (!aligned16(coors)) return BEVPOOL_E_ALIGN;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const unsigned grid = (unsigned)ceil_div64(num_voxels * channels, 256);
  switch (dtype) {
    case BEVPOOL_F32:
      scatter_backward_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float *>(grad_canvas), coors, num_voxels, channels, batch, nz, ny, nx, static_cast<float *>(grad_voxel_features));
      break;
    case BEVPOOL_F16:
      scatter_backward_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half *>(grad_canvas), coors, num_voxels, channels, batch, nz, ny, nx, static_cast<__half *>(grad_voxel_features));
      break;
    case BEVPOOL_BF16:
      scatter_backward_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(grad_canvas), coors, num_voxels, channels, batch, nz, ny, nx, static_cast<__nv_bfloat16 *>(grad_voxel_features));
      break;
    default: return BEVPOOL_E_DTYPE;
  }
  BEVPOOL_LAUNCH_CHECK();
  return BEVPOOL_OK;
}
