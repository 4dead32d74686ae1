// Depth-label generation for the depth loss (sm_100a): LiDAR -> image projection, "last point wins" depth map,
// 16x16 min-pool, bin index, one-hot.
//
// Reference being replaced (exps/mm_training_aim.py): get_depth_labels :115-141 (python triple loop over batch x
// sweep x camera, ~25 ATen kernels and a matrix inverse per camera), get_depth_image :143-162 (projection, bounds
// mask, truncation to pixels, `depth_map[v, u] = depth` -- for several points in one pixel the LAST point in cloud
// order wins on the CPU, and an arbitrary one on the GPU), get_downsampled_gt_depth :180-215 (zeros -> 1e5, min over
// each downsample x downsample block, (d - (d0 - step)) / step, out-of-range -> 0, truncation, one_hot).
//
// Here: ONE launch projects every point into every image of its sample, a 64-bit atomicMax on (point index + 1,
// depth bits) per pixel implements "last point in cloud order wins" deterministically, and a second launch reduces
// the blocks and writes the (images * h * w, D) one-hot rows once, coalesced.
//
// Arithmetic (float32, every product and sum rounded separately, left to right -- no FMA contraction -- so that the
// CPU oracle reproduces it bit for bit):
//   un-augment  q_j = (x * A[j][0] + y * A[j][1]) + z * A[j][2]                 A = inverse(bda[:3,:3])
//   extrinsic   p_j = ((E[j][0] * q0 + E[j][1] * q1) + E[j][2] * q2) + E[j][3]
//   intrinsic   r_j = ((K[j][0] * p0 + K[j][1] * p1) + K[j][2] * p2) + K[j][3] * p3
//   u = r0 / r2, v = r1 / r2 (IEEE division); depth = p2
#include "common.cuh"

namespace bevpool {

struct Mat4 {
  float m[16];
};

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }

// grid: (ceil(max_points / 256), images_per_sample, batch)
__global__ void __launch_bounds__(256)
depth_project_kernel(const float *const *__restrict__ sample_ptrs, const int32_t *__restrict__ sample_counts, int F,
                     const float *__restrict__ bda_inv, const float *__restrict__ extrinsics,
                     const float *__restrict__ intrinsics, int images_per_sample, int img_h, int img_w,
                     unsigned long long *__restrict__ winner) {
  const int b = blockIdx.z, img = b * images_per_sample + blockIdx.y;
  const int n = sample_counts[b];
  __shared__ float s_a[9], s_e[16], s_k[16];
  if (threadIdx.x < 9) s_a[threadIdx.x] = bda_inv[b * 9 + threadIdx.x];
  if (threadIdx.x >= 32 && threadIdx.x < 48) s_e[threadIdx.x - 32] = extrinsics[(int64_t)img * 16 + threadIdx.x - 32];
  if (threadIdx.x >= 64 && threadIdx.x < 80) s_k[threadIdx.x - 64] = intrinsics[(int64_t)img * 16 + threadIdx.x - 64];
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float *pt = sample_ptrs[b] + (int64_t)p * F;
  const float x = __ldg(pt), y = __ldg(pt + 1), z = __ldg(pt + 2);
  float q[3], c[4], r[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) q[j] = add_rn(add_rn(mul_rn(x, s_a[j * 3 + 0]), mul_rn(y, s_a[j * 3 + 1])), mul_rn(z, s_a[j * 3 + 2]));
#pragma unroll
  for (int j = 0; j < 4; ++j)
    c[j] = add_rn(add_rn(add_rn(mul_rn(s_e[j * 4 + 0], q[0]), mul_rn(s_e[j * 4 + 1], q[1])), mul_rn(s_e[j * 4 + 2], q[2])), s_e[j * 4 + 3]);
#pragma unroll
  for (int j = 0; j < 3; ++j)
    r[j] = add_rn(add_rn(add_rn(mul_rn(s_k[j * 4 + 0], c[0]), mul_rn(s_k[j * 4 + 1], c[1])), mul_rn(s_k[j * 4 + 2], c[2])), mul_rn(s_k[j * 4 + 3], c[3]));
  const float depth = c[2];
  const float u = __fdiv_rn(r[0], r[2]), v = __fdiv_rn(r[1], r[2]);
  // mm_training_aim.py:152-157: strict inequalities; NaN compares false
  const bool ok = depth > 1.0f && u > 1.0f && u < (float)(img_w - 1) && v > 1.0f && v < (float)(img_h - 1);
  if (!ok) return;
  const int ui = (int)u, vi = (int)v;                                 // .to(torch.long): truncation
  const unsigned long long key = ((unsigned long long)(unsigned)(p + 1) << 32) | (unsigned long long)__float_as_uint(depth);
  atomicMax(winner + ((int64_t)img * img_h + vi) * img_w + ui, key);
}

// one warp per downsampled cell; the winner map is cleared behind the read (the scratch is clean for the next call)
__global__ void __launch_bounds__(256)
depth_minpool_onehot_kernel(unsigned long long *__restrict__ winner, int img_h, int img_w, int ds, float bin_offset,
                            float bin_step, int D, int64_t num_cells, float *__restrict__ labels,
                            int32_t *__restrict__ bins) {
  const int lane = threadIdx.x & 31;
  const int64_t cell = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (cell >= num_cells) return;
  const int fw = img_w / ds, fh = img_h / ds;
  const int wb = (int)(cell % fw), hb = (int)((cell / fw) % fh);
  const int64_t img = cell / ((int64_t)fw * fh);
  float m = 1e5f;                                                      // zeros -> 1e5 (:199-201)
  for (int e = lane; e < ds * ds; e += 32) {
    const int dy = e / ds, dx = e % ds;
    unsigned long long *w = winner + (img * img_h + (int64_t)hb * ds + dy) * img_w + (int64_t)wb * ds + dx;
    const unsigned long long key = *w;
    if (key) {
      *w = 0ull;
      const float d = __uint_as_float((unsigned)(key & 0xffffffffull));
      if (d != 0.0f) m = fminf(m, d);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float g = __fdiv_rn(__fsub_rn(m, bin_offset), bin_step);       // (:207-208)
  const float gz = (g < (float)D && g >= 0.0f) ? g : 0.0f;            // (:209-211)
  const int bin = (int)gz;                                             // .long()
  if (bins && lane == 0) bins[cell] = bin;
  float *row = labels + cell * D;
  for (int d = lane; d < D; d += 32) row[d] = d == bin ? 1.0f : 0.0f;
}

}  // namespace bevpool

using namespace bevpool;

extern "C" int bevlabel_scratch_bytes(int num_images, int img_h, int img_w, size_t *bytes) {
  if (num_images <= 0 || img_h <= 0 || img_w <= 0 || !bytes) return BEVPOOL_E_ARG;
  *bytes = (size_t)num_images * img_h * img_w * 8;
  return BEVPOOL_OK;
}

extern "C" int bevlabel_depth_labels(const float *const *sample_ptrs, const int32_t *sample_counts, int num_features,
                                     int batch, int images_per_sample, int64_t max_points, const float *bda_inv,
                                     const float *extrinsics, const float *intrinsics, int img_h, int img_w,
                                     int downsample, float bin_offset, float bin_step, int depth_channels,
                                     float *labels, int32_t *bins, void *scratch, int scratch_is_clean, void *stream) {
  if (batch <= 0 || images_per_sample <= 0 || num_features < 3 || img_h <= 0 || img_w <= 0 || downsample <= 0 ||
      depth_channels <= 0 || max_points < 0)
    return BEVPOOL_E_ARG;
  if (!sample_ptrs || !sample_counts || !bda_inv || !extrinsics || !intrinsics || !labels || !scratch) return BEVPOOL_E_ARG;
  if (img_h % downsample || img_w % downsample) return BEVPOOL_E_ARG;        // the reference's view() needs it too
  if (max_points >= (int64_t)INT32_MAX) return BEVPOOL_E_RANGE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t images = (int64_t)batch * images_per_sample;
  unsigned long long *winner = static_cast<unsigned long long *>(scratch);
  if (!scratch_is_clean) BEVPOOL_RETURN_IF_CUDA(cudaMemsetAsync(winner, 0, (size_t)images * img_h * img_w * 8, s));
  if (max_points > 0) {
    dim3 grid((unsigned)ceil_div64(max_points, 256), (unsigned)images_per_sample, (unsigned)batch);
    depth_project_kernel<<<grid, 256, 0, s>>>(sample_ptrs, sample_counts, num_features, bda_inv, extrinsics, intrinsics,
                                               images_per_sample, img_h, img_w, winner);
    BEVPOOL_LAUNCH_CHECK();
  }
  const int64_t cells = images * (img_h / downsample) * (img_w / downsample);
  depth_minpool_onehot_kernel<<<(unsigned)ceil_div64(cells, 8), 256, 0, s>>>(winner, img_h, img_w, downsample, bin_offset,
                                                                              bin_step, depth_channels, cells, labels, bins);
  BEVPOOL_LAUNCH_CHECK();
  return BEVPOOL_OK;
}
