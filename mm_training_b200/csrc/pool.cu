// BEV voxel pooling kernels (sm_100a): deterministic sorted-interval segmented
// reduction (forward) and gather (backward), for both the drop-in op and the fused
// depth (x) context entry point.
//
// Reference behaviour being replaced:
//   forward  : ops/voxel_pooling/src/voxel_pooling_forward_cuda.cu:30-34 (C scalar fp32
//              atomicAdd per point) and, for the fused entry, the materialised outer
//              product of layers/backbones/lss_fpn.py:441-463
//   backward : ops/voxel_pooling/voxel_pooling.py:58-69 (mask + advanced-index gather)
//
// HBM-bound gather/scatter with <= 2 flop/byte: no tensor cores.  Every output
// element is written exactly once by exactly one thread, in a fixed order, so results
// are bit-stable run to run (the reference's are not).
#include "common.cuh"

namespace bevpool {

// ---- element access: 4 consecutive channels <-> float4 -------------------------------
template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ float4 load(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
  static __device__ __forceinline__ float4 load_stream(const float *p) { return ldg_stream_f4(reinterpret_cast<const float4 *>(p)); }
  static __device__ __forceinline__ void store(float *p, const float4 &v) { stg_stream_f4(reinterpret_cast<float4 *>(p), v); }
  static __device__ __forceinline__ float to_float(float v) { return v; }
  static __device__ __forceinline__ float from_float(float v) { return v; }
};
template <> struct Vec4<__half> {
  static __device__ __forceinline__ float4 load(const __half *p) {
    const uint2 u = __ldg(reinterpret_cast<const uint2 *>(p));
    const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ float4 load_stream(const __half *p) { return load(p); }
  static __device__ __forceinline__ void store(__half *p, const float4 &v) {
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<unsigned *>(&a);
    u.y = *reinterpret_cast<unsigned *>(&b);
    *reinterpret_cast<uint2 *>(p) = u;
  }
  static __device__ __forceinline__ float to_float(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_float(float v) { return __float2half_rn(v); }
};
template <> struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 load(const __nv_bfloat16 *p) {
    const uint2 u = __ldg(reinterpret_cast<const uint2 *>(p));
    // bf16 -> f32 is a 16-bit shift
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u),
                       __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
  }
  static __device__ __forceinline__ float4 load_stream(const __nv_bfloat16 *p) { return load(p); }
  static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float4 &v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<unsigned *>(&a);
    u.y = *reinterpret_cast<unsigned *>(&b);
    *reinterpret_cast<uint2 *>(p) = u;
  }
  static __device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_float(float v) { return __float2bfloat16_rn(v); }
};

constexpr int kPoolThreads = 256;
constexpr int kPoolWarps = kPoolThreads / 32;
constexpr int kUnroll = 8;  // feature rows in flight per warp

// ---- forward: one warp per BEV cell, lanes own float4 channel chunks -----------------
// kFused = false: rows = feature rows (B*Np, C), streamed from HBM once.
// kFused = true : rows = context rows (B*N*H*W, C) (L1/L2 resident), scaled by depth[p].
// The cell's points are visited in ascending point order (stable plan), with separate
// multiply and add (no FMA contraction), so the fp32 result is bit-identical to a
// sequential scatter-add over the materialised tensor.
template <typename T, int CPL, bool kFused>
__global__ void __launch_bounds__(kPoolThreads)
pool_forward_kernel(const int32_t *__restrict__ cell_start, const int32_t *__restrict__ sorted_ids,
                    const T *__restrict__ rows, const T *__restrict__ depth, T *__restrict__ out,
                    int64_t total_cells, int C, int dhw, int hw) {
  const int lane = threadIdx.x & 31;
  const int64_t cell = (int64_t)blockIdx.x * kPoolWarps + (threadIdx.x >> 5);
  if (cell >= total_cells) return;
  const int C4 = C >> 2;
  const int start = cell_start[cell], end = cell_start[cell + 1];

  float4 acc[CPL];
#pragma unroll
  for (int k = 0; k < CPL; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int base = start; base < end; base += 32) {
    const int n = min(32, end - base);
    int my_row = 0;
    float my_d = 1.f;
    if (lane < n) {
      const int gp = sorted_ids[base + lane];
      if (kFused) {
        my_d = Vec4<T>::to_float(depth[gp]);
        my_row = (gp / dhw) * hw + gp % hw;   // pixel row: (b*N+n)*H*W + h*W + w
      } else {
        my_row = gp;
      }
    }
    for (int j = 0; j < n; j += kUnroll) {
      float4 v[kUnroll][CPL];
      float d[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int src = (j + u) & 31;
        const int row = __shfl_sync(0xffffffffu, my_row, src);
        d[u] = __shfl_sync(0xffffffffu, my_d, src);
        if (j + u < n) {
          const T *rp = rows + (int64_t)row * C;
#pragma unroll
          for (int k = 0; k < CPL; ++k) {
            const int ch = lane + 32 * k;
            if (ch < C4) v[u][k] = kFused ? Vec4<T>::load(rp + ch * 4) : Vec4<T>::load_stream(rp + ch * 4);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        if (j + u < n) {
#pragma unroll
          for (int k = 0; k < CPL; ++k) {
            if (lane + 32 * k < C4) {
              if (kFused) {
                acc[k].x = __fadd_rn(acc[k].x, __fmul_rn(d[u], v[u][k].x));
                acc[k].y = __fadd_rn(acc[k].y, __fmul_rn(d[u], v[u][k].y));
                acc[k].z = __fadd_rn(acc[k].z, __fmul_rn(d[u], v[u][k].z));
                acc[k].w = __fadd_rn(acc[k].w, __fmul_rn(d[u], v[u][k].w));
              } else {
                acc[k].x += v[u][k].x;
                acc[k].y += v[u][k].y;
                acc[k].z += v[u][k].z;
                acc[k].w += v[u][k].w;
              }
            }
          }
        }
      }
    }
  }
  T *op = out + cell * C;
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    const int ch = lane + 32 * k;
    if (ch < C4) Vec4<T>::store(op + ch * 4, acc[k]);
  }
}

// ---- backward of the drop-in op: point-centric gather ---------------------------------
// grad_features[p, :] = kept(p) ? grad_out_nhwc[b, cell(p), :] : 0 ; each row written once.
constexpr int kBwdPoints = 64;  // points per CTA
template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
pool_backward_kernel(const int32_t *__restrict__ cell_of_point, const T *__restrict__ grad_nhwc,
                     T *__restrict__ grad_rows, int64_t total_points, int64_t num_points,
                     int64_t cells_per_sample, int C) {
  __shared__ int64_t s_row[kBwdPoints];
  const int C4 = C >> 2;
  const int64_t gp0 = (int64_t)blockIdx.x * kBwdPoints;
  if (threadIdx.x < kBwdPoints) {
    const int64_t gp = gp0 + threadIdx.x;
    int64_t row = -1;
    if (gp < total_points) {
      const int cell = cell_of_point[gp];
      if (cell >= 0) row = (gp / num_points) * cells_per_sample + cell;
    }
    s_row[threadIdx.x] = row;
  }
  __syncthreads();
  const int npts = (int)min((int64_t)kBwdPoints, total_points - gp0);
  const int work = npts * C4;
  for (int i = threadIdx.x; i < work; i += kPoolThreads) {
    const int pl = i / C4, ch = i - pl * C4;
    const int64_t row = s_row[pl];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row >= 0) v = Vec4<T>::load(grad_nhwc + row * C + ch * 4);
    Vec4<T>::store(grad_rows + (gp0 + pl) * C + ch * 4, v);
  }
}

// ---- backward of the fused op: pixel-centric, no atomics, no sort ----------------------
// One CTA = kPixTile consecutive pixels of one camera image, one warp per pixel.  The warp
// keeps its context row in registers and walks the D depth bins of its ray:
//   grad_depth[d, pix]  = <grad_out[cell(d, pix), :], context[pix, :]>
//   grad_context[pix,:] = sum_d depth[d, pix] * grad_out[cell(d, pix), :]
// depth / cell / grad_depth columns and the NCHW context rows are staged through shared
// memory so that every global access is a full 32-byte sector.
constexpr int kPixTile = 8;
template <typename T, int CPL>
__global__ void __launch_bounds__(kPixTile * 32)
fused_backward_kernel(const int32_t *__restrict__ cell_of_point, const T *__restrict__ grad_nhwc,
                      const T *__restrict__ depth, const T *__restrict__ ctx_nchw,
                      T *__restrict__ grad_depth, T *__restrict__ grad_ctx_nchw, int num_cams, int D,
                      int HW, int C, int64_t cells_per_sample) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int32_t *s_cell = reinterpret_cast<int32_t *>(smem_raw);             // [D][kPixTile]
  float *s_depth = reinterpret_cast<float *>(s_cell + D * kPixTile);   // [D][kPixTile]
  float *s_gd = s_depth + D * kPixTile;                                // [D][kPixTile]
  float *s_ctx = s_gd + D * kPixTile;                                  // [C][kPixTile] (in: ctx, out: grad)
  const int bn = blockIdx.y;
  const int hw0 = blockIdx.x * kPixTile;
  const int npix = min(kPixTile, HW - hw0);
  const int lane = threadIdx.x & 31, pix = threadIdx.x >> 5;
  const int C4 = C >> 2;
  const int64_t img_base = (int64_t)bn * D * HW;   // == first global point id of this image

  for (int i = threadIdx.x; i < D * kPixTile; i += kPixTile * 32) {
    const int d = i / kPixTile, j = i - d * kPixTile;
    int cell = -1;
    float dv = 0.f;
    if (j < npix) {
      const int64_t gp = img_base + (int64_t)d * HW + hw0 + j;
      cell = cell_of_point[gp];
      dv = Vec4<T>::to_float(depth[gp]);
    }
    s_cell[i] = cell;
    s_depth[i] = dv;
  }
  for (int i = threadIdx.x; i < C * kPixTile; i += kPixTile * 32) {
    const int c = i / kPixTile, j = i - c * kPixTile;
    s_ctx[i] = j < npix ? Vec4<T>::to_float(ctx_nchw[((int64_t)bn * C + c) * HW + hw0 + j]) : 0.f;
  }
  __syncthreads();

  float4 cx[CPL], gacc[CPL];
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    const int ch = lane + 32 * k;
    gacc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    cx[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ch < C4) {
      cx[k].x = s_ctx[(ch * 4 + 0) * kPixTile + pix];
      cx[k].y = s_ctx[(ch * 4 + 1) * kPixTile + pix];
      cx[k].z = s_ctx[(ch * 4 + 2) * kPixTile + pix];
      cx[k].w = s_ctx[(ch * 4 + 3) * kPixTile + pix];
    }
  }
  const T *gbase = grad_nhwc + (int64_t)(bn / num_cams) * cells_per_sample * C;
  constexpr int U = 4;
  for (int d0 = 0; d0 < D; d0 += U) {
    float4 g[U][CPL];
    int cell[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      cell[u] = d0 + u < D ? s_cell[(d0 + u) * kPixTile + pix] : -1;
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const int ch = lane + 32 * k;
        g[u][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cell[u] >= 0 && ch < C4) g[u][k] = Vec4<T>::load(gbase + (int64_t)cell[u] * C + ch * 4);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (d0 + u < D) {
        float dot = 0.f;
        const float dv = s_depth[(d0 + u) * kPixTile + pix];
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          dot += g[u][k].x * cx[k].x + g[u][k].y * cx[k].y + g[u][k].z * cx[k].z + g[u][k].w * cx[k].w;
          gacc[k].x += dv * g[u][k].x;
          gacc[k].y += dv * g[u][k].y;
          gacc[k].z += dv * g[u][k].z;
          gacc[k].w += dv * g[u][k].w;
        }
        if (cell[u] >= 0) dot = warp_sum(dot);     // warp-uniform branch
        if (lane == 0) s_gd[(d0 + u) * kPixTile + pix] = dot;
      }
    }
  }
  __syncthreads();   // everyone is done reading s_ctx as input
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    const int ch = lane + 32 * k;
    if (ch < C4) {
      s_ctx[(ch * 4 + 0) * kPixTile + pix] = gacc[k].x;
      s_ctx[(ch * 4 + 1) * kPixTile + pix] = gacc[k].y;
      s_ctx[(ch * 4 + 2) * kPixTile + pix] = gacc[k].z;
      s_ctx[(ch * 4 + 3) * kPixTile + pix] = gacc[k].w;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D * kPixTile; i += kPixTile * 32) {
    const int d = i / kPixTile, j = i - d * kPixTile;
    if (j < npix) grad_depth[img_base + (int64_t)d * HW + hw0 + j] = Vec4<T>::from_float(s_gd[i]);
  }
  for (int i = threadIdx.x; i < C * kPixTile; i += kPixTile * 32) {
    const int c = i / kPixTile, j = i - c * kPixTile;
    if (j < npix) grad_ctx_nchw[((int64_t)bn * C + c) * HW + hw0 + j] = Vec4<T>::from_float(s_ctx[i]);
  }
}

// ---- (batch, R, Cc) -> (batch, Cc, R) tiled transpose ----------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
transpose_kernel(const T *__restrict__ in, T *__restrict__ out, int64_t R, int64_t Cc) {
  __shared__ T tile[32][33];
  const int64_t batch_off = (int64_t)blockIdx.z * R * Cc;
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i, c = c0 + tx;
    if (r < R && c < Cc) tile[i][tx] = in[batch_off + r * Cc + c];
  }
  __syncthreads();
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int64_t c = c0 + i, r = r0 + tx;
    if (r < R && c < Cc) out[batch_off + c * R + r] = tile[tx][i];
  }
}

// ---- launchers ----------------------------------------------------------------------------
template <typename T, bool kFused>
static int launch_forward(const PlanView &pv, const T *rows, const T *depth, T *out,
                          int64_t total_cells, int C, int dhw, int hw, cudaStream_t s) {
  const unsigned grid = (unsigned)ceil_div64(total_cells, kPoolWarps);
  const int C4 = C >> 2;
  if (C4 <= 32)
    pool_forward_kernel<T, 1, kFused><<<grid, kPoolThreads, 0, s>>>(pv.cell_start, pv.sorted_ids, rows, depth, out, total_cells, C, dhw, hw);
  else if (C4 <= 64)
    pool_forward_kernel<T, 2, kFused><<<grid, kPoolThreads, 0, s>>>(pv.cell_start, pv.sorted_ids, rows, depth, out, total_cells, C, dhw, hw);
  else
    pool_forward_kernel<T, 4, kFused><<<grid, kPoolThreads, 0, s>>>(pv.cell_start, pv.sorted_ids, rows, depth, out, total_cells, C, dhw, hw);
  BEVPOOL_LAUNCH_CHECK();
  return BEVPOOL_OK;
}

template <typename T, int CPL>
static int launch_fused_backward_cpl(const PlanView &pv, const T *grad, const T *depth, const T *ctx,
                                     T *gdepth, T *gctx, int batch, int N, int D, int HW, int C,
                                     int64_t cells, cudaStream_t s) {
  const size_t smem = (size_t)D * kPixTile * 12 + (size_t)C * kPixTile * 4;
  if (smem > 48 * 1024) {
    BEVPOOL_RETURN_IF_CUDA(cudaFuncSetAttribute(fused_backward_kernel<T, CPL>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const dim3 grid((unsigned)ceil_div64(HW, kPixTile), (unsigned)(batch * N));
  fused_backward_kernel<T, CPL><<<grid, kPixTile * 32, smem, s>>>(pv.cell_of_point, grad, depth, ctx,
                                                                gdepth, gctx, N, D, HW, C, cells);
  BEVPOOL_LAUNCH_CHECK();
  return BEVPOOL_OK;
}

static int check_channels(int C) {
  if (C <= 0 || (C & 3) != 0 || C > 512) return BEVPOOL_E_CHANNELS;
  return BEVPOOL_OK;
}

template <typename T>
static int forward_t(const void *plan, const void *feats, void *out, int B, int64_t Np, int C, int X,
                     int Y, cudaStream_t s) {
  const PlanView pv = plan_view(plan, B, Np, X, Y);
  return launch_forward<T, false>(pv, static_cast<const T *>(feats), nullptr, static_cast<T *>(out),
                                  (int64_t)B * X * Y, C, 1, 1, s);
}
template <typename T>
static int backward_t(const void *plan, const void *grad, void *gfeats, int B, int64_t Np, int C, int X,
                      int Y, cudaStream_t s) {
  const PlanView pv = plan_view(plan, B, Np, X, Y);
  const int64_t total = (int64_t)B * Np;
  pool_backward_kernel<T><<<(unsigned)ceil_div64(total, kBwdPoints), kPoolThreads, 0, s>>>(
      pv.cell_of_point, static_cast<const T *>(grad), static_cast<T *>(gfeats), total, Np,
      (int64_t)X * Y, C);
  BEVPOOL_LAUNCH_CHECK();
  return BEVPOOL_OK;
}
template <typename T>
static int fused_forward_t(const void *plan, const void *depth, const void *ctx, void *out, int B, int N,
                           int D, int H, int W, int C, int X, int Y, cudaStream_t s) {
  const int64_t Np = (int64_t)N * D * H * W;
  const PlanView pv = plan_view(plan, B, Np, X, Y);
  return launch_forward<T, true>(pv, static_cast<const T *>(ctx), static_cast<const T *>(depth),
                                 static_cast<T *>(out), (int64_t)B * X * Y, C, D * H * W, H * W, s);
}
template <typename T>
static int fused_backward_t(const void *plan, const void *grad, const void *depth, const void *ctx,
                            void *gdepth, void *gctx, int B, int N, int D, int H, int W, int C, int X,
                            int Y, cudaStream_t s) {
  const int64_t Np = (int64_t)N * D * H * W;
  const PlanView pv = plan_view(plan, B, Np, X, Y);
  const int C4 = C >> 2;
  const T *g = static_cast<const T *>(grad), *dp = static_cast<const T *>(depth), *cx = static_cast<const T *>(ctx);
  T *gd = static_cast<T *>(gdepth), *gc = static_cast<T *>(gctx);
  const int64_t cells = (int64_t)X * Y;
  if (C4 <= 32) return launch_fused_backward_cpl<T, 1>(pv, g, dp, cx, gd, gc, B, N, D, H * W, C, cells, s);
  if (C4 <= 64) return launch_fused_backward_cpl<T, 2>(pv, g, dp, cx, gd, gc, B, N, D, H * W, C, cells, s);
  return launch_fused_backward_cpl<T, 4>(pv, g, dp, cx, gd, gc, B, N, D, H * W, C, cells, s);
}

}  // namespace bevpool

using namespace bevpool;

#define BEVPOOL_DISPATCH_DTYPE(dtype, CALL)                        \
  switch (dtype) {                                                 \
    case BEVPOOL_F32: { using T = float; return CALL; }            \
    case BEVPOOL_F16: { using T = __half; return CALL; }           \
    case BEVPOOL_BF16: { using T = __nv_bfloat16; return CALL; }   \
    default: return BEVPOOL_E_DTYPE;                               \
  }

extern "C" int bevpool_forward(const void *plan, const void *features, void *out_nhwc, int dtype,
                               int batch, int64_t num_points, int channels, int X, int Y, void *stream) {
  int rc = check_plan_dims(batch, num_points, X, Y);
  if (rc) return rc;
  if ((rc = check_channels(channels))) return rc;
  if (!plan || !features || !out_nhwc) return BEVPOOL_E_ARG;
  if (!aligned16(features) || !aligned16(out_nhwc)) return BEVPOOL_E_ALIGN;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  BEVPOOL_DISPATCH_DTYPE(dtype, (forward_t<T>(plan, features, out_nhwc, batch, num_points, channels, X, Y, s)));
}

extern "C" int bevpool_backward(const void *plan, const void *grad_out_nhwc, void *grad_features,
                                int dtype, int batch, int64_t num_points, int channels, int X, int Y,
                                void *stream) {
  int rc = check_plan_dims(batch, num_points, X, Y);
  if (rc) return rc;
  if ((rc = check_channels(channels))) return rc;
  if (!plan || !grad_out_nhwc || !grad_features) return BEVPOOL_E_ARG;
  if (!aligned16(grad_out_nhwc) || !aligned16(grad_features)) return BEVPOOL_E_ALIGN;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  BEVPOOL_DISPATCH_DTYPE(dtype, (backward_t<T>(plan, grad_out_nhwc, grad_features, batch, num_points, channels, X, Y, s)));
}

extern "C" int bevpool_fused_forward(const void *plan, const void *depth, const void *context_nhwc,
                                     void *out_nhwc, int dtype, int batch, int num_cams, int depth_bins,
                                     int feat_h, int feat_w, int channels, int X, int Y, void *stream) {
  if (num_cams <= 0 || depth_bins <= 0 || feat_h <= 0 || feat_w <= 0) return BEVPOOL_E_ARG;
  const int64_t np = (int64_t)num_cams * depth_bins * feat_h * feat_w;
  int rc = check_plan_dims(batch, np, X, Y);
  if (rc) return rc;
  if ((rc = check_channels(channels))) return rc;
  if (!plan || !depth || !context_nhwc || !out_nhwc) return BEVPOOL_E_ARG;
  if (!aligned16(context_nhwc) || !aligned16(out_nhwc)) return BEVPOOL_E_ALIGN;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  BEVPOOL_DISPATCH_DTYPE(dtype, (fused_forward_t<T>(plan, depth, context_nhwc, out_nhwc, batch, num_cams, depth_bins, feat_h, feat_w, channels, X, Y, s)));
}

extern "C" int bevpool_fused_backward(const void *plan, const void *grad_out_nhwc, const void *depth,
                                      const void *context_nchw, void *grad_depth, void *grad_context_nchw,
                                      int dtype, int batch, int num_cams, int depth_bins, int feat_h,
                                      int feat_w, int channels, int X, int Y, void *stream) {
  if (num_cams <= 0 || depth_bins <= 0 || feat_h <= 0 || feat_w <= 0) return BEVPOOL_E_ARG;
  const int64_t np = (int64_t)num_cams * depth_bins * feat_h * feat_w;
  int rc = check_plan_dims(batch, np, X, Y);
  if (rc) return rc;
  if ((rc = check_channels(channels))) return rc;
  if (!plan || !grad_out_nhwc || !depth || !context_nchw || !grad_depth || !grad_context_nchw) return BEVPOOL_E_ARG;
  if (!aligned16(grad_out_nhwc)) return BEVPOOL_E_ALIGN;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  BEVPOOL_DISPATCH_DTYPE(dtype, (fused_backward_t<T>(plan, grad_out_nhwc, depth, context_nchw, grad_depth, grad_context_nchw, batch, num_cams, depth_bins, feat_h, feat_w, channels, X, Y, s)));
}

extern "C" int bevpool_transpose(const void *in, void *out, int dtype, int batch, int64_t rows,
                                 int64_t cols, void *stream) {
  if (!in || !out || batch <= 0 || rows <= 0 || cols <= 0) return BEVPOOL_E_ARG;
  if (batch > 65535 || ceil_div64(rows, 32) > 65535) return BEVPOOL_E_RANGE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const dim3 grid((unsigned)ceil_div64(cols, 32), (unsigned)ceil_div64(rows, 32), (unsigned)batch);
  if (dtype == BEVPOOL_F32) {
    transpose_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float *>(in), static_cast<float *>(out), rows, cols);
  } else if (dtype == BEVPOOL_F16 || dtype == BEVPOOL_BF16) {
    transpose_kernel<uint16_t><<<grid, 256, 0, s>>>(static_cast<const uint16_t *>(in), static_cast<uint16_t *>(out), rows, cols);
  } else {
    return BEVPOOL_E_DTYPE;
  }
  BEVPOOL_LAUNCH_CHECK();
  return BEVPOOL_OK;
}
