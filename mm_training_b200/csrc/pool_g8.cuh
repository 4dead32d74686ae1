// fp32 fast path of the pooling kernels: 8 lanes per feature row ("g8" layout).
//
// A row of C = 16*NV2 floats (C = 80 -> NV2 = 5) is owned by a group of 8 lanes; lane l holds
// channels [32k + 4l, +4) for k < NV2/2 (float4 loads) plus, when NV2 is odd, the float2 at
// [32*(NV2/2) + 2l, +2).  A warp therefore works on 4 rows per instruction with all 32 lanes
// busy (the generic float4-per-lane layout keeps 20 of 32 lanes busy at C = 80 and spends
// more than half of its issue slots on addressing and predicates).
//
// Same numerics contract as the generic kernels: per-cell sums run in ascending point order
// with separate multiply and add, so the forward is bit-identical to a sequential scatter-add.
#pragma once
#include "common.cuh"

namespace bevpool {

constexpr int kG8Chunk = 256;   // points of a warp's 4 cells staged in shared memory at a time

__device__ __forceinline__ float4 ld_stream_or_cached_f4(const char *p, bool stream) {
  return stream ? ldg_stream_f4(reinterpret_cast<const float4 *>(p)) : __ldg(reinterpret_cast<const float4 *>(p));
}

// loads the NREG floats of one row for lane l8; `row` points at the row's first byte
template <int NV2, bool kStream>
__device__ __forceinline__ void g8_load_row(const char *row, int l8, float (&v)[2 * NV2]) {
  constexpr int NF4 = NV2 / 2, NF2 = NV2 & 1;
#pragma unroll
  for (int k = 0; k < NF4; ++k) {
    const float4 t = ld_stream_or_cached_f4(row + 128 * k + 16 * l8, kStream);
    v[4 * k + 0] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
  }
  if (NF2) {
    const float2 t = __ldg(reinterpret_cast<const float2 *>(row + 128 * NF4 + 8 * l8));
    v[4 * NF4 + 0] = t.x; v[4 * NF4 + 1] = t.y;
  }
}

template <int NV2>
__device__ __forceinline__ void g8_store_row(char *row, int l8, const float (&v)[2 * NV2]) {
  constexpr int NF4 = NV2 / 2, NF2 = NV2 & 1;
#pragma unroll
  for (int k = 0; k < NF4; ++k)
    stg_stream_f4(reinterpret_cast<float4 *>(row + 128 * k + 16 * l8),
                  make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
  if (NF2) *reinterpret_cast<float2 *>(row + 128 * NF4 + 8 * l8) = make_float2(v[4 * NF4], v[4 * NF4 + 1]);
}

// channel index of register r of lane l8 (for the NCHW context accesses of the backward)
template <int NV2>
__device__ __forceinline__ int g8_channel(int r, int l8) {
  constexpr int NF4 = NV2 / 2;
  return r < 4 * NF4 ? 32 * (r >> 2) + 4 * l8 + (r & 3) : 32 * NF4 + 2 * l8 + (r - 4 * NF4);
}

// ---- forward ---------------------------------------------------------------------------------
// Warp-autonomous: every warp owns 4 consecutive BEV cells (one 8-lane group per cell) and never
// synchronises with the rest of its CTA, so the SM's warp scheduler hides the dependent
// cell_start -> sorted ids -> depth -> context-row latencies across ~24 independent warps.
// The 4 cells' points are contiguous in the plan's sorted list; the warp stages (row index,
// depth) for up to kG8Chunk of them in its private shared-memory slice with coalesced loads,
// then every group walks its own cell's interval in order.
constexpr int kG8FwdWarps = 4;
constexpr int kG8FwdCellsPerCta = 4 * kG8FwdWarps;

template <int NV2, bool kFused>
__global__ void __launch_bounds__(kG8FwdWarps * 32, 6)
pool_forward_g8_kernel(const int32_t *__restrict__ cell_start, const int32_t *__restrict__ sorted_ids,
                       const float *__restrict__ rows, const float *__restrict__ depth,
                       float *__restrict__ out, int64_t total_cells, int dhw, int hw) {
  constexpr int C = 16 * NV2, NREG = 2 * NV2, U = 4;
  __shared__ uint2 s_pts_all[kG8FwdWarps][kG8Chunk];
  const int lane = threadIdx.x & 31, l8 = lane & 7, grp = lane >> 3, warp = threadIdx.x >> 5;
  uint2 *s_pts = s_pts_all[warp];
  const int64_t cell0 = ((int64_t)blockIdx.x * kG8FwdWarps + warp) * 4;
  if (cell0 >= total_cells) return;
  const int ncell = (int)min((int64_t)4, total_cells - cell0);
  int cs = 0;
  if (lane <= ncell) cs = __ldg(cell_start + cell0 + lane);
  const int wstart = __shfl_sync(0xffffffffu, cs, 0), wend = __shfl_sync(0xffffffffu, cs, ncell);
  const bool mine = grp < ncell;
  int my_start = __shfl_sync(0xffffffffu, cs, min(grp, ncell));
  int my_end = __shfl_sync(0xffffffffu, cs, min(grp + 1, ncell));
  if (!mine) my_start = my_end = wend;

  float acc[NREG];
#pragma unroll
  for (int r = 0; r < NREG; ++r) acc[r] = 0.f;
  const char *rows_b = reinterpret_cast<const char *>(rows);

  for (int lo = wstart; lo < wend; lo += kG8Chunk) {
    const int hi = min(lo + kG8Chunk, wend);
    if (lo != wstart) __syncwarp();
    for (int i = lane; i < hi - lo; i += 32) {
      const int gp = __ldg(sorted_ids + lo + i);
      uint2 e;
      if (kFused) {
        e.x = (unsigned)((gp / dhw) * hw + gp % hw);   // context row of the point's pixel
        e.y = __float_as_uint(__ldg(depth + gp));
      } else {
        e.x = (unsigned)gp;
        e.y = 0u;
      }
      s_pts[i] = e;
    }
    __syncwarp();
    const int a = max(my_start, lo), b = min(my_end, hi);
    int trips = max(b - a, 0);
    trips = max(trips, __shfl_xor_sync(0xffffffffu, trips, 8));
    trips = max(trips, __shfl_xor_sync(0xffffffffu, trips, 16));
    for (int t = 0; t < trips; t += U) {
      float v[U][NREG];
      float d[U];
      bool on[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = a + t + u;
        on[u] = p < b;
        if (on[u]) {
          const uint2 e = s_pts[p - lo];
          d[u] = __uint_as_float(e.y);
          g8_load_row<NV2, !kFused>(rows_b + (size_t)e.x * (C * 4), l8, v[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (on[u]) {
#pragma unroll
          for (int r = 0; r < NREG; ++r)
            acc[r] = kFused ? __fadd_rn(acc[r], __fmul_rn(d[u], v[u][r])) : acc[r] + v[u][r];
        }
      }
    }
  }
  if (mine) g8_store_row<NV2>(reinterpret_cast<char *>(out + (cell0 + grp) * C), l8, acc);
}

// ---- fused backward ---------------------------------------------------------------------------
// warp = one pixel at a time (rows h of the same image columns share a CTA for L1 locality);
// the 4 groups of the warp take 4 kept depth bins per step.  Kept bins of a 32-bin chunk are
// compacted through a per-warp shared-memory list so no issue slots are spent on dropped bins.
constexpr int kG8BwdWarps = 8;
constexpr int kG8BwdTileW = 4;

template <int NV2>
__global__ void __launch_bounds__(kG8BwdWarps * 32)
fused_backward_g8_kernel(const int32_t *__restrict__ cell_of_point, const float *__restrict__ grad_rows,
                         const float *__restrict__ depth, const float *__restrict__ ctx_nchw,
                         float *__restrict__ grad_depth, float *__restrict__ grad_ctx_nchw, int num_cams,
                         int D, int H, int W, int64_t cells_per_sample) {
  constexpr int C = 16 * NV2, NREG = 2 * NV2;
  __shared__ uint2 s_list[kG8BwdWarps][32];   // (cell, depth bits) of kept bins, compacted
  __shared__ int s_bin[kG8BwdWarps][32];      // their depth-bin index
  const int lane = threadIdx.x & 31, l8 = lane & 7, grp = lane >> 3, warp = threadIdx.x >> 5;
  const int bn = blockIdx.z;
  const int h = blockIdx.y * (blockDim.x >> 5) + warp;
  if (h >= H) return;
  const int HW = H * W;
  const int64_t img_base = (int64_t)bn * D * HW;
  const char *gbase = reinterpret_cast<const char *>(grad_rows + (int64_t)(bn / num_cams) * cells_per_sample * C);
  const int w_end = min(W, (int)(blockIdx.x + 1) * kG8BwdTileW);

  for (int w = blockIdx.x * kG8BwdTileW; w < w_end; ++w) {
    const int hw = h * W + w;
    float cx[NREG], gacc[NREG];
    const float *cp = ctx_nchw + (int64_t)bn * C * HW + hw;
#pragma unroll
    for (int r = 0; r < NREG; ++r) {
      cx[r] = __ldg(cp + (int64_t)g8_channel<NV2>(r, l8) * HW);
      gacc[r] = 0.f;
    }
    for (int d0 = 0; d0 < D; d0 += 32) {
      const int d = d0 + lane;
      const int64_t gp = img_base + (int64_t)d * HW + hw;
      int cell = -1;
      float dv = 0.f;
      if (d < D) {
        cell = __ldg(cell_of_point + gp);
        dv = __ldg(depth + gp);
      }
      const unsigned mask = __ballot_sync(0xffffffffu, cell >= 0);
      if (d < D && cell < 0) grad_depth[gp] = 0.f;
      if (mask == 0) continue;
      const int cnt = __popc(mask);
      if (cell >= 0) {
        const int rank = __popc(mask & ((1u << lane) - 1u));
        s_list[warp][rank] = make_uint2((unsigned)cell, __float_as_uint(dv));
        s_bin[warp][rank] = d;
      }
      __syncwarp();
      for (int j0 = 0; j0 < cnt; j0 += 8) {
        float g[2][NREG];
        float dvj[2];
        int bin[2];
        bool on[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int j = j0 + 4 * q + grp;
          on[q] = j < cnt;
          if (on[q]) {
            const uint2 e = s_list[warp][j];
            bin[q] = s_bin[warp][j];
            dvj[q] = __uint_as_float(e.y);
            g8_load_row<NV2, false>(gbase + (size_t)e.x * (C * 4), l8, g[q]);
          }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float dot = 0.f;
          if (on[q]) {
#pragma unroll
            for (int r = 0; r < NREG; ++r) {
              dot = fmaf(g[q][r], cx[r], dot);
              gacc[r] = fmaf(dvj[q], g[q][r], gacc[r]);
            }
          }
          dot += __shfl_xor_sync(0xffffffffu, dot, 4);
          dot += __shfl_xor_sync(0xffffffffu, dot, 2);
          dot += __shfl_xor_sync(0xffffffffu, dot, 1);
          if (on[q] && l8 == 0) grad_depth[img_base + (int64_t)bin[q] * HW + hw] = dot;
        }
      }
      __syncwarp();
    }
    // combine the 4 groups' partial context gradients in a fixed order
#pragma unroll
    for (int r = 0; r < NREG; ++r) {
      gacc[r] += __shfl_xor_sync(0xffffffffu, gacc[r], 8);
      gacc[r] += __shfl_xor_sync(0xffffffffu, gacc[r], 16);
    }
    if (grp == 0) {
      float *op = grad_ctx_nchw + (int64_t)bn * C * HW + hw;
#pragma unroll
      for (int r = 0; r < NREG; ++r) op[(int64_t)g8_channel<NV2>(r, l8) * HW] = gacc[r];
    }
  }
}

inline bool g8_supported(int C) {
  return C == 32 || C == 64 || C == 80 || C == 96 || C == 128;
}

#define BEVPOOL_G8_DISPATCH(C, CALL)                    \
  switch (C) {                                          \
    case 32: { constexpr int NV2 = 2; CALL; break; }    \
    case 64: { constexpr int NV2 = 4; CALL; break; }    \
    case 80: { constexpr int NV2 = 5; CALL; break; }    \
    case 96: { constexpr int NV2 = 6; CALL; break; }    \
    case 128: { constexpr int NV2 = 8; CALL; break; }   \
    default: return BEVPOOL_E_CHANNELS;                 \
  }

}  // namespace bevpool
