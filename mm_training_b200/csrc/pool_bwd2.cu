// Fused lift-splat backward, column kernel (sm_100a, fp32, g8 channel counts).
//
//   grad_depth[d, pix]   = < grad_out[cell(d, pix), :], context[pix, :] >
//   grad_context[pix, :] = sum_d depth[d, pix] * grad_out[cell(d, pix), :]
// (the gradient of layers/backbones/lss_fpn.py:441-464 + ops/voxel_pooling/voxel_pooling.py:58-69 of the
// reference, without the (B, N, D, H, W, C) tensor).  Pixel-centric, no sort, no atomics, every sum in
// ascending-d order (bit-stable).
//
// Why a second kernel (ncu of fused_backward_tile_kernel, profiles/r01zz_ncu_kernels.txt): that kernel is
// bound by instruction issue, not by HBM -- 81 M warp instructions per 32 frames of which 12 M are FFMA2,
// 117 registers, 24 % occupancy, 46 % issue utilisation.  A lane there owns ONE pixel x C/4 channels, so
// every gradient-row quarter it reads from shared memory (5 LDS.128) feeds only 20 FFMA2, and the per-bin
// bookkeeping (scalar fetches, find-first-set walk, two-step shuffle reduction) is paid per 8 pixels.
// Here
//  * warp = one image column x all 16 rows; lane = (4 consecutive rows, channel eighth): a gradient-row
//    eighth (2 LDS.128 + 1 LDS.64 at C = 80, the same addresses for the four row groups = one wavefront
//    each) feeds 40 FFMA2, and the per-bin bookkeeping is paid per 16 pixels;
//  * the four dot products a lane holds are reduced over the 8 lanes of its group with a transposing
//    butterfly: 4 shuffles per (bin, column) instead of 2 x 2 per half column;
//  * the depth values arrive transposed ([bin][column][row]: one LDS.128 per lane) together with a 16-bit
//    "rows in the primary cell" mask per (bin, column), both produced by the staging threads while they
//    look for the primary cells, so the reduction loop has no per-row scalar fetches and no bit scans;
//  * 128-thread CTAs, 3 per SM: latency is hidden by the 40 independent FFMA2 per bin rather than by
//    occupancy;
//  * context is read from the caller's NCHW tensor and the context gradient written back NCHW through
//    TMA tensor maps (cp.async.bulk.tensor, boxes of 4 columns x 16 rows x C channels): the two layout
//    passes (bevpool_transpose, 15 us each per 32 frames) are gone.
// Correct for any geometry: rows of a (bin, column) that do not fall into its primary cell (tilted
// cameras, random geometry) take a per-row path that gathers their gradient row from global memory.
#include "common.cuh"
#include "pool_g8.cuh"
#include "tma.cuh"

namespace bevpool {

constexpr int kBcTW = 4;                            // image columns per CTA = warps per CTA
constexpr int kBcTH = 16;                           // image rows per CTA
constexpr int kBcDC = 16;                           // depth bins per chunk
constexpr int kBcThreads = 32 * kBcTW;
constexpr int kBcDepStride = kBcTW * kBcTH + 16;    // floats per bin of the transposed depth stage (+16: the two
                                                    // half-warps of a staging warp hit disjoint banks)

template <int NV2>
struct BcSmem {
  static constexpr int C = 16 * NV2;
  static constexpr size_t kRowFloats = (size_t)kBcDC * kBcTW * C;                  // one stage of gradient rows
  static constexpr size_t off_g = 0;                                                  // [2][bin][column][channel]; also the context / context-gradient TMA box
  static constexpr size_t off_cell = off_g + 2 * kRowFloats * 4;                      // int4   [2][bin][row]   (4 columns)
  static constexpr size_t off_dep = off_cell + 2 * kBcDC * kBcTH * 16;                // float4 [2][bin][row]
  static constexpr size_t off_depT = off_dep + 2 * kBcDC * kBcTH * 16;                // float  [2][bin][kBcDepStride]
  static constexpr size_t off_mask = off_depT + 2 * kBcDC * kBcDepStride * 4;         // uint32 [2][bin][column]
  static constexpr size_t off_pc = off_mask + 2 * kBcDC * kBcTW * 4;                  // int    [2][bin][column]
  static constexpr size_t off_res = off_pc + 2 * kBcDC * kBcTW * 4;                   // float4 [bin][row]
  static constexpr size_t off_bar = off_res + kBcDC * kBcTH * 16;
  static constexpr size_t bytes = off_bar + 16;
};

// transposing butterfly over the 8 lanes of a group: in = 4 partial sums per lane, out = the full sum of
// value (l8 >> 1) (both lanes of a pair hold it).  Fixed association order.
__device__ __forceinline__ float reduce4_over8(const float (&s)[4], int l8) {
  constexpr unsigned kFull = 0xffffffffu;
  const bool hi4 = (l8 & 4) != 0;
  const float t0 = (hi4 ? s[2] : s[0]) + __shfl_xor_sync(kFull, hi4 ? s[0] : s[2], 4);
  const float t1 = (hi4 ? s[3] : s[1]) + __shfl_xor_sync(kFull, hi4 ? s[1] : s[3], 4);
  const bool hi2 = (l8 & 2) != 0;
  const float u = (hi2 ? t1 : t0) + __shfl_xor_sync(kFull, hi2 ? t0 : t1, 2);
  return u + __shfl_xor_sync(kFull, u, 1);
}

template <int NV2, bool kNchw>
__global__ void __launch_bounds__(kBcThreads, (NV2 <= 5 ? 3 : 2))
fused_backward_col_kernel(const __grid_constant__ CUtensorMap ctx_map, const __grid_constant__ CUtensorMap gctx_map,
                          const int32_t *__restrict__ cell_of_point, const float *__restrict__ grad_rows,
                          const float *__restrict__ depth, const float *__restrict__ ctx_nhwc,
                          float *__restrict__ grad_depth, float *__restrict__ grad_ctx_nhwc, int num_cams, int D,
                          int H, int W, int64_t cells_per_sample, int tiles_h, int tiles_w) {
  pdl_wait();
  pdl_trigger();
  using S = BcSmem<NV2>;
  constexpr int C = S::C, C4 = C / 4, NREG = 2 * NV2;
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int kRowFloats = (int)S::kRowFloats;
  extern __shared__ __align__(128) unsigned char s_raw[];
  float *s_g = reinterpret_cast<float *>(s_raw + S::off_g);
  int4 (*s_cell)[kBcDC][kBcTH] = reinterpret_cast<int4 (*)[kBcDC][kBcTH]>(s_raw + S::off_cell);
  float4 (*s_dep)[kBcDC][kBcTH] = reinterpret_cast<float4 (*)[kBcDC][kBcTH]>(s_raw + S::off_dep);
  float *s_depT = reinterpret_cast<float *>(s_raw + S::off_depT);
  uint32_t *s_mask = reinterpret_cast<uint32_t *>(s_raw + S::off_mask);
  int *s_pc = reinterpret_cast<int *>(s_raw + S::off_pc);
  float4 (*s_res)[kBcTH] = reinterpret_cast<float4 (*)[kBcTH]>(s_raw + S::off_res);
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_raw + S::off_bar);

  const int tid = threadIdx.x, lane = tid & 31, wl = tid >> 5;      // warp = image column of the tile
  const int l8 = lane & 7, rg = lane >> 3;                          // lane = (rows 4*rg .. 4*rg+3, channel eighth)
  int bid = blockIdx.x;
  const int th = bid % tiles_h; bid /= tiles_h;
  const int tw = bid % tiles_w;
  const int bn = bid / tiles_w;
  const int h0 = th * kBcTH, w0 = tw * kBcTW;
  const int HW = H * W;
  const int64_t img_base = (int64_t)bn * D * HW;
  const float *gbase = grad_rows + (int64_t)(bn / num_cams) * cells_per_sample * C;
  const int nchunks = (D + kBcDC - 1) / kBcDC;

  // ---- context tile: TMA box [channel][row][4 columns] of the NCHW tensor (rows beyond H read as zeros)
  if (kNchw) {
    if (tid == 0) {
      mbar_init(s_bar, 1);
      fence_proxy_async();
    }
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(s_bar, (uint32_t)(kBcTW * kBcTH * C * 4));
      tma_load_4d(s_g, &ctx_map, w0, h0, 0, bn, s_bar);
    }
  }

  // staging role: a half-warp = the 16 rows of one bin; a thread stages bins sd and sd + 8
  const int sh = tid & 15, sd = tid >> 4;
  const bool srow = h0 + sh < H;
  const int64_t sbase = img_base + (int64_t)(h0 + sh) * W + w0;

  auto issue_cells = [&](int c) {            // (cell, depth) segments of chunk c -> raw stage c & 1
    if (c < nchunks) {
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int bin = sd + 8 * pass, d = c * kBcDC + bin;
        int4 *dc = &s_cell[c & 1][bin][sh];
        float4 *dd = &s_dep[c & 1][bin][sh];
        if (srow && d < D) {
          const int64_t gp = sbase + (int64_t)d * HW;
          cp_async16(dc, cell_of_point + gp);
          cp_async16(dd, depth + gp);
        } else {
          *dc = make_int4(-1, -1, -1, -1);
          *dd = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    cp_async_commit();
  };
  // primary cell of every (bin, column) of chunk c, the mask of rows lying in it (low 16 bits) and of kept
  // rows lying elsewhere (high 16 bits), and the depths transposed to [bin][column][row]
  auto prep = [&](int c) -> int {
    if (c >= nchunks) return 0;
    const int st = c & 1, half = lane & 16;
    int any = 0;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int bin = sd + 8 * pass;
      const int4 pc = s_cell[st][bin][sh];
      const float4 pd = s_dep[st][bin][sh];
      float *dT = s_depT + (st * kBcDC + bin) * kBcDepStride + sh;
      auto one = [&](int cv, float dv, int col) {
        const unsigned m = (__ballot_sync(kFull, cv >= 0) >> half) & 0xffffu;
        const int v = __shfl_sync(kFull, cv, half + (m ? __ffs(m) - 1 : 0));
        const bool fast = cv >= 0 && cv == v;
        const unsigned fm = (__ballot_sync(kFull, fast) >> half) & 0xffffu;
        dT[col * kBcTH] = fast ? dv : 0.f;
        if (sh == 0) {
          s_mask[(st * kBcDC + bin) * kBcTW + col] = fm | ((m & ~fm) << 16);
          s_pc[(st * kBcDC + bin) * kBcTW + col] = m ? v : -1;
        }
        any |= (int)m;
      };
      one(pc.x, pd.x, 0);
      one(pc.y, pd.y, 1);
      one(pc.z, pd.z, 2);
      one(pc.w, pd.w, 3);
    }
    return any;
  };
  auto issue_rows = [&](int c, int live) {   // gradient rows of chunk c's primary cells -> row stage c & 1
    if (c < nchunks && live) {
      const int *pcs = s_pc + (c & 1) * kBcDC * kBcTW;
      float4 *dst = reinterpret_cast<float4 *>(s_g + (c & 1) * kRowFloats);
#pragma unroll
      for (int i = tid; i < kBcDC * kBcTW * C4; i += kBcThreads) {
        const int row = i / C4, v = i - row * C4;
        const int cell = pcs[row];
        if (cell >= 0) cp_async16(dst + i, reinterpret_cast<const float4 *>(gbase + (int64_t)cell * C) + v);
      }
    }
    cp_async_commit();
  };

  issue_cells(0);
  issue_cells(1);

  // ---- this lane's context rows: 4 pixels (rows 4*rg + j of column wl) x NREG channels
  float cx[4][NREG], gacc[4][NREG];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int r = 0; r < NREG; ++r) cx[j][r] = gacc[j][r] = 0.f;
  if (kNchw) {
    mbar_wait(s_bar, 0);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < NREG; ++r) cx[j][r] = s_g[(g8_channel<NV2>(r, l8) * kBcTH + 4 * rg + j) * kBcTW + wl];
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int h = h0 + 4 * rg + j;
      if (h < H && w0 + wl < W)
        g8_load_row<NV2, true>(reinterpret_cast<const char *>(ctx_nhwc + ((int64_t)bn * HW + h * W + w0 + wl) * C), l8, cx[j]);
    }
  }
  for (int i = tid; i < kBcDC * kBcTH; i += kBcThreads) (&s_res[0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  cp_async_wait_1();
  __syncthreads();                           // cells(0) visible; the context box is no longer needed
  int live_cur = __syncthreads_or(prep(0));
  issue_rows(0, live_cur);

  for (int c = 0; c < nchunks; ++c) {
    // in flight here: cells(c+1), rows(c)
    cp_async_wait_all();
    __syncthreads();
    const int live_next = __syncthreads_or(prep(c + 1));
    issue_rows(c + 1, live_next);
    issue_cells(c + 2);                      // into the raw stage prep(c) consumed an iteration ago

    if (live_cur) {
      const int st = c & 1;
      const float *g_col = s_g + st * kRowFloats + wl * C;
      const float *dT = s_depT + st * kBcDC * kBcDepStride + wl * kBcTH + 4 * rg;
      const uint32_t *mk = s_mask + st * kBcDC * kBcTW + wl;
#pragma unroll 4
      for (int b = 0; b < kBcDC; ++b) {
        const uint32_t m = mk[b * kBcTW];
        if (m == 0u) continue;                                  // warp-uniform: no kept row in this (bin, column)
        const uint32_t fast = m & 0xffffu, slow = m >> 16;
        if (fast) {
          float g[NREG];
          g8_lds_row<NV2>(g_col + b * (kBcTW * C), l8, g);
          const float4 dp4 = *reinterpret_cast<const float4 *>(dT + b * kBcDepStride);
          const float dp[4] = {dp4.x, dp4.y, dp4.z, dp4.w};
          const uint32_t mine = (fast >> (4 * rg)) & 0xfu;
          // Branch-free: the dot products of rows outside the primary cell are computed and discarded, and their
          // staged depth is 0, so their accumulators receive +-0 (exact for finite gradients; see DESIGN.md 4.4 for
          // the non-finite case).  40 independent FFMA2 per bin hide the shared-memory latency of the next bin.
          float s[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 da = make_float2(0.f, 0.f), db = make_float2(0.f, 0.f);
#pragma unroll
            for (int r = 0; r < NREG; r += 4) {
              da = __ffma2_rn(make_float2(g[r], g[r + 1]), make_float2(cx[j][r], cx[j][r + 1]), da);
              if (r + 2 < NREG) db = __ffma2_rn(make_float2(g[r + 2], g[r + 3]), make_float2(cx[j][r + 2], cx[j][r + 3]), db);
            }
            axpy_row<NREG>(gacc[j], dp[j], g);
            const float2 dab = __fadd2_rn(da, db);
            s[j] = dab.x + dab.y;
          }
          const float tot = reduce4_over8(s, l8);
          const int p = l8 >> 1;
          if (!(l8 & 1) && ((mine >> p) & 1u))
            reinterpret_cast<float *>(&s_res[b][4 * rg + p])[wl] = tot;
        }
        if (slow) {
          // rows of this (bin, column) that are kept but lie outside the primary cell: gather their own row
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (((slow >> j) & 0x1111u) == 0u) continue;        // warp-uniform
            const bool p = (slow >> (4 * rg + j)) & 1u;
            float dot = 0.f;
            if (p) {
              const int64_t gp = img_base + (int64_t)(c * kBcDC + b) * HW + (int64_t)(h0 + 4 * rg + j) * W + w0 + wl;
              const int cell = __ldg(cell_of_point + gp);
              const float dv = __ldg(depth + gp);
              float g[NREG];
              g8_load_row<NV2, false>(reinterpret_cast<const char *>(gbase + (int64_t)cell * C), l8, g);
              float2 da = make_float2(0.f, 0.f);
#pragma unroll
              for (int r = 0; r < NREG; r += 2)
                da = __ffma2_rn(make_float2(g[r], g[r + 1]), make_float2(cx[j][r], cx[j][r + 1]), da);
              dot = da.x + da.y;
              axpy_row<NREG>(gacc[j], dv, g);
            }
            dot += __shfl_xor_sync(kFull, dot, 4);
            dot += __shfl_xor_sync(kFull, dot, 2);
            dot += __shfl_xor_sync(kFull, dot, 1);
            if (p && l8 == 0) reinterpret_cast<float *>(&s_res[b][4 * rg + j])[wl] = dot;
          }
        }
      }
    }
    __syncthreads();
    // ---- grad_depth of the chunk: one 16-byte segment per (bin, row)
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int bin = sd + 8 * pass, d = c * kBcDC + bin;
      if (srow && d < D) stg_stream_f4(reinterpret_cast<float4 *>(grad_depth + sbase + (int64_t)d * HW), s_res[bin][sh]);
      if (live_cur) s_res[bin][sh] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    live_cur = live_next;
  }
  cp_async_wait_all();

  // ---- context gradient
  if (kNchw) {
    __syncthreads();                         // every copy into the row stages has landed and been consumed
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < NREG; ++r) s_g[(g8_channel<NV2>(r, l8) * kBcTH + 4 * rg + j) * kBcTW + wl] = gacc[j][r];
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tma_store_4d(&gctx_map, w0, h0, 0, bn, s_g);
      tma_store_commit();
      tma_store_wait_read();
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int h = h0 + 4 * rg + j;
      if (h < H && w0 + wl < W)
        g8_store_row<NV2>(reinterpret_cast<char *>(grad_ctx_nhwc + ((int64_t)bn * HW + h * W + w0 + wl) * C), l8, gacc[j]);
    }
  }
}

template <int NV2, bool kNchw>
static int launch_bc(const CUtensorMap &ctx_map, const CUtensorMap &gctx_map, const int32_t *cell_of_point,
                     const float *grad_rows, const float *depth, const float *ctx_nhwc, float *grad_depth,
                     float *grad_ctx_nhwc, int num_cams, int D, int H, int W, int64_t cells_per_sample, int64_t ctas,
                     int tiles_h, int tiles_w, cudaStream_t s) {
  constexpr size_t smem = BcSmem<NV2>::bytes;
  static bool configured = false;
  if (!configured) {
    BEVPOOL_RETURN_IF_CUDA(cudaFuncSetAttribute(fused_backward_col_kernel<NV2, kNchw>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  BEVPOOL_RETURN_IF_CUDA(launch_pdl(fused_backward_col_kernel<NV2, kNchw>, dim3((unsigned)ctas), dim3(kBcThreads), smem, s,
                                    ctx_map, gctx_map, cell_of_point, grad_rows, depth, ctx_nhwc, grad_depth, grad_ctx_nhwc,
                                    num_cams, D, H, W, cells_per_sample, tiles_h, tiles_w));
  BEVPOOL_LAUNCH_CHECK();
  return BEVPOOL_OK;
}

bool fused_backward_col_supported(int C, int W, const void *depth, const void *grad_depth, const void *cell_of_point) {
  return g8_supported(C) && (W % 4 == 0) && aligned16(depth) && aligned16(grad_depth) && aligned16(cell_of_point);
}

// context / grad_context: NCHW (B*N, C, H, W) when `nchw`, else pixel rows (B*N, H, W, C)
int launch_fused_backward_col(const int32_t *cell_of_point, const float *grad_rows, const float *depth,
                              const float *ctx, float *grad_depth, float *grad_ctx, bool nchw, int batch, int num_cams,
                              int D, int H, int W, int C, int64_t cells_per_sample, cudaStream_t s) {
  const int64_t tiles_h = ceil_div64(H, kBcTH), tiles_w = ceil_div64(W, kBcTW);
  const int64_t ctas = (int64_t)batch * num_cams * tiles_h * tiles_w;
  if (ctas >= (int64_t)INT32_MAX) return BEVPOOL_E_RANGE;
  CUtensorMap ctx_map{}, gctx_map{};
  if (nchw) {
    const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)C, (uint64_t)batch * num_cams};
    const uint64_t strides[3] = {(uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)C * H * W * 4};
    const uint32_t box[4] = {kBcTW, kBcTH, (uint32_t)C, 1};
    int rc = make_tensor_map_f32(&ctx_map, ctx, 4, dims, strides, box);
    if (rc) return rc;
    if ((rc = make_tensor_map_f32(&gctx_map, grad_ctx, 4, dims, strides, box))) return rc;
  }
  int rc = BEVPOOL_OK;
#define BEVPOOL_BC_ARGS ctx_map, gctx_map, cell_of_point, grad_rows, depth, ctx, grad_depth, grad_ctx, num_cams, D, H, W, \
                        cells_per_sample, ctas, (int)tiles_h, (int)tiles_w, s
  if (nchw) { BEVPOOL_G8_DISPATCH(C, (rc = launch_bc<NV2, true>(BEVPOOL_BC_ARGS))); }
  else { BEVPOOL_G8_DISPATCH(C, (rc = launch_bc<NV2, false>(BEVPOOL_BC_ARGS))); }
#undef BEVPOOL_BC_ARGS
  return rc;
}

}  // namespace bevpool
