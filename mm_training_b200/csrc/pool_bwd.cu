// Fused lift-splat backward, tile kernel (sm_100a, fp32, channel counts 32..96).
//
//   grad_depth[d, pix]   = < grad_out[cell(d, pix), :], context[pix, :] >
//   grad_context[pix, :] = sum_d depth[d, pix] * grad_out[cell(d, pix), :]
// (the gradient of layers/backbones/lss_fpn.py:441-464 + ops/voxel_pooling/voxel_pooling.py:58-69 of the
// reference, without the (B, N, D, H, W, C) tensor).  Pixel-centric, no sort, no atomics, every sum in
// ascending-d order (bit-stable).
//
// CTA = 4 image columns x 16 rows of one camera image, walked through all depth bins in chunks of 16.
// warp = one column x 8 rows; lane = (row, channel quarter): a pixel's context row and its
// context-gradient accumulator live in the registers of 4 lanes (C/4 channels each) for the whole ray.
//
// Why this shape (ncu of the two earlier kernels, profiles/): both were bound by instruction issue and
// shared/L1 latency, not by HBM -- 100 M warp instructions per 32 frames, 53 per group of 4 points, most
// of them loop control, addressing, 3-step shuffle reductions and per-point operand loads.  Here
//  * for a level camera the rows of a (depth bin, column) pair share one BEV cell.  Per chunk the CTA
//    finds that "primary" cell of each of its 32 x 4 pairs and stages their gradient rows in shared
//    memory with coalesced, fully independent 16-byte loads; a warp then reads ONE gradient row per
//    depth bin for all its 8 pixels (5 LDS.128 per lane, 4 distinct addresses per instruction) instead
//    of one per pixel, and the dot product needs 2 shuffle steps instead of 3;
//  * points whose cell differs from the primary one (tilted cameras, random geometry) load their
//    row directly -- correct for any geometry, fast for the common one;
//  * every global read is an asynchronous copy (cp.async, 16 bytes) issued one or two chunks ahead:
//    while chunk c is reduced, the gradient rows of chunk c+1 and the (cell, depth) segments of chunk
//    c+2 are in flight (3 + 2 shared-memory stages), so the reduction never waits on HBM/L2 latency
//    after the prologue.  The previous version staged synchronously and spent 40 % of its time there
//    at 13 % issue utilisation.
#include "common.cuh"
#include "pool_g8.cuh"

#include <cstdlib>
#include <type_traits>

namespace bevpool {

constexpr int kBtTW = 4;      // image columns per CTA
constexpr int kBtTH = 16;     // image rows per CTA
constexpr int kBtDC = 16;     // depth bins per chunk
constexpr int kBtThreads = 256;
constexpr int kBtCellStages = 3, kBtRowStages = 2;

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

template <int NV2 /* C = 16 * NV2; a lane owns NV2 float4 = C / 4 channels */, bool kVec, int kMinCtas>
__global__ void __launch_bounds__(kBtThreads, kMinCtas)
fused_backward_tile_kernel(const int32_t *__restrict__ cell_of_point, const float *__restrict__ grad_rows,
                           const float *__restrict__ depth, const float *__restrict__ ctx_nhwc,
                           float *__restrict__ grad_depth, float *__restrict__ grad_ctx_nhwc, int num_cams,
                           int D, int H, int W, int64_t cells_per_sample, int tiles_h, int tiles_w) {
  pdl_wait();
  pdl_trigger();
  constexpr int C = 16 * NV2, C4 = C / 4, NQ = NV2;      // NQ float4 per lane
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int kRowFloats = kBtDC * kBtTW * C;            // one stage of gradient rows
  extern __shared__ __align__(16) unsigned char s_raw[];
  float *s_g = reinterpret_cast<float *>(s_raw);                                              // [stage][bin][column][channel]
  int4 (*s_cell)[kBtDC][kBtTH] = reinterpret_cast<int4 (*)[kBtDC][kBtTH]>(s_g + kBtRowStages * kRowFloats);   // [stage][bin][row] x 4 columns
  float4 (*s_dep)[kBtDC][kBtTH] = reinterpret_cast<float4 (*)[kBtDC][kBtTH]>(s_cell + kBtCellStages);
  float4 (*s_res)[kBtTH] = reinterpret_cast<float4 (*)[kBtTH]>(s_dep + kBtCellStages);      // grad_depth of the chunk
  int4 (*s_pc)[kBtDC] = reinterpret_cast<int4 (*)[kBtDC]>(s_res + kBtDC);                     // [stage][bin] primary cells x 4 columns
  int (*s_flag)[kBtDC] = reinterpret_cast<int (*)[kBtDC]>(s_pc + kBtRowStages);               // [stage][bin]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int bid = blockIdx.x;
  const int th = bid % tiles_h; bid /= tiles_h;
  const int tw = bid % tiles_w;
  const int bn = bid / tiles_w;
  const int h0 = th * kBtTH, w0 = tw * kBtTW;
  const int HW = H * W;
  const int64_t img_base = (int64_t)bn * D * HW;
  const float *gbase = grad_rows + (int64_t)(bn / num_cams) * cells_per_sample * C;
  const int nchunks = (D + kBtDC - 1) / kBtDC;

  // staging role: thread = (bin sd, row sh) of a chunk, 4 columns
  const int sh = tid & 15, sd = tid >> 4;
  const bool srow = h0 + sh < H;
  const int64_t sbase = img_base + (int64_t)(h0 + sh) * W + w0;
  // reducing role: warp = (column wl, rows 8*hh .. 8*hh+7); lane = (row r8, channel quarter q)
  const int wl = warp & 3, hh = warp >> 2, r8 = lane >> 2, q = lane & 3;
  const int hl = 8 * hh + r8;
  const bool pix_ok = (w0 + wl < W) && (h0 + hl < H);
  const int hw = (h0 + hl) * W + w0 + wl;
  const int my_off = hl * 4 + wl;                                 // scalar index of (row, column) inside a bin

  float4 cx[NQ], gacc[NQ];
#pragma unroll
  for (int j = 0; j < NQ; ++j) cx[j] = gacc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (pix_ok) {
    const float4 *src = reinterpret_cast<const float4 *>(ctx_nhwc + ((int64_t)bn * HW + hw) * C) + q * NQ;
#pragma unroll
    for (int j = 0; j < NQ; ++j) cx[j] = ldg_stream_f4(src + j);
  }
  s_res[sd][sh] = make_float4(0.f, 0.f, 0.f, 0.f);

  // ---- pipeline stages --------------------------------------------------------------------------
  auto issue_cells = [&](int c) {            // (cell, depth) segments of chunk c -> stage c % 3
    if (c < nchunks) {
      const int st = c % kBtCellStages, d = c * kBtDC + sd;
      int4 *dc = &s_cell[st][sd][sh];
      float4 *dd = &s_dep[st][sd][sh];
      if (srow && d < D) {
        const int64_t gp = sbase + (int64_t)d * HW;
        if (kVec) {
          cp_async16(dc, cell_of_point + gp);
          cp_async16(dd, depth + gp);
        } else {
          int4 pc = make_int4(-1, -1, -1, -1);
          float4 pd = make_float4(0.f, 0.f, 0.f, 0.f);
          if (w0 + 0 < W) { pc.x = __ldg(cell_of_point + gp + 0); pd.x = __ldg(depth + gp + 0); }
          if (w0 + 1 < W) { pc.y = __ldg(cell_of_point + gp + 1); pd.y = __ldg(depth + gp + 1); }
          if (w0 + 2 < W) { pc.z = __ldg(cell_of_point + gp + 2); pd.z = __ldg(depth + gp + 2); }
          if (w0 + 3 < W) { pc.w = __ldg(cell_of_point + gp + 3); pd.w = __ldg(depth + gp + 3); }
          *dc = pc;
          *dd = pd;
        }
      } else {
        *dc = make_int4(-1, -1, -1, -1);
        *dd = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    cp_async_commit();
  };
  // primary cell of every (bin, column) of chunk c (+ per-half "any row kept" / "some row elsewhere" flags);
  // returns whether this thread saw a kept point.  The 16 lanes of a half-warp hold the 16 rows of a bin.
  auto primary = [&](int c) -> int {
    if (c >= nchunks) return 0;
    const int4 pc = s_cell[c % kBtCellStages][sd][sh];
    const int half = lane & 16;
    int flags = 0;     // bit 2*col + half: some row kept; bit 8 + 2*col + half: some kept row is NOT in the primary cell
    auto prim = [&](int cv, int col) -> int {
      const unsigned m = (__ballot_sync(kFull, cv >= 0) >> half) & 0xffffu;
      const int v = __shfl_sync(kFull, cv, half + (m ? __ffs(m) - 1 : 0));
      const unsigned x = (__ballot_sync(kFull, cv >= 0 && cv != v) >> half) & 0xffffu;
      flags |= (((m & 0xffu) ? 1 : 0) | ((m >> 8) ? 2 : 0) | ((x & 0xffu) ? 256 : 0) | ((x >> 8) ? 512 : 0)) << (2 * col);
      return m ? v : -1;
    };
    const int4 pp = make_int4(prim(pc.x, 0), prim(pc.y, 1), prim(pc.z, 2), prim(pc.w, 3));
    if (sh == 0) {
      s_pc[c & 1][sd] = pp;
      s_flag[c & 1][sd] = flags;
    }
    return flags & 0xff;
  };
  auto issue_rows = [&](int c, int live) {   // gradient rows of chunk c's primary cells -> stage c & 1
    if (c < nchunks && live) {
      const int *pcs = reinterpret_cast<const int *>(&s_pc[c & 1][0]);
      float4 *dst = reinterpret_cast<float4 *>(s_g + (c & 1) * kRowFloats);
#pragma unroll
      for (int i = tid; i < kBtDC * kBtTW * C4; i += kBtThreads) {
        const int row = i / C4, v = i - row * C4;
        const int cell = pcs[row];
        if (cell >= 0) cp_async16(dst + i, reinterpret_cast<const float4 *>(gbase + (int64_t)cell * C) + v);
      }
    }
    cp_async_commit();
  };

  // ---- prologue: cells(0), cells(1) in flight; primary(0); rows(0) in flight
  issue_cells(0);
  issue_cells(1);
  cp_async_wait_1();
  __syncthreads();
  int live_cur = __syncthreads_or(primary(0));
  issue_rows(0, live_cur);

  for (int c = 0; c < nchunks; ++c) {
    // in flight here: cells(c+1), rows(c)
    cp_async_wait_all();
    __syncthreads();
    const int live_next = __syncthreads_or(primary(c + 1));
    issue_rows(c + 1, live_next);
    issue_cells(c + 2);

    if (live_cur) {
      // this lane's (row, column) scalars of bin 0 and the gradient-row quarter of (bin 0, column wl); a bin
      // further on is one multiply-add away (kBtTH * 4 scalars, kBtTW * C floats per bin)
      const int *cell_p = reinterpret_cast<const int *>(&s_cell[c % kBtCellStages][0][0]) + my_off;
      const float *dep_p = reinterpret_cast<const float *>(&s_dep[c % kBtCellStages][0][0]) + my_off;
      const float4 *g_p = reinterpret_cast<const float4 *>(s_g + (c & 1) * kRowFloats + wl * C) + q * NQ;
      const int *pc_p = reinterpret_cast<const int *>(&s_pc[c & 1][0]) + wl;
      float *res_p = reinterpret_cast<float *>(s_res) + my_off;
      // bins of the chunk kept by at least one of the warp's 8 rows (lane = bin), and those that need
      // the per-lane path because some row left the primary cell
      const int fl = lane < kBtDC ? s_flag[c & 1][lane] >> (2 * wl + hh) : 0;
      unsigned dmask = __ballot_sync(kFull, fl & 1);
      const unsigned smask = __ballot_sync(kFull, (fl >> 8) & 1);
      // Software-pipelined walk over the kept bins: the scalars of bin n+1 are fetched and the 4-lane
      // reduction + store of bin n-1 are issued while the gradient row of bin n is in flight, so a warp
      // (in-order issue) is not parked on the shuffle tail before it can start the next loads.  The body
      // is instantiated twice with the roles of the two scalar sets swapped (no register shuffling), and
      // once per chunk flavour: kFast = every kept row of every bin lies in its primary cell.
      // carried state: `pend` = the next kept bin (its find-first-set runs on the slow XU pipe, so it is
      // computed one iteration before it is used); prev_h = the previous bin's dot product after the first
      // of its two shuffle steps (issued at the end of its body, consumed in the middle of the next one).
      int pend = -1;
      auto advance = [&]() {
        pend = dmask ? __ffs(dmask) - 1 : -1;
        dmask &= dmask - 1u;
      };
      advance();
      auto fetch = [&](int &dq_, int &cell_, float &dv_) {
        dq_ = pend;
        if (pend >= 0) {
          cell_ = cell_p[pend * (kBtTH * 4)];
          dv_ = dep_p[pend * (kBtTH * 4)];
        }
        advance();
      };
      float prev_h = 0.f;
      int prev_dst = -1;                                      // bin whose dot product is still to be stored, -1: none
      auto body = [&](auto fast, int dq, int cell, float dv, int &ndq, int &ncell, float &ndv) {
        const bool on = cell >= 0;
        float4 g[NQ];
        bool staged = true;
        if (!decltype(fast)::value) {
          if ((smask >> dq) & 1u) staged = cell == pc_p[dq * 4] || !on;
        }
        if (staged) {
          const float4 *gs = g_p + dq * (kBtTW * C4);
#pragma unroll
          for (int j = 0; j < NQ; ++j) g[j] = gs[j];
        } else {
          const float4 *gg = reinterpret_cast<const float4 *>(gbase + (int64_t)cell * C) + q * NQ;
#pragma unroll
          for (int j = 0; j < NQ; ++j) g[j] = __ldg(gg + j);
        }
        fetch(ndq, ncell, ndv);
        // second reduction step of the previous bin: in flight behind this bin's multiply-adds
        const float prev_s = __shfl_xor_sync(kFull, prev_h, 1);
        // this bin
        float2 dot_a = make_float2(0.f, 0.f), dot_b = make_float2(0.f, 0.f);
        if (on) {
          const float2 dv2 = make_float2(dv, dv);
#pragma unroll
          for (int j = 0; j < NQ; ++j) {
            dot_a = __ffma2_rn(make_float2(g[j].x, g[j].y), make_float2(cx[j].x, cx[j].y), dot_a);
            dot_b = __ffma2_rn(make_float2(g[j].z, g[j].w), make_float2(cx[j].z, cx[j].w), dot_b);
            const float2 t0 = __ffma2_rn(dv2, make_float2(g[j].x, g[j].y), make_float2(gacc[j].x, gacc[j].y));
            const float2 t1 = __ffma2_rn(dv2, make_float2(g[j].z, g[j].w), make_float2(gacc[j].z, gacc[j].w));
            gacc[j] = make_float4(t0.x, t0.y, t1.x, t1.y);
          }
        }
        if (prev_dst >= 0) res_p[prev_dst * (kBtTH * 4)] = prev_h + prev_s;
        const float dot = (dot_a.x + dot_a.y) + (dot_b.x + dot_b.y);
        prev_h = dot + __shfl_xor_sync(kFull, dot, 2);         // first step: in flight across the loop back-edge
        prev_dst = (on && q == 0) ? dq : -1;
      };
      auto walk = [&](auto fast) {
        int dq0 = -1, cell0 = -1, dq1 = -1, cell1 = -1;
        float dv0 = 0.f, dv1 = 0.f;
        fetch(dq0, cell0, dv0);
        while (dq0 >= 0) {                                    // warp-uniform
          body(fast, dq0, cell0, dv0, dq1, cell1, dv1);
          if (dq1 < 0) break;
          body(fast, dq1, cell1, dv1, dq0, cell0, dv0);
        }
      };
      if (smask == 0u) walk(std::true_type{}); else walk(std::false_type{});
      const float last_s = __shfl_xor_sync(kFull, prev_h, 1);
      if (prev_dst >= 0) res_p[prev_dst * (kBtTH * 4)] = prev_h + last_s;
    }
    __syncthreads();
    // ---- grad_depth of the chunk: one 16-byte segment per (bin, row); the entry is this thread's own
    {
      const int d = c * kBtDC + sd;
      if (srow && d < D) {
        const int64_t gp = sbase + (int64_t)d * HW;
        const float4 r4 = s_res[sd][sh];
        if (kVec) {
          stg_stream_f4(reinterpret_cast<float4 *>(grad_depth + gp), r4);
        } else {
          if (w0 + 0 < W) grad_depth[gp + 0] = r4.x;
          if (w0 + 1 < W) grad_depth[gp + 1] = r4.y;
          if (w0 + 2 < W) grad_depth[gp + 2] = r4.z;
          if (w0 + 3 < W) grad_depth[gp + 3] = r4.w;
        }
      }
      if (live_cur) s_res[sd][sh] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    live_cur = live_next;
  }
  cp_async_wait_all();
  if (pix_ok) {
    float4 *dst = reinterpret_cast<float4 *>(grad_ctx_nhwc + ((int64_t)bn * HW + hw) * C) + q * NQ;
#pragma unroll
    for (int j = 0; j < NQ; ++j) dst[j] = gacc[j];
  }
}

inline bool bt_supported(int C) { return C == 32 || C == 64 || C == 80 || C == 96; }

template <int NV2, bool kVec, int kMinCtas>
static int launch_bt(const int32_t *cell_of_point, const float *grad_rows, const float *depth,
                     const float *ctx_nhwc, float *grad_depth, float *grad_ctx_nhwc, int num_cams, int D, int H,
                     int W, int64_t cells_per_sample, int64_t ctas, int tiles_h, int tiles_w, cudaStream_t s) {
  constexpr int C = 16 * NV2;
  const size_t smem = (size_t)kBtRowStages * kBtDC * kBtTW * C * 4 + (size_t)kBtCellStages * kBtDC * kBtTH * 32 +
                      (size_t)kBtDC * kBtTH * 16 + (size_t)kBtRowStages * kBtDC * 20;
  if (smem > 48 * 1024)
    BEVPOOL_RETURN_IF_CUDA(cudaFuncSetAttribute(fused_backward_tile_kernel<NV2, kVec, kMinCtas>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BEVPOOL_RETURN_IF_CUDA(launch_pdl(fused_backward_tile_kernel<NV2, kVec, kMinCtas>, dim3((unsigned)ctas), dim3(kBtThreads),
                                    smem, s, cell_of_point, grad_rows, depth, ctx_nhwc, grad_depth, grad_ctx_nhwc, num_cams, D,
                                    H, W, cells_per_sample, tiles_h, tiles_w));
  BEVPOOL_LAUNCH_CHECK();
  return BEVPOOL_OK;
}

#define BEVPOOL_BT_DISPATCH(C, CALL)                    \
  switch (C) {                                          \
    case 32: { constexpr int NV2 = 2; CALL; break; }    \
    case 64: { constexpr int NV2 = 4; CALL; break; }    \
    case 80: { constexpr int NV2 = 5; CALL; break; }    \
    case 96: { constexpr int NV2 = 6; CALL; break; }    \
    default: return BEVPOOL_E_CHANNELS;                 \
  }

int launch_fused_backward_tile(const int32_t *cell_of_point, const float *grad_rows, const float *depth,
                               const float *ctx_nhwc, float *grad_depth, float *grad_ctx_nhwc, int batch,
                               int num_cams, int D, int H, int W, int C, int64_t cells_per_sample,
                               cudaStream_t s) {
  const int64_t tiles_h = ceil_div64(H, kBtTH), tiles_w = ceil_div64(W, kBtTW);
  const int64_t ctas = (int64_t)batch * num_cams * tiles_h * tiles_w;
  if (ctas >= (int64_t)INT32_MAX) return BEVPOOL_E_RANGE;
  const bool vec = (W % 4 == 0) && aligned16(depth) && aligned16(grad_depth) && aligned16(cell_of_point);
  int rc = BEVPOOL_OK;
  static const int occ = [] { const char *e = std::getenv("BEVPOOL_BW_OCC"); return e && e[0] == '3' ? 3 : 2; }();   // 2 resident CTAs: no spills (measured faster than 3 with spills)
#define BEVPOOL_BT_ARGS cell_of_point, grad_rows, depth, ctx_nhwc, grad_depth, grad_ctx_nhwc, num_cams, D, H, W, \
                        cells_per_sample, ctas, (int)tiles_h, (int)tiles_w, s
  if (vec && occ == 3) { BEVPOOL_BT_DISPATCH(C, (rc = launch_bt<NV2, true, 3>(BEVPOOL_BT_ARGS))); }
  else if (vec) { BEVPOOL_BT_DISPATCH(C, (rc = launch_bt<NV2, true, 2>(BEVPOOL_BT_ARGS))); }
  else { BEVPOOL_BT_DISPATCH(C, (rc = launch_bt<NV2, false, 3>(BEVPOOL_BT_ARGS))); }
#undef BEVPOOL_BT_ARGS
  return rc;
}

bool fused_backward_tile_supported(int C) { return bt_supported(C); }

}  // namespace bevpool
