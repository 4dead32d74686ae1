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
// Warp-autonomous: every warp owns 32 consecutive BEV cells and never synchronises with the rest
// of its CTA, so the SM's warp scheduler hides the dependent cell_start -> sorted ids -> depth ->
// context-row latencies across independent warps.
//  * empty cells are zero-filled with coalesced 16-byte stores (every output element is written
//    exactly once, no memset);
//  * the points of the 32 cells are contiguous in the plan's sorted list; the range is cut into 4
//    equal slices (point granularity, NOT cell granularity: near-camera cells hold 100x the median)
//    and each 8-lane group reduces one slice, 8 list entries per batch (ids + depth prefetched one
//    batch ahead, 4 context rows in flight);
//  * a cell that straddles a slice boundary is finished by a fixed-order in-warp fix-up
//    (head partial of slice g+1 added to the open tail of slice g), so the result is a pure
//    function of the plan: bit-stable run to run.
constexpr int kG8FwdWarps = 4;
constexpr int kG8FwdCellsPerWarp = 32;
constexpr int kG8FwdCellsPerCta = kG8FwdCellsPerWarp * kG8FwdWarps;

struct FastDiv {          // exact n / d for 0 <= n < 2^31 (round-up magic, 64-bit product)
  uint32_t mul, shift, div;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  uint32_t s = 0;
  while ((1ull << s) < d) ++s;
  f.shift = 31 + s;
  f.mul = (uint32_t)((1ull << f.shift) / d + 1ull);
  f.div = d;
  return f;
}
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, const FastDiv &f) {
  return (uint32_t)(((uint64_t)n * f.mul) >> f.shift);
}

template <int NV2, bool kFused>
__global__ void __launch_bounds__(kG8FwdWarps * 32)
pool_forward_g8_kernel(const int32_t *__restrict__ cell_start, const int32_t *__restrict__ sorted_ids,
                       const float *__restrict__ rows, const float *__restrict__ depth,
                       float *__restrict__ out, int64_t total_cells, FastDiv div_dhw, FastDiv div_hw,
                       int row_pitch) {
  constexpr int C = 16 * NV2, NREG = 2 * NV2, C4 = C / 4, U = 4;
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31, l8 = lane & 7, grp = lane >> 3, warp = threadIdx.x >> 5;
  const int64_t cell0 = ((int64_t)blockIdx.x * kG8FwdWarps + warp) * kG8FwdCellsPerWarp;
  if (cell0 >= total_cells) return;
  const int ncell = (int)min((int64_t)kG8FwdCellsPerWarp, total_cells - cell0);
  const int cs = __ldg(cell_start + cell0 + min(lane, ncell));
  const int ce = __ldg(cell_start + cell0 + min(lane + 1, ncell));
  const unsigned occ = __ballot_sync(kFull, ce > cs);
  float *out_w = out + cell0 * C;
  {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int i = lane; i < ncell * C4; i += 32)
      if (!((occ >> (i / C4)) & 1u)) stg_stream_f4(reinterpret_cast<float4 *>(out_w) + i, z);
  }
  if (occ == 0u) return;
  const int wstart = __shfl_sync(kFull, cs, 0), wend = __shfl_sync(kFull, ce, 31);
  const int n = wend - wstart;

  // my slice of the warp's point range, and the cell its first point belongs to
  int lo = wstart, hi = wstart, cur = 0;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int glo = wstart + (int)(((int64_t)n * g) >> 2);
    const unsigned m = __ballot_sync(kFull, ce > cs && cs <= glo && glo < ce);
    if (g == grp) {
      lo = glo;
      hi = wstart + (int)(((int64_t)n * (g + 1)) >> 2);
      cur = m ? __ffs(m) - 1 : 31;
    }
  }
  int cur_end = __shfl_sync(kFull, ce, cur);
  const int cur_begin = __shfl_sync(kFull, cs, cur);     // (all lanes: no short-circuit around a shuffle)
  const bool starts_mid = hi > lo && lo > cur_begin;
  bool in_head = starts_mid, had_boundary = false, open = false;

  // head partial of a slice that starts inside a cell: parked in shared memory (rarely touched)
  __shared__ float s_head[kG8FwdWarps][4][NREG][8];
  float acc[NREG];
#pragma unroll
  for (int r = 0; r < NREG; ++r) acc[r] = 0.f;
  const char *rows_b = reinterpret_cast<const char *>(rows);
  const size_t pitch_b = (size_t)row_pitch * 4;

  auto load_entry = [&](int idx, unsigned &erow, float &ed) {
    erow = 0u;
    ed = 0.f;
    if (idx < hi) {
      const unsigned gp = (unsigned)ldg_stream_i32(sorted_ids + idx);
      if (kFused) {
        const unsigned img = fastdiv(gp, div_dhw);                  // b*N + n
        const unsigned rem = gp - fastdiv(gp, div_hw) * div_hw.div; // h*W + w
        erow = img * div_hw.div + rem;                             // context row of the point's pixel
        ed = ldg_stream_f32(depth + gp);
      } else {
        erow = gp;
      }
    }
  };

  unsigned e_row, n_row;
  float e_d, n_d;
  load_entry(lo + l8, e_row, e_d);
  const int maxlen = (n + 3) >> 2;
  for (int t = 0; t < maxlen; t += 8) {
    const int pos = lo + t;
    load_entry(pos + 8 + l8, n_row, n_d);     // next batch: in flight while this one is reduced
#pragma unroll
    for (int j0 = 0; j0 < 8; j0 += U) {
      float v[U][NREG];
      float d[U];
      bool on[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned row = __shfl_sync(kFull, e_row, j0 + u, 8);
        d[u] = __shfl_sync(kFull, e_d, j0 + u, 8);
        on[u] = pos + j0 + u < hi;
        if (on[u]) g8_load_row<NV2, !kFused>(rows_b + (size_t)row * pitch_b, l8, v[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (on[u]) {
#pragma unroll
          for (int r = 0; r < NREG; ++r)
            acc[r] = kFused ? __fadd_rn(acc[r], __fmul_rn(d[u], v[u][r])) : acc[r] + v[u][r];
          open = true;
        }
        const bool flush = on[u] && (pos + j0 + u + 1 == cur_end);
        if (__any_sync(kFull, flush)) {
          if (flush) {
            if (in_head) {
#pragma unroll
              for (int r = 0; r < NREG; ++r) s_head[warp][grp][r][l8] = acc[r];
              in_head = false;
            } else {
              g8_store_row<NV2>(reinterpret_cast<char *>(out_w + (size_t)cur * C), l8, acc);
            }
#pragma unroll
            for (int r = 0; r < NREG; ++r) acc[r] = 0.f;
            had_boundary = true;
            open = false;
            const unsigned rest = cur < 31 ? occ & ~((2u << cur) - 1u) : 0u;
            cur = rest ? __ffs(rest) - 1 : 31;
          }
          const int nce = __shfl_sync(kFull, ce, cur);
          if (flush) cur_end = nce;
        }
      }
    }
    e_row = n_row;
    e_d = n_d;
  }
  if (in_head) {          // the whole slice lies inside one cell that started in an earlier slice
#pragma unroll
    for (int r = 0; r < NREG; ++r) { s_head[warp][grp][r][l8] = acc[r]; acc[r] = 0.f; }
    open = false;
  }
  __syncwarp();

  // fixed-order fix-up of cells that straddle slice boundaries
  float carry[NREG];
#pragma unroll
  for (int r = 0; r < NREG; ++r) carry[r] = 0.f;
  int carry_cell = 0;
  bool carry_open = false;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int src = g * 8;
    const bool g_nonempty = __shfl_sync(kFull, (int)(hi > lo), src) != 0;
    const bool g_mid = __shfl_sync(kFull, (int)starts_mid, src) != 0;
    const bool g_bound = __shfl_sync(kFull, (int)had_boundary, src) != 0;
    const bool g_open = __shfl_sync(kFull, (int)open, src) != 0;
    const int g_cur = __shfl_sync(kFull, cur, src);
    if (!g_nonempty) continue;
    if (g_mid) {
#pragma unroll
      for (int r = 0; r < NREG; ++r) carry[r] += s_head[warp][g][r][l8];
      if (g_bound) {
        if (grp == 0) g8_store_row<NV2>(reinterpret_cast<char *>(out_w + (size_t)carry_cell * C), l8, carry);
        carry_open = false;
      }
    }
    if (g_open) {
#pragma unroll
      for (int r = 0; r < NREG; ++r) carry[r] = __shfl_sync(kFull, acc[r], src + l8);
      carry_cell = g_cur;
      carry_open = true;
    }
  }
  if (carry_open && grp == 0) g8_store_row<NV2>(reinterpret_cast<char *>(out_w + (size_t)carry_cell * C), l8, carry);
}

// ---- fused backward ---------------------------------------------------------------------------
// Pixel-centric, no atomics, no sort:
//   grad_depth[d, pix]  = <grad_out[cell(d, pix), :], context[pix, :]>
//   grad_context[pix,:] = sum_d depth[d, pix] * grad_out[cell(d, pix), :]
// CTA tile = 4 image columns x (4*HG) image rows of one camera image; warp = one column x 4 rows,
// 8-lane group = one pixel, which it follows through all depth bins (its context row and its
// context-gradient accumulator live in registers, summed in ascending-d order).
//  * The 4 rows of a warp project to the same BEV cell for a level camera, so the 4 groups ask for
//    the SAME gradient row: one L1 wavefront per 128 B instead of four, and the other row-warps of
//    the CTA hit the line in L1.  Nothing relies on that: rows with different cells just cost more
//    wavefronts.
//  * cell_of_point / depth / grad_depth are (d, h, w)-major, i.e. strided by H*W along a pixel's
//    ray.  They are moved 32 depth bins at a time as 16-byte (4-column) segments through shared
//    memory, one segment per thread, and the next chunk's segments are prefetched into registers
//    while the current chunk is reduced.
constexpr int kBwTW = 4;     // image columns per CTA = one 16-byte segment
constexpr int kBwDC = 32;    // depth bins per staged chunk
constexpr int kBwHG = 2;     // row groups (of 4 rows) per CTA
constexpr int kBwU = 2;      // depth bins in flight per warp

template <int NV2, int HG, bool kVec>
__global__ void __launch_bounds__(128 * HG)
fused_backward_g8_kernel(const int32_t *__restrict__ cell_of_point, const float *__restrict__ grad_rows,
                         const float *__restrict__ depth, const float *__restrict__ ctx_nchw,
                         float *__restrict__ grad_depth, float *__restrict__ grad_ctx_nchw, int num_cams,
                         int D, int H, int W, int64_t cells_per_sample, int tiles_h, int tiles_w) {
  constexpr int C = 16 * NV2, NREG = 2 * NV2, TH = 4 * HG, LD = kBwDC + 1;
  __shared__ uint2 s_cd[kBwTW][TH][LD];    // (cell, depth bits) of the chunk, [column][row][bin]
  __shared__ float s_res[kBwTW][TH][LD];   // grad_depth of the chunk
  const int tid = threadIdx.x, lane = tid & 31, l8 = lane & 7, grp = lane >> 3, warp = tid >> 5;
  int bid = blockIdx.x;
  const int th = bid % tiles_h; bid /= tiles_h;
  const int tw = bid % tiles_w;
  const int bn = bid / tiles_w;
  const int h0 = th * TH, w0 = tw * kBwTW;
  const int HW = H * W;
  const int64_t img_base = (int64_t)bn * D * HW;
  const char *gbase = reinterpret_cast<const char *>(grad_rows + (int64_t)(bn / num_cams) * cells_per_sample * C);

  // staging role: thread = (bin sd of the chunk, row sh of the tile), 4 columns
  const int sd = lane, sh = warp;
  const bool srow = h0 + sh < H;
  const int64_t sbase = img_base + (int64_t)(h0 + sh) * W + w0;
  // reducing role: warp = (column wl, row group hg); group = row hl
  const int wl = warp & 3, hl = 4 * (warp >> 2) + grp;
  const bool pix_ok = (w0 + wl < W) && (h0 + hl < H);
  const int hw = (h0 + hl) * W + w0 + wl;

  float cx[NREG], gacc[NREG];
  {
    const float *cp = ctx_nchw + (int64_t)bn * C * HW + hw;
#pragma unroll
    for (int r = 0; r < NREG; ++r) {
      cx[r] = pix_ok ? __ldg(cp + (int64_t)g8_channel<NV2>(r, l8) * HW) : 0.f;
      gacc[r] = 0.f;
    }
  }

  int4 pc = make_int4(-1, -1, -1, -1);
  float4 pd = make_float4(0.f, 0.f, 0.f, 0.f);
  auto prefetch = [&](int d0) {
    const int d = d0 + sd;
    pc = make_int4(-1, -1, -1, -1);
    if (srow && d < D) {
      const int64_t gp = sbase + (int64_t)d * HW;
      if (kVec) {
        pc = ldg_stream_i4(reinterpret_cast<const int4 *>(cell_of_point + gp));
        pd = ldg_stream_f4(reinterpret_cast<const float4 *>(depth + gp));
      } else {
        if (w0 + 0 < W) { pc.x = __ldg(cell_of_point + gp + 0); pd.x = __ldg(depth + gp + 0); }
        if (w0 + 1 < W) { pc.y = __ldg(cell_of_point + gp + 1); pd.y = __ldg(depth + gp + 1); }
        if (w0 + 2 < W) { pc.z = __ldg(cell_of_point + gp + 2); pd.z = __ldg(depth + gp + 2); }
        if (w0 + 3 < W) { pc.w = __ldg(cell_of_point + gp + 3); pd.w = __ldg(depth + gp + 3); }
      }
    }
  };
  prefetch(0);

  for (int d0 = 0; d0 < D; d0 += kBwDC) {
    s_cd[0][sh][sd] = make_uint2((unsigned)pc.x, __float_as_uint(pd.x));
    s_cd[1][sh][sd] = make_uint2((unsigned)pc.y, __float_as_uint(pd.y));
    s_cd[2][sh][sd] = make_uint2((unsigned)pc.z, __float_as_uint(pd.z));
    s_cd[3][sh][sd] = make_uint2((unsigned)pc.w, __float_as_uint(pd.w));
    s_res[0][sh][sd] = 0.f; s_res[1][sh][sd] = 0.f; s_res[2][sh][sd] = 0.f; s_res[3][sh][sd] = 0.f;
    __syncthreads();
    if (d0 + kBwDC < D) prefetch(d0 + kBwDC);     // in flight while this chunk is reduced

    // bins of the chunk kept by at least one of the warp's 4 rows
    unsigned dmask = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = (int)s_cd[wl][hl][l8 + 8 * k].x;
      unsigned m = __ballot_sync(0xffffffffu, c >= 0);
      m |= m >> 16;
      m |= m >> 8;
      dmask |= (m & 0xffu) << (8 * k);
    }
    while (dmask) {
      int dq[kBwU];
      uint2 e[kBwU];
      bool on[kBwU];
      float g[kBwU][NREG];
#pragma unroll
      for (int q = 0; q < kBwU; ++q) {
        dq[q] = dmask ? __ffs(dmask) - 1 : 0;
        on[q] = dmask != 0u;
        dmask &= dmask - 1u;
        e[q] = s_cd[wl][hl][dq[q]];
        on[q] = on[q] && (int)e[q].x >= 0;
        if (on[q]) g8_load_row<NV2, false>(gbase + (size_t)e[q].x * (C * 4), l8, g[q]);
      }
#pragma unroll
      for (int q = 0; q < kBwU; ++q) {
        float dot = 0.f;
        if (on[q]) {
          const float dv = __uint_as_float(e[q].y);
#pragma unroll
          for (int r = 0; r < NREG; ++r) {
            dot = fmaf(g[q][r], cx[r], dot);
            gacc[r] = fmaf(dv, g[q][r], gacc[r]);
          }
        }
        dot += __shfl_xor_sync(0xffffffffu, dot, 4);
        dot += __shfl_xor_sync(0xffffffffu, dot, 2);
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        if (on[q] && l8 == 0) s_res[wl][hl][dq[q]] = dot;
      }
    }
    __syncthreads();
    const int d = d0 + sd;
    if (srow && d < D) {
      const int64_t gp = sbase + (int64_t)d * HW;
      const float4 r4 = make_float4(s_res[0][sh][sd], s_res[1][sh][sd], s_res[2][sh][sd], s_res[3][sh][sd]);
      if (kVec) {
        stg_stream_f4(reinterpret_cast<float4 *>(grad_depth + gp), r4);
      } else {
        if (w0 + 0 < W) grad_depth[gp + 0] = r4.x;
        if (w0 + 1 < W) grad_depth[gp + 1] = r4.y;
        if (w0 + 2 < W) grad_depth[gp + 2] = r4.z;
        if (w0 + 3 < W) grad_depth[gp + 3] = r4.w;
      }
    }
  }
  if (pix_ok) {
    float *op = grad_ctx_nchw + (int64_t)bn * C * HW + hw;
#pragma unroll
    for (int r = 0; r < NREG; ++r) op[(int64_t)g8_channel<NV2>(r, l8) * HW] = gacc[r];
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
