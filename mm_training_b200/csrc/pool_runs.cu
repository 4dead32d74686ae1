// Fused lift-splat forward on a RUN plan (sm_100a, fp32, g8 channel layout).
//
// Replaces the materialised outer product + scatter of layers/backbones/lss_fpn.py:441-464 and
// ops/voxel_pooling/src/voxel_pooling_forward_cuda.cu:30-34.  A run (common.cuh) is a set of
// vertically adjacent frustum points -- same image, depth bin and column -- that land in the same BEV
// cell; for a level camera that is every kept point of a (depth bin, column) pair (11 points per run at
// the aiMotive shape).  Two stages, no atomics, every sum in a fixed order (bit-stable):
//
//   A  frustum_reduce_kernel   run_rows[slot] = sum_h depth[d, h, w] * context[h, w, :]
//        pixel-centric: a CTA owns (image, 4 columns x 16 rows, 32 depth bins); its 64 context rows sit
//        in shared memory (staged once, coalesced) and every depth (x) context product reads them from
//        there -- the forward issues NO data-dependent global gather at all.
//        The context tile arrives by TMA bulk copies (cp.async.bulk + mbarrier, one elected thread), and
//        the empty BEV cells (80 % of the grid) are zero-filled by TMA bulk stores from a zeroed
//        shared-memory buffer that every CTA issues for its share of the grid before it starts
//        reducing: the DRAM-write-bound fill runs in the copy engine behind the issue-bound reduction.
//   B  pool_forward_share_kernel<kIdent>   out[cell] = sum of the cell's run rows (contiguous: the slot
//        order IS the cell order) -- the even-share segmented reduction of pool_g8.cuh over 11x fewer
//        entries; every output row written exactly once.
//
// The point-sorted kernel this replaces gathered one 320-byte context row per kept point through L2
// (1.5 GB per 32 frames, L2->SM fabric bound); here the only per-run traffic is one row written and
// read back through L2 (11x fewer rows), processed a few frames at a time so it never leaves L2.
#include "common.cuh"
#include "pool_g8.cuh"

#include <cstdlib>

namespace bevpool {

static int env_int_runs(const char *name, int dflt) {
  const char *e = std::getenv(name);
  return e && e[0] ? std::atoi(e) : dflt;
}

constexpr int kRaTW = 4;                     // image columns per CTA = one 16-byte segment
constexpr int kRaDC = 16;                    // depth bins per staged chunk
constexpr int kRaStages = 3;                 // chunks of (code, depth) in shared memory: one reduced, two in flight
constexpr int kRaThreads = 256;
constexpr int kRaZeroCells = 32;             // cells covered by the zeroed shared-memory buffer

// ---- TMA (bulk async copy) + mbarrier helpers, sm_90+ PTX ---------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, completion counted in bytes on the mbarrier (16-byte aligned addresses and size)
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global, tracked by the thread's bulk async-group
__device__ __forceinline__ void tma_store_1d(void *dst_gmem, const void *src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
// L2 eviction-priority policies: the zero-fill is written once and never read by these kernels
// (evict_first), the run rows are read back by stage B right after (evict_last)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_store_1d_hint(void *dst_gmem, const void *src_smem, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
               ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg_hint_f4(float *p, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
               ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg_hint_f2(float *p, float2 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void cp_async16_runs(void *dst_smem, const void *src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}

template <int NV2>
__device__ __forceinline__ void g8_lds_row(const float *row, int l8, float (&v)[2 * NV2]) {
  constexpr int NF4 = NV2 / 2, NF2 = NV2 & 1;
#pragma unroll
  for (int k = 0; k < NF4; ++k) {
    const float4 t = *reinterpret_cast<const float4 *>(row + 32 * k + 4 * l8);
    v[4 * k + 0] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
  }
  if (NF2) {
    const float2 t = *reinterpret_cast<const float2 *>(row + 32 * NF4 + 2 * l8);
    v[4 * NF4 + 0] = t.x; v[4 * NF4 + 1] = t.y;
  }
}

template <int NV2>
__device__ __forceinline__ void g8_store_row_hint(float *row, int l8, const float (&v)[2 * NV2], uint64_t pol) {
  constexpr int NF4 = NV2 / 2, NF2 = NV2 & 1;
#pragma unroll
  for (int k = 0; k < NF4; ++k)
    stg_hint_f4(row + 32 * k + 4 * l8, make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]), pol);
  if (NF2) stg_hint_f2(row + 32 * NF4 + 2 * l8, make_float2(v[4 * NF4], v[4 * NF4 + 1]), pol);
}

template <int NV2>
__device__ __forceinline__ void g8_store_row_plain(float *row, int l8, const float (&v)[2 * NV2]) {
  constexpr int NF4 = NV2 / 2, NF2 = NV2 & 1;
#pragma unroll
  for (int k = 0; k < NF4; ++k)
    *reinterpret_cast<float4 *>(row + 32 * k + 4 * l8) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  if (NF2) *reinterpret_cast<float2 *>(row + 32 * NF4 + 2 * l8) = make_float2(v[4 * NF4], v[4 * NF4 + 1]);
}

// one (depth bin, column) pair of an 8-lane group: codes / depths of its 16 rows (lane l8 holds rows l8 and
// l8 + 8), the group's kept-row and first-of-run masks
struct RunCol {
  int c_lo, c_hi;
  float p_lo, p_hi;
  unsigned km, hm;
};

// ---- stage A ----------------------------------------------------------------------------------
// CTA = (image, 4 columns x 16 rows, one of `d_split` depth ranges), walked in chunks of 32 depth bins.
// warp = (column wl, 8 consecutive bins per round): every 8-lane group walks the rows of TWO
// (bin, column) pairs, so one shared-memory read of a context row (one wavefront for the whole warp:
// all four groups read the same row) feeds 8 depth (x) context products.
template <int NV2>
__global__ void __launch_bounds__(kRaThreads, 4)
frustum_reduce_kernel(const int32_t *__restrict__ run_code, const float *__restrict__ depth,
                      const float *__restrict__ ctx_nhwc, float *__restrict__ run_rows,
                      const int32_t *__restrict__ cell_start, float *__restrict__ out, int img0,
                      int64_t cell_base, int64_t num_cells, int D, int H, int W, int d_split, int d_per_cta,
                      int tiles_h, int tiles_w, int64_t capacity, int vec, int fill, int hints) {
  pdl_wait();
  pdl_trigger();
  constexpr int C = 16 * NV2, NREG = 2 * NV2;
  constexpr unsigned kFull = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char s_raw[];
  float *s_ctx = reinterpret_cast<float *>(s_raw);                                  // [row][column][channel]
  int4 (*s_code)[kRaDC][kRunHB] = reinterpret_cast<int4 (*)[kRaDC][kRunHB]>(s_ctx + kRunHB * kRaTW * C);   // [stage][bin][row] x 4 columns
  float4 (*s_dep)[kRaDC][kRunHB] = reinterpret_cast<float4 (*)[kRaDC][kRunHB]>(s_code + kRaStages);
  float *s_zero = reinterpret_cast<float *>(s_dep + kRaStages);                     // kRaZeroCells * C zeros
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_zero + kRaZeroCells * C);
  const int tid = threadIdx.x, lane = tid & 31, l8 = lane & 7, grp = lane >> 3, warp = tid >> 5;
  int bid = blockIdx.x;
  const int ts = bid % d_split; bid /= d_split;
  const int th = bid % tiles_h; bid /= tiles_h;
  const int tw = bid % tiles_w;
  const int bn = img0 + bid / tiles_w;
  const int h0 = th * kRunHB, w0 = tw * kRaTW;
  const int d_begin = ts * d_per_cta, d_end = min(D, d_begin + d_per_cta);
  const int HW = H * W;
  const int rows_here = min(kRunHB, H - h0), cols_here = min(kRaTW, W - w0);

  // ---- context tile by TMA: one bulk copy per image row (cols_here * C contiguous floats)
  if (tid == 0) {
    mbar_init(s_bar, 1);
    fence_proxy_async();
  }
  if (fill) {
    for (int i = tid; i < kRaZeroCells * C / 4; i += kRaThreads) reinterpret_cast<float4 *>(s_zero)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (rows_here < kRunHB || cols_here < kRaTW) {       // ragged tile: rows / columns the copies do not cover read as zeros
    for (int i = tid; i < kRunHB * kRaTW * C / 4; i += kRaThreads) reinterpret_cast<float4 *>(s_ctx)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    const uint32_t row_bytes = (uint32_t)cols_here * C * 4;
    mbar_expect_tx(s_bar, row_bytes * rows_here);
    for (int r = 0; r < rows_here; ++r)
      tma_load_1d(s_ctx + r * (kRaTW * C), ctx_nhwc + (((int64_t)bn * H + h0 + r) * W + w0) * C, row_bytes, s_bar);
  }

  // ---- zero-fill of this CTA's share of the empty BEV cells.  A block of 32 cells that is entirely
  // empty (most of the grid) is ONE 10-KB TMA bulk store from the zero buffer; a mixed block is
  // written with ordinary 16-byte stores (many small bulk copies would serialise in the copy engine
  // and delay the context tiles queued behind them).  One block per warp per round, so the
  // cell_start loads of a CTA's blocks are all in flight together.
  bool issued_bulk = false;
  if (fill) {
    constexpr int kWarps = kRaThreads / 32, C4 = C / 4;
    const int64_t blocks = (num_cells + kRaZeroCells - 1) / kRaZeroCells;
    const int64_t per_cta = (blocks + gridDim.x - 1) / gridDim.x;
    const int64_t blk_begin = (int64_t)blockIdx.x * per_cta, blk_end = min(blocks, blk_begin + per_cta);
    for (int64_t blk = blk_begin + warp; blk < blk_end; blk += kWarps) {
      const int64_t off = blk * kRaZeroCells;
      const int ncell = (int)min((int64_t)kRaZeroCells, num_cells - off);
      const int64_t c0 = cell_base + off;
      const int cs = __ldg(cell_start + c0 + min(lane, ncell)), ce = __ldg(cell_start + c0 + min(lane + 1, ncell));
      const unsigned em = __ballot_sync(kFull, lane < ncell && ce == cs);
      if (ncell == kRaZeroCells && em == kFull) {
        if (lane == 0) {
          if (hints) tma_store_1d_hint(out + c0 * C, s_zero, kRaZeroCells * C * 4, l2_policy_evict_first());
          else tma_store_1d(out + c0 * C, s_zero, kRaZeroCells * C * 4);
          issued_bulk = true;
        }
      } else if (em) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 *o4 = reinterpret_cast<float4 *>(out + c0 * C);
#pragma unroll 4
        for (int i = lane; i < ncell * C4; i += 32)
          if ((em >> (i / C4)) & 1u) stg_stream_f4(o4 + i, z);
      }
    }
    if (issued_bulk) tma_store_commit();
  }

  // staging role: thread = (bin sd, row sh) of a chunk, 4 columns
  const int sh = tid & 15, sd = tid >> 4;
  const bool srow = h0 + sh < H;
  const int64_t sbase = (int64_t)bn * D * HW + (int64_t)(h0 + sh) * W + w0;
  const int nchunks = (d_end - d_begin + kRaDC - 1) / kRaDC;

  // (code, depth) segments of chunk c -> stage c % 3, as asynchronous 16-byte copies
  auto issue_chunk = [&](int cidx) {
    if (cidx < nchunks) {
      const int st = cidx % kRaStages, d = d_begin + cidx * kRaDC + sd;
      int4 *dc = &s_code[st][sd][sh];
      float4 *dd = &s_dep[st][sd][sh];
      if (srow && d < d_end) {
        const int64_t gp = sbase + (int64_t)d * HW;
        if (vec) {
          cp_async16_runs(dc, run_code + gp);
          cp_async16_runs(dd, depth + gp);
        } else {
          int4 pc = make_int4(kRunDropped, kRunDropped, kRunDropped, kRunDropped);
          float4 pd = make_float4(0.f, 0.f, 0.f, 0.f);
          if (w0 + 0 < W) { pc.x = __ldg(run_code + gp + 0); pd.x = __ldg(depth + gp + 0); }
          if (w0 + 1 < W) { pc.y = __ldg(run_code + gp + 1); pd.y = __ldg(depth + gp + 1); }
          if (w0 + 2 < W) { pc.z = __ldg(run_code + gp + 2); pd.z = __ldg(depth + gp + 2); }
          if (w0 + 3 < W) { pc.w = __ldg(run_code + gp + 3); pd.w = __ldg(depth + gp + 3); }
          *dc = pc;
          *dd = pd;
        }
      } else {
        *dc = make_int4(kRunDropped, kRunDropped, kRunDropped, kRunDropped);
        *dd = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int64_t slot0 = __ldg(cell_start + cell_base);
  const int wl = warp & 3, dh = warp >> 2;
  const float *ctx_col = s_ctx + wl * C;
  const uint64_t pol_rows = l2_policy_evict_last();
  auto store_run = [&](int64_t slot, const float (&acc)[NREG]) {
    if (slot >= 0 && slot < capacity) {
      if (hints) g8_store_row_hint<NV2>(run_rows + slot * C, l8, acc, pol_rows);
      else g8_store_row_plain<NV2>(run_rows + slot * C, l8, acc);
    }
  };
  auto warp_union = [&](unsigned m) -> unsigned {       // any-group union of a per-group mask (warp-uniform)
    m |= __shfl_xor_sync(kFull, m, 8);
    m |= __shfl_xor_sync(kFull, m, 16);
    return m;
  };

  issue_chunk(0);
  issue_chunk(1);
  bool ctx_ready = false;
  for (int cidx = 0; cidx < nchunks; ++cidx) {
    asm volatile("cp.async.wait_group 1;" ::: "memory");   // chunk cidx has landed (cidx + 1 may still be in flight)
    __syncthreads();                                        // ... for every thread; and chunk cidx - 1 is fully reduced
    issue_chunk(cidx + 2);                                  // into the stage chunk cidx - 1 just left
    if (!ctx_ready) {                                       // first chunk: the context tile must have landed
      mbar_wait(s_bar, 0);
      ctx_ready = true;
    }
    const int st = cidx % kRaStages;
    auto load_col = [&](int dl, RunCol &q) {
      const int *cw = reinterpret_cast<const int *>(&s_code[st][dl][0]) + wl;
      const float *pw = reinterpret_cast<const float *>(&s_dep[st][dl][0]) + wl;
      q.c_lo = cw[4 * l8];
      q.c_hi = cw[4 * (8 + l8)];
      q.p_lo = pw[4 * l8];
      q.p_hi = pw[4 * (8 + l8)];
      const unsigned k_lo = __ballot_sync(kFull, q.c_lo != kRunDropped), k_hi = __ballot_sync(kFull, q.c_hi != kRunDropped);
      const unsigned h_lo = __ballot_sync(kFull, q.c_lo >= 0), h_hi = __ballot_sync(kFull, q.c_hi >= 0);
      q.km = ((k_lo >> (8 * grp)) & 0xffu) | (((k_hi >> (8 * grp)) & 0xffu) << 8);
      q.hm = ((h_lo >> (8 * grp)) & 0xffu) | (((h_hi >> (8 * grp)) & 0xffu) << 8);
    };
    {
      const int dl = dh * (kRaDC / 2) + 2 * grp;            // this group's bins: dl, dl + 1 (8 bins per warp)
      RunCol a, b;
      load_col(dl, a);
      load_col(dl + 1, b);
      const unsigned any = warp_union(a.km | b.km);
      if (any == 0u) continue;
      float acc_a[NREG], acc_b[NREG];
#pragma unroll
      for (int r = 0; r < NREG; ++r) acc_a[r] = acc_b[r] = 0.f;
      const bool single = __popc(a.hm) <= 1 && __popc(b.hm) <= 1;
      if (__all_sync(kFull, single)) {
        // fast path (level camera): at most one run per (bin, column) pair -> no per-row bookkeeping
#pragma unroll
        for (int h = 0; h < kRunHB; ++h) {
          if (!((any >> h) & 1u)) continue;                     // warp-uniform
          const float da = __shfl_sync(kFull, h < 8 ? a.p_lo : a.p_hi, h & 7, 8);
          const float db = __shfl_sync(kFull, h < 8 ? b.p_lo : b.p_hi, h & 7, 8);
          if (((a.km | b.km) >> h) & 1u) {
            float v[NREG];
            g8_lds_row<NV2>(ctx_col + h * (kRaTW * C), l8, v);
            if ((a.km >> h) & 1u) axpy_row<NREG>(acc_a, da, v);
            if ((b.km >> h) & 1u) axpy_row<NREG>(acc_b, db, v);
          }
        }
        // the run's slot sits in the code of its first row
        const int ha = a.hm ? __ffs(a.hm) - 1 : 0, hb = b.hm ? __ffs(b.hm) - 1 : 0;
        const int sa_lo = __shfl_sync(kFull, a.c_lo, ha & 7, 8), sa_hi = __shfl_sync(kFull, a.c_hi, ha & 7, 8);
        const int sb_lo = __shfl_sync(kFull, b.c_lo, hb & 7, 8), sb_hi = __shfl_sync(kFull, b.c_hi, hb & 7, 8);
        if (a.hm) store_run((int64_t)(ha < 8 ? sa_lo : sa_hi) - slot0, acc_a);
        if (b.hm) store_run((int64_t)(hb < 8 ? sb_lo : sb_hi) - slot0, acc_b);
      } else {
        // general geometry: several runs per pair; walk the rows once per pair with explicit run boundaries
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
          const RunCol &q = pass == 0 ? a : b;
          const unsigned any_q = warp_union(q.km);
          float acc[NREG];
#pragma unroll
          for (int r = 0; r < NREG; ++r) acc[r] = 0.f;
          int64_t slot = -1;
#pragma unroll
          for (int h = 0; h < kRunHB; ++h) {
            if (!((any_q >> h) & 1u)) continue;                 // warp-uniform
            const int cv = __shfl_sync(kFull, h < 8 ? q.c_lo : q.c_hi, h & 7, 8);
            const float dv = __shfl_sync(kFull, h < 8 ? q.p_lo : q.p_hi, h & 7, 8);
            if ((q.km >> h) & 1u) {
              if (cv >= 0) {                                     // first row of a run
                store_run(slot, acc);
                if (slot >= 0) {
#pragma unroll
                  for (int r = 0; r < NREG; ++r) acc[r] = 0.f;
                }
                slot = (int64_t)cv - slot0;
              }
              float v[NREG];
              g8_lds_row<NV2>(ctx_col + h * (kRaTW * C), l8, v);
              axpy_row<NREG>(acc, dv, v);
            }
          }
          store_run(slot, acc);
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (!ctx_ready) mbar_wait(s_bar, 0);                   // (empty depth range) never leave a copy in flight
  if (issued_bulk) tma_store_wait_read();                // the zero buffer must outlive the bulk stores reading it
}

}  // namespace bevpool

using namespace bevpool;

template <int NV2>
static int launch_stage_a(const PlanView &pv, const float *dp, const float *cx, float *rr, float *out, int img0,
                          int64_t cell_base, int64_t num_cells, int nb, int num_cams, int D, int H, int W,
                          int64_t capacity, int vec, int fill, int hints, int d_split, cudaStream_t s) {
  constexpr int C = 16 * NV2;
  const int tiles_h = (int)ceil_div64(H, kRunHB), tiles_w = (int)ceil_div64(W, kRaTW);
  const int d_per_cta = (int)(ceil_div64(ceil_div64(D, d_split), kRaDC) * kRaDC);      // whole chunks per CTA
  const int splits = (int)ceil_div64(D, d_per_cta);
  const int64_t ctas = (int64_t)nb * num_cams * tiles_w * tiles_h * splits;
  if (ctas >= (int64_t)INT32_MAX) return BEVPOOL_E_RANGE;
  const size_t smem = (size_t)kRunHB * kRaTW * C * 4 + (size_t)kRaStages * kRaDC * kRunHB * 32 + (size_t)kRaZeroCells * C * 4 + 16;
  if (smem > 48 * 1024)
    BEVPOOL_RETURN_IF_CUDA(cudaFuncSetAttribute(frustum_reduce_kernel<NV2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BEVPOOL_RETURN_IF_CUDA(launch_pdl_if(pdl_forward_enabled(), frustum_reduce_kernel<NV2>, dim3((unsigned)ctas), dim3(kRaThreads), smem, s,
      pv.run_code, dp, cx, rr, pv.cell_start, out, img0, cell_base, num_cells, D, H, W, splits, d_per_cta, tiles_h,
      tiles_w, capacity, vec, fill, hints));
  BEVPOOL_LAUNCH_CHECK();
  return BEVPOOL_OK;
}

static int sm_count_runs() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = kSMs;
  }
  return cached;
}

extern "C" int bevpool_fused_forward_runs(const void *plan, const void *depth, const void *context_nhwc,
                                          void *out_nhwc, int dtype, int batch, int num_cams, int depth_bins,
                                          int feat_h, int feat_w, int channels, int X, int Y, void *run_rows,
                                          int64_t run_rows_capacity, void *workspace, void *stream) {
  if (num_cams <= 0 || depth_bins <= 0 || feat_h <= 0 || feat_w <= 0) return BEVPOOL_E_ARG;
  const int64_t np = (int64_t)num_cams * depth_bins * feat_h * feat_w;
  int rc = check_plan_dims(batch, np, X, Y);
  if (rc) return rc;
  if (dtype != BEVPOOL_F32) return BEVPOOL_E_DTYPE;
  if (!g8_supported(channels)) return BEVPOOL_E_CHANNELS;
  if (!plan || !depth || !context_nhwc || !out_nhwc || !run_rows || !workspace || run_rows_capacity <= 0) return BEVPOOL_E_ARG;
  if (!aligned16(context_nhwc) || !aligned16(out_nhwc) || !aligned16(run_rows) || !aligned16(workspace)) return BEVPOOL_E_ALIGN;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const PlanView pv = plan_view(plan, batch, np, X, Y);
  const int64_t G = (int64_t)X * Y;
  const int vec = (feat_w % 4 == 0) && aligned16(depth);
  int fpc = env_int_runs("BEVPOOL_RUN_CHUNK", 0);           // frames per (stage A, stage B) pair; 0 = all
  if (fpc <= 0 || fpc > batch) fpc = batch;
  const int fill = env_int_runs("BEVPOOL_RUN_FILL", 1) != 0;        // 0: stage B fills the empty cells itself
  const int hints = env_int_runs("BEVPOOL_RUN_HINTS", 1) != 0;      // L2 eviction-priority hints on the fill / run-row stores
  int cps_b = env_int_runs("BEVPOOL_RUN_CPSB", 5);                   // stage B CTAs per SM (4 warps each)
  cps_b = cps_b < 1 ? 1 : (cps_b > kFwMaxCtasPerSm - 1 ? kFwMaxCtasPerSm - 1 : cps_b);
  const float *dp = static_cast<const float *>(depth), *cx = static_cast<const float *>(context_nhwc);
  float *rr = static_cast<float *>(run_rows), *out = static_cast<float *>(out_nhwc);
  const FastDiv one = make_fastdiv(1u);
  for (int b0 = 0; b0 < batch; b0 += fpc) {
    const int nb = batch - b0 < fpc ? batch - b0 : fpc;
    const int64_t cell_base = (int64_t)b0 * G, ncells = (int64_t)nb * G;
    // depth ranges per tile: enough CTAs for >= ~4 waves of 4 resident CTAs per SM
    const int64_t tiles = (int64_t)nb * num_cams * ceil_div64(feat_w, kRaTW) * ceil_div64(feat_h, kRunHB);
    int d_split = env_int_runs("BEVPOOL_RUN_DSPLIT", 0);
    if (d_split <= 0) {
      d_split = (int)ceil_div64((int64_t)16 * sm_count_runs(), tiles);
      const int max_split = (int)ceil_div64(depth_bins, kRaDC);
      d_split = d_split < 1 ? 1 : (d_split > max_split ? max_split : d_split);
    }
    BEVPOOL_G8_DISPATCH(channels, (rc = launch_stage_a<NV2>(pv, dp, cx, rr, out, b0 * num_cams, cell_base, ncells, nb, num_cams,
                                                            depth_bins, feat_h, feat_w, run_rows_capacity, vec, fill, hints, d_split, s)));
    if (rc) return rc;
    // stage B: even-share segmented sum of the run rows (identity ids), fill CTAs only if stage A did not fill
    const int period_b = fill ? 0 : cps_b + 1;
    const unsigned ctas_b = (unsigned)(sm_count_runs() * (period_b ? period_b : cps_b));
    const int slices = (int)(period_b ? ctas_b - ctas_b / period_b : ctas_b) * kFwWarpsPerCta * 4;
    float *ws_head = static_cast<float *>(workspace);
    float *ws_tail = ws_head + (size_t)slices * channels;
    cudaError_t le = cudaSuccess;
    BEVPOOL_G8_DISPATCH(channels, (le = launch_pdl_if(pdl_forward_enabled(), pool_forward_share_kernel<NV2, false, 4, true>, dim3(ctas_b),
                                                   dim3(kFwWarpsPerCta * 32), 0, s, pv.cell_start, (const int32_t *)nullptr,
                                                   pv.sorted_cells, (const float *)rr, (const float *)nullptr, out, ws_head,
                                                   ws_tail, cell_base, ncells, one, one, period_b, -1, 0)));
    BEVPOOL_RETURN_IF_CUDA(le);
    BEVPOOL_LAUNCH_CHECK();
    BEVPOOL_G8_DISPATCH(channels, (le = launch_pdl_if(pdl_forward_enabled(), pool_forward_fixup_kernel<NV2>,
                                                   dim3((unsigned)ceil_div64((int64_t)slices * 8, 128)), dim3(128), 0, s,
                                                   pv.cell_start, pv.sorted_cells, (const float *)ws_head,
                                                   (const float *)ws_tail, out, cell_base, ncells, slices)));
    BEVPOOL_RETURN_IF_CUDA(le);
    BEVPOOL_LAUNCH_CHECK();
  }
  return BEVPOOL_OK;
}
