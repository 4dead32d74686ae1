// Single-pass device-wide exclusive scan of uint32 (decoupled look-back), sm_100a.
//
// Used for: radix-pass tile histograms, the per-cell CSR offsets of the pooling plan
// and the first-occurrence flags of the voxelizer.  One launch, reads n and writes n
// elements (8 B/element of traffic), in-place allowed.
#pragma once
#include "common.cuh"

namespace bevpool {

constexpr int kScanThreads = 256;
constexpr int kScanTile = 4096;  // 8 warps x 4 rounds x 32 lanes x 4 elements

// workspace: status word per tile + one ticket counter; must be zero before launch
__host__ __device__ inline int64_t scan_num_tiles(int64_t n) { return ceil_div64(n, kScanTile); }
inline size_t scan_workspace_bytes(int64_t n) {
  return align_up((size_t)(scan_num_tiles(n) + 1) * 8, 256);
}

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

constexpr unsigned long long kScanAggregate = 1ull << 32;
constexpr unsigned long long kScanPrefix = 2ull << 32;

static __global__ void __launch_bounds__(kScanThreads)
scan_exclusive_kernel(const uint32_t *in, uint32_t *out, int64_t n,
                      unsigned long long *status /* [tiles] then ticket */) {
  __shared__ uint32_t s_warp[kScanThreads / 32];
  __shared__ uint32_t s_tile, s_prefix;
  unsigned int *ticket = reinterpret_cast<unsigned int *>(status + scan_num_tiles(n));
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warp_base = (int64_t)tile * kScanTile + warp * 512;

  uint4 v[4];
  uint32_t excl[4];
  uint32_t run = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t idx = warp_base + r * 128 + lane * 4;
    if (idx + 3 < n) {
      v[r] = *reinterpret_cast<const uint4 *>(in + idx);
    } else {
      v[r].x = idx < n ? in[idx] : 0u;
      v[r].y = idx + 1 < n ? in[idx + 1] : 0u;
      v[r].z = idx + 2 < n ? in[idx + 2] : 0u;
      v[r].w = 0u;
    }
    const uint32_t s = v[r].x + v[r].y + v[r].z + v[r].w;
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    excl[r] = run + incl - s;
    run += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) s_warp[warp] = run;
  __syncthreads();

  if (warp == 0) {
    const uint32_t w = lane < kScanThreads / 32 ? s_warp[lane] : 0u;
    uint32_t incl = w;
#pragma unroll
    for (int o = 1; o < kScanThreads / 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, kScanThreads / 32 - 1);
    if (lane < kScanThreads / 32) s_warp[lane] = incl - w;

    uint32_t exclusive = 0;
    if (tile == 0) {
      if (lane == 0) st_volatile_u64(status, kScanPrefix | total);
    } else {
      if (lane == 0) st_volatile_u64(status + tile, kScanAggregate | total);
      int64_t look = (int64_t)tile - 1;
      while (true) {
        const int64_t idx = look - lane;
        unsigned long long st;
        do {
          st = idx >= 0 ? ld_volatile_u64(status + idx) : kScanPrefix;
        } while (__any_sync(0xffffffffu, (st >> 32) == 0ull));
        const unsigned pm = __ballot_sync(0xffffffffu, (st >> 32) == 2ull);
        const int first = pm ? __ffs(pm) - 1 : 32;
        uint32_t contrib = lane <= first ? (uint32_t)(st & 0xffffffffull) : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        exclusive += contrib;
        if (pm) break;
        look -= 32;
      }
      if (lane == 0) st_volatile_u64(status + tile, kScanPrefix | (uint64_t)(exclusive + total));
    }
    if (lane == 0) s_prefix = exclusive;
  }
  __syncthreads();

  const uint32_t base = s_prefix + s_warp[warp];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t idx = warp_base + r * 128 + lane * 4;
    uint4 o;
    o.x = base + excl[r];
    o.y = o.x + v[r].x;
    o.z = o.y + v[r].y;
    o.w = o.z + v[r].z;
    if (idx + 3 < n) {
      *reinterpret_cast<uint4 *>(out + idx) = o;
    } else {
      if (idx < n) out[idx] = o.x;
      if (idx + 1 < n) out[idx + 1] = o.y;
      if (idx + 2 < n) out[idx + 2] = o.z;
    }
  }
}

// `workspace` must be zeroed (scan_workspace_bytes(n)); in/out 16-byte aligned; in == out allowed.
static inline int launch_scan_exclusive(const uint32_t *in, uint32_t *out, int64_t n, void *workspace,
                                 cudaStream_t stream) {
  if (n <= 0) return BEVPOOL_OK;
  const int64_t tiles = scan_num_tiles(n);
  scan_exclusive_kernel<<<(unsigned)tiles, kScanThreads, 0, stream>>>(
      in, out, n, static_cast<unsigned long long *>(workspace));
  BEVPOOL_LAUNCH_CHECK();
  return BEVPOOL_OK;
}

}  // namespace bevpool
