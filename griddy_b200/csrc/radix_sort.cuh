// radix_sort.cuh — stable LSD radix sort of (u64 key, u32 value) pairs, hand-written for sm_100a.
//
// Used by the periodic reorder of the actor slots (griddy_cuda.cu): keys are
// (locality key of the actor's container, coarse pos-param), values are the old slot numbers, so
// that after the sort actors of one lane sit next to each other in ascending pos-param and the
// leader-follower gathers of k_advance stay inside one cache line.  This is the "segmented sort by
// (lane, position)" of the north star, restricted to the bits that matter for locality; exact order
// is carried by the ahead pointers, which the reorder remaps rather than rebuilds.
//
// One pass = three launches over tiles of TILE keys (8-bit digits):
//   k_rs_hist     per-tile digit histogram (warp-aggregated shared atomics via __match_any_sync)
//   scan          exclusive scan of the digit-major histogram matrix (exclusive_scan_u32 below)
//   k_rs_scatter  re-reads the tile, ranks every key inside its digit with warp match/ballot and a
//                 per-warp count table in shared memory, stages the tile in shared memory ordered by
//                 digit, and writes each digit's run to its global position — so global stores are
//                 runs of ~TILE/256 consecutive pairs instead of one sector per pair.
// The element count is padded to a multiple of TILE by the caller with all-ones keys, which sort
// last; there are no bounds checks in the sort kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace rsort {

constexpr int RADIX = 256;
constexpr int THREADS = 256;
constexpr int ITEMS = 16;
constexpr int TILE = THREADS * ITEMS;   // 4096 pairs per block
constexpr int WARPS = THREADS / 32;
constexpr size_t SCATTER_SMEM = TILE * (sizeof(uint64_t) + sizeof(uint32_t)) + (WARPS + 3) * RADIX * sizeof(uint32_t);

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// hist[d * n_tiles + tile] = number of keys of the tile whose digit is d
__global__ void __launch_bounds__(THREADS) k_rs_hist(const uint64_t* __restrict__ keys, int shift,
                                                      uint32_t* __restrict__ hist, int n_tiles) {
  __shared__ uint32_t h[RADIX];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t* tile = keys + (size_t)blockIdx.x * TILE;
#pragma unroll 4
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t d = (uint32_t)(tile[i * THREADS + threadIdx.x] >> shift) & (RADIX - 1);
    const uint32_t m = __match_any_sync(0xffffffffu, d);
    if ((m & lanemask_lt()) == 0) atomicAdd(&h[d], __popc(m));
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(THREADS) k_rs_scatter(const uint64_t* __restrict__ kin,
                                                         const uint32_t* __restrict__ vin,
                                                         uint64_t* __restrict__ kout,
                                                         uint32_t* __restrict__ vout, int shift,
                                                         const uint32_t* __restrict__ hist_scanned,
                                                         int n_tiles) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint64_t* s_keys = (uint64_t*)smem;
  uint32_t* s_vals = (uint32_t*)(s_keys + TILE);
  uint32_t* s_wcnt = s_vals + TILE;            // [WARPS][RADIX] counts of the current round
  uint32_t* s_base = s_wcnt + WARPS * RADIX;   // keys of each digit seen in earlier rounds
  uint32_t* s_start = s_base + RADIX;          // tile-local start of each digit's run
  uint32_t* s_goff = s_start + RADIX;          // global start of each digit's run for this tile
  const int t = threadIdx.x, w = t >> 5;
  for (int k = 0; k < WARPS; ++k) s_wcnt[k * RADIX + t] = 0;
  s_base[t] = 0;
  s_goff[t] = hist_scanned[(size_t)t * n_tiles + blockIdx.x];
  __syncthreads();

  const size_t base = (size_t)blockIdx.x * TILE;
  uint64_t key[ITEMS];
  uint32_t val[ITEMS], rank[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    key[i] = kin[base + i * THREADS + t];
    val[i] = vin[base + i * THREADS + t];
  }
  // stable rank of every key inside its digit: rounds in order, warps in order, lanes in order
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t d = (uint32_t)(key[i] >> shift) & (RADIX - 1);
    const uint32_t m = __match_any_sync(0xffffffffu, d);
    const uint32_t r = __popc(m & lanemask_lt());
    if (r == 0) s_wcnt[w * RADIX + d] = __popc(m);
    __syncthreads();
    uint32_t before = s_base[d];
    for (int k = 0; k < w; ++k) before += s_wcnt[k * RADIX + d];
    rank[i] = before + r;
    __syncthreads();
    uint32_t tot = 0;   // thread t owns digit t
#pragma unroll
    for (int k = 0; k < WARPS; ++k) {
      tot += s_wcnt[k * RADIX + t];
      s_wcnt[k * RADIX + t] = 0;
    }
    s_base[t] += tot;
    __syncthreads();
  }
  // exclusive scan of the 256 digit counts (s_base) -> s_start
  {
    const uint32_t c = s_base[t];
    uint32_t x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if ((t & 31) >= o) x += y;
    }
    if ((t & 31) == 31) s_wcnt[w] = x;   // s_wcnt is all zero here; reuse its first WARPS words
    __syncthreads();
    uint32_t wbase = 0;
    for (int k = 0; k < w; ++k) wbase += s_wcnt[k];
    s_start[t] = wbase + x - c;
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t d = (uint32_t)(key[i] >> shift) & (RADIX - 1);
    const uint32_t at = s_start[d] + rank[i];
    s_keys[at] = key[i];
    s_vals[at] = val[i];
  }
  __syncthreads();
#pragma unroll 4
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t at = i * THREADS + t;
    const uint64_t k = s_keys[at];
    const uint32_t d = (uint32_t)(k >> shift) & (RADIX - 1);
    const size_t g = (size_t)s_goff[d] + (at - s_start[d]);
    kout[g] = k;
    vout[g] = s_vals[at];
  }
}

// ---- device-wide exclusive scan of a u32 array (three launches) --------------------------------
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t* total) {
  const int t = threadIdx.x, w = t >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if ((t & 31) >= o) x += y;
  }
  if ((t & 31) == 31) s_warp[w] = x;
  __syncthreads();
  uint32_t wbase = 0, all = 0;
#pragma unroll
  for (int k = 0; k < WARPS; ++k) {
    if (k < w) wbase += s_warp[k];
    all += s_warp[k];
  }
  __syncthreads();
  if (total) *total = all;
  return wbase + x - v;
}

__global__ void __launch_bounds__(THREADS) k_scan_reduce(const uint32_t* __restrict__ in, size_t n,
                                                          uint32_t* __restrict__ sums) {
  __shared__ uint32_t s_warp[WARPS];
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n) acc += in[base + i];
  uint32_t total;
  block_exclusive_scan(acc, s_warp, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// single block: exclusive scan of sums[0..m) in place
__global__ void __launch_bounds__(THREADS) k_scan_sums(uint32_t* sums, size_t m) {
  __shared__ uint32_t s_warp[WARPS];
  uint32_t carry = 0;
  for (size_t off = 0; off < m; off += THREADS) {
    const size_t i = off + threadIdx.x;
    const uint32_t v = i < m ? sums[i] : 0;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan(v, s_warp, &total);
    if (i < m) sums[i] = carry + ex;
    carry += total;
  }
}

// `out` may alias `in`: every thread reads its items before it writes them
__global__ void __launch_bounds__(THREADS) k_scan_apply(const uint32_t* in, size_t n,
                                                         const uint32_t* __restrict__ sums, uint32_t* out) {
  __shared__ uint32_t s_warp[WARPS];
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS], acc = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = base + i < n ? in[base + i] : 0;
    acc += v[i];
  }
  uint32_t run = sums[blockIdx.x] + block_exclusive_scan(acc, s_warp, nullptr);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
}

inline size_t scan_blocks(size_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }

// out may alias in.  `sums` needs scan_blocks(n) words.  Returns the number of launches.
inline int exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* sums, cudaStream_t s) {
  if (n == 0) return 0;
  const unsigned nb = (unsigned)scan_blocks(n);
  k_scan_reduce<<<nb, THREADS, 0, s>>>(in, n, sums);
  k_scan_sums<<<1, THREADS, 0, s>>>(sums, nb);
  k_scan_apply<<<nb, THREADS, 0, s>>>(in, n, sums, out);
  return 3;
}

// Scratch the sort needs for n_pad = n_tiles * TILE pairs.
struct Scratch {
  uint64_t* keys_alt = nullptr;
  uint32_t* vals_alt = nullptr;
  uint32_t* hist = nullptr;   // RADIX * n_tiles
  uint32_t* sums = nullptr;   // scan_blocks(RADIX * n_tiles)
};

inline size_t hist_words(int n_tiles) { return (size_t)RADIX * n_tiles; }

inline cudaError_t prepare() {
  return cudaFuncSetAttribute(k_rs_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCATTER_SMEM);
}

// Sorts n_tiles * TILE pairs by the low `key_bits` bits of the key.  The result is left in
// (*keys, *vals); the pointers are swapped with the scratch buffers when the pass count is odd.
inline int sort_pairs(uint64_t** keys, uint32_t** vals, int n_tiles, int key_bits, Scratch* sc, cudaStream_t s) {
  int launches = 0;
  for (int shift = 0; shift < key_bits; shift += 8) {
    k_rs_hist<<<n_tiles, THREADS, 0, s>>>(*keys, shift, sc->hist, n_tiles);
    launches += 1 + exclusive_scan_u32(sc->hist, sc->hist, hist_words(n_tiles), sc->sums, s);
    k_rs_scatter<<<n_tiles, THREADS, SCATTER_SMEM, s>>>(*keys, *vals, sc->keys_alt, sc->vals_alt, shift, sc->hist, n_tiles);
    launches += 1;
    uint64_t* tk = *keys; *keys = sc->keys_alt; sc->keys_alt = tk;
    uint32_t* tv = *vals; *vals = sc->vals_alt; sc->vals_alt = tv;
  }
  return launches;
}

}  // namespace rsort
