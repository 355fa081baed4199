// setup.cuh — part of griddy_cuda.cu (one translation unit).  Kernels that assemble the network records on the device at upload time.
#pragma once

namespace griddy {

// ------------------------------------------------------------------------------------------------
// Set-up: the network is assembled on the device from the arrays that cross the ABI (staged through
// a reusable device buffer), so that no host copy of the 32-byte records is ever made.  Every kernel works
// on one id range of the held items: `item` / `lane` point at the range's first record (global id g0), the
// source arrays at its first compact entry.
__global__ void __launch_bounds__(256) k_item_build(Item* __restrict__ item, int64_t n, const uint8_t* __restrict__ kind,
                                                     const double* __restrict__ len, const uint8_t* __restrict__ dir,
                                                     const int32_t* __restrict__ seg, const int32_t* __restrict__ out,
                                                     const int32_t* __restrict__ junction, const int32_t* __restrict__ of,
                                                     const int32_t* __restrict__ ob, int64_t g0) {
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (r >= n) return;
  Item it;
  it.kind = kind[r];
  it.dir = dir[r];
  it.pad = 0;
  it.len = len[r];
  it.link = it.kind == GRIDDY_SEGMENT_LANE ? seg[r] : it.kind == GRIDDY_JUNCTION_LANE ? out[r] : it.kind == GRIDDY_SEGMENT ? of[r] : -1;
  it.link2 = it.kind == GRIDDY_JUNCTION_LANE ? junction[r] : it.kind == GRIDDY_SEGMENT ? ob[r] : -1;
  it.from_j = it.to_j = -1;
  it.loc_key = (uint32_t)(g0 + r);   // default locality key: the registration index
  item[r] = it;
}

__global__ void __launch_bounds__(256) k_item_set_grid(Item* __restrict__ item, int64_t n, const int32_t* __restrict__ from_j,
                                                        const int32_t* __restrict__ to_j) {
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (r < n) { item[r].from_j = from_j[r]; item[r].to_j = to_j[r]; }
}

__global__ void __launch_bounds__(256) k_item_set_key(Item* __restrict__ item, int64_t n, const uint32_t* __restrict__ key) {
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (r < n) item[r].loc_key = key[r];
}

__global__ void __launch_bounds__(256) k_lane_init(LaneDyn* __restrict__ lane, const Item* __restrict__ item, int64_t n) {
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (r >= n) return;
  const Item it = item[r];
  lane[r] = LaneDyn{-1, -1, -1, -1, 0, it.kind == GRIDDY_JUNCTION_LANE ? it.link2 : -1, 0, 0};
}

__global__ void __launch_bounds__(256) k_lane_set_tail(LaneDyn* __restrict__ lane, const int32_t* __restrict__ which,
                                                        const int32_t* __restrict__ tail, int64_t m) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < m) lane[which[i]].tail = tail[i];
}

}  // namespace griddy
