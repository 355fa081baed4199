// setup.cuh — part of griddy_cuda.cu (one translation unit).  Kernels that assemble the network records on the device at upload time.
#pragma once

namespace griddy {

// ------------------------------------------------------------------------------------------------
// Set-up: the network is assembled on the device from the arrays that cross the ABI (staged through
// a reusable device buffer), so that no host copy of the 64-byte records is ever made.  Every kernel works
// on one id range of the held items: `lane` points at the range's first record, the source arrays at its
// first compact entry.
__global__ void __launch_bounds__(256) k_item_build(LaneRec* __restrict__ lane, int64_t n, const uint8_t* __restrict__ kind,
                                                     const double* __restrict__ len, const uint8_t* __restrict__ dir,
                                                     const int32_t* __restrict__ seg, const int32_t* __restrict__ out,
                                                     const int32_t* __restrict__ junction, const int32_t* __restrict__ of,
                                                     const int32_t* __restrict__ ob) {
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (r >= n) return;
  LaneRec it;
  const int k = kind[r];
  it.len = len[r];
  it.link = k == GRIDDY_SEGMENT_LANE ? seg[r] : k == GRIDDY_JUNCTION_LANE ? out[r] : k == GRIDDY_SEGMENT ? of[r] : -1;
  it.link2 = k == GRIDDY_JUNCTION_LANE ? junction[r] : k == GRIDDY_SEGMENT ? ob[r] : -1;
  it.to_j = -1;
  it.meta = (uint32_t)(k & 7) | (dir[r] ? META_DIR : 0u);
  it.e_pos = 0.0;
  it.tail = it.inbox = it.outbox = it.ghosts = -1;
  it.turn[0] = it.turn[1] = it.turn[2] = it.turn[3] = -1;
  lane[r] = it;
}

__global__ void __launch_bounds__(256) k_item_set_grid(LaneRec* __restrict__ lane, int64_t n, const int32_t* __restrict__ to_j,
                                                        const int32_t* __restrict__ turn4) {
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (r >= n) return;
  lane[r].to_j = to_j[r];
  *(int4*)lane[r].turn = ((const int4*)turn4)[r];
}

// A new world on the same network: empty lists, no marks.
__global__ void __launch_bounds__(256) k_lane_init(LaneRec* __restrict__ lane, int64_t n) {
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (r >= n) return;
  lane[r].meta &= META_KIND | META_DIR | META_TOJ_REMOTE;   // (the partition's mark belongs to the network)
  *(int4*)&lane[r].tail = make_int4(-1, -1, -1, -1);
}

__global__ void __launch_bounds__(256) k_lane_set_tail(LaneRec* __restrict__ lane, const int32_t* __restrict__ which,
                                                        const int32_t* __restrict__ tail, int64_t m) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < m) lane[which[i]].tail = tail[i];
}

// Partitioned world: mark the lanes (ids in `which`) that reach a junction another rank owns, so that a lane change
// knows it from the target lane's record instead of a gather into the owner table.
__global__ void __launch_bounds__(256) k_lane_clear_remote(LaneRec* __restrict__ lane, int64_t m) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < m) lane[i].meta &= ~META_TOJ_REMOTE;
}
__global__ void __launch_bounds__(256) k_lane_mark_remote(LaneRec* __restrict__ lane, const int32_t* __restrict__ which, int64_t m) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < m) lane[which[i]].meta |= META_TOJ_REMOTE;
}

}  // namespace griddy
