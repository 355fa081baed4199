// readback.cuh — part of griddy_cuda.cu (one translation unit).  Read-back: ghosts, population count, state formatting.
#pragma once

namespace griddy {

// ------------------------------------------------------------------------------------------------
// Read-back: format the world on the device (by actor id), then one copy per output array.

// The device's scalars (error bits, slots in use, dead slots) into the host's mapped mirror, for griddy_step_async.
__global__ void k_mirror_scalars(const int32_t* __restrict__ scal, int32_t* __restrict__ host) {
  if (threadIdx.x < SC_COUNT) host[threadIdx.x] = scal[threadIdx.x];
  __threadfence_system();
}

// Ghost rule for downloads: flag every leader that shares its follower's key.
__global__ void __launch_bounds__(256) k_flag_ghosts(Params p, int par, uint8_t* ghost) {
  const int a = blockIdx.x * 256 + threadIdx.x;
  if (a >= p.scal[SC_NSLOTS]) return;
  const uint32_t st = p.hot[a].st;
  if ((st & (ST_ALIVE | ST_ONROAD)) != (ST_ALIVE | ST_ONROAD)) return;
  const double* pos = p.pos[par];
  const double cur = pos[a];
  for (int32_t x = p.hot[a].ahead; x >= 0 && pos[x] == cur; x = p.hot[x].ahead) ghost[x] = 1;
}

// Living actors of the world: alive and not a ghost.
__global__ void __launch_bounds__(256) k_count_alive(Params p, const uint8_t* ghost, unsigned long long* out) {
  const int a = blockIdx.x * 256 + threadIdx.x;
  const bool live = a < p.scal[SC_NSLOTS] && (p.hot[a].st & ST_ALIVE) && !ghost[a];
  const uint32_t m = __ballot_sync(0xffffffffu, live);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}

struct PackOut {
  uint8_t *alive, *loc_kind, *side, *route_status;
  int32_t *container, *agenda_cur, *sleep_left, *route_head, *gid;
  double* pos;
  int by_slot;   // index the outputs by slot (multi-GPU read-back) instead of by actor id
};

__global__ void __launch_bounds__(256) k_pack_state(Params p, int par, const uint8_t* ghost, PackOut o) {
  const int a = blockIdx.x * 256 + threadIdx.x;
  if (a >= p.scal[SC_NSLOTS]) return;
  const uint32_t st = p.hot[a].st;
  const int32_t id = o.by_slot ? a : p.cold[a].gid;
  const int head = (st >> ST_HEAD_SHIFT) & 3, rs = (st >> ST_RS_SHIFT) & 3;
  const int32_t aux = p.hot[a].tj;
  if (o.gid) o.gid[id] = p.cold[a].gid;
  if (o.alive) o.alive[id] = ((st & ST_ALIVE) && !ghost[a]) ? 1 : 0;
  if (o.loc_kind) o.loc_kind[id] = (st & ST_ONROAD) ? 1 : 0;
  if (o.container) o.container[id] = p.cold[a].container;
  if (o.side) o.side[id] = (!(st & ST_ONROAD) && (st & ST_SIDE)) ? 1 : 0;
  if (o.pos) o.pos[id] = p.pos[par][a];
  if (o.agenda_cur) o.agenda_cur[id] = p.cold[a].agenda_cur - p.coldf[a].agenda_beg;
  if (o.sleep_left) o.sleep_left[id] = head == HEAD_SLEEP ? aux : 0;
  if (o.route_status) o.route_status[id] = (uint8_t)rs;
  if (o.route_head) o.route_head[id] = rs == GRIDDY_ROUTE_SOME ? p.cold[a].hop : -2;
}

// What the get-actors order (world.scm:24-30) of a slot depends on, for the read-back's sort: 32 bytes per slot.
constexpr uint32_t ORDER_GHOST = 0x80000000u;
struct __align__(16) OrderRec {
  double key;        // on the road: pos-param; off the road: the landing stamp's pos-param
  uint32_t st;       // ST_* | ORDER_GHOST
  int32_t container, land_tick, land_lane, gid, pad;
};
static_assert(sizeof(OrderRec) == 32, "one sector per slot");

__global__ void __launch_bounds__(256) k_pack_order(Params p, int par, const uint8_t* __restrict__ ghost, OrderRec* __restrict__ out) {
  const int a = blockIdx.x * 256 + threadIdx.x;
  if (a >= p.scal[SC_NSLOTS]) return;
  const uint32_t st = p.hot[a].st;
  const ColdF f = p.coldf[a];
  OrderRec r;
  r.key = (st & ST_ONROAD) ? p.pos[par][a] : f.land_pos;
  r.st = st | (ghost[a] ? ORDER_GHOST : 0u);
  r.container = p.cold[a].container;
  r.land_tick = f.land_tick;
  r.land_lane = f.land_lane;
  r.gid = p.cold[a].gid;
  r.pad = 0;
  out[a] = r;
}

// ------------------------------------------------------------------------------------------------
// State digest (griddy_state_digest): a 64-bit hash per living actor of its id and of everything
// griddy_download_actors reports about it, summed and xor-ed over the world.  Both reductions commute, so
// the digest does not depend on slot order or on how the world is partitioned over ranks: the ranks'
// digests add / xor up to the single world's.  griddy_b200/digest.py restates it in numpy for read-backs.
__device__ __forceinline__ uint64_t mix64(uint64_t x) {   // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256) k_digest(Params p, int par, const uint8_t* __restrict__ ghost,
                                                 unsigned long long* __restrict__ out /* sum, xor, living */) {
  const int a = blockIdx.x * 256 + threadIdx.x;
  uint64_t h = 0;
  bool live = false;
  if (a < p.scal[SC_NSLOTS]) {
    const uint32_t st = p.hot[a].st;
    if ((st & ST_ALIVE) && !ghost[a]) {
      live = true;
      const Cold k = p.cold[a];
      const int head = (st >> ST_HEAD_SHIFT) & 3, rs = (st >> ST_RS_SHIFT) & 3;
      const uint64_t on_road = (st & ST_ONROAD) ? 1 : 0, side = (!on_road && (st & ST_SIDE)) ? 1 : 0;
      const uint32_t popped = (uint32_t)(k.agenda_cur - p.coldf[a].agenda_beg);
      const uint32_t sleep_left = head == HEAD_SLEEP ? (uint32_t)p.hot[a].tj : 0u;
      const uint32_t route_head = (uint32_t)(rs == GRIDDY_ROUTE_SOME ? k.hop : -2);
      h = mix64((uint64_t)(int64_t)k.gid + 1);
      h = mix64(h ^ ((uint64_t)(uint32_t)k.container | (on_road << 32) | (side << 33) | ((uint64_t)rs << 34)));
      h = mix64(h ^ (uint64_t)__double_as_longlong(p.pos[par][a]));
      h = mix64(h ^ ((uint64_t)popped | ((uint64_t)sleep_left << 32)));
      h = mix64(h ^ (uint64_t)route_head);
    }
  }
  uint64_t sum = h, x = h;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, d);
    x ^= __shfl_xor_sync(0xffffffffu, x, d);
  }
  const uint32_t m = __ballot_sync(0xffffffffu, live);
  if ((threadIdx.x & 31) == 0 && m) {
    atomicAdd(&out[0], (unsigned long long)sum);
    atomicXor(&out[1], (unsigned long long)x);
    atomicAdd(&out[2], (unsigned long long)__popc(m));
  }
}

// Occupancy read-back (griddy_download_occupancy): actors per container, and per junction the device's own
// incrementally maintained counter less the ghosts standing on its lanes.
__global__ void __launch_bounds__(256) k_occupancy(Params p, const uint8_t* __restrict__ ghost, int32_t* __restrict__ lane_count,
                                                    int32_t* __restrict__ junction_count) {
  const int a = blockIdx.x * 256 + threadIdx.x;
  if (a >= p.scal[SC_NSLOTS]) return;
  const uint32_t st = p.hot[a].st;
  if (!(st & ST_ALIVE)) return;
  const int32_t c = p.cold[a].container;
  if (!ghost[a]) {
    if (lane_count) atomicAdd(&lane_count[c], 1);
  } else if (junction_count && meta_kind(p.lane[c].meta) == GRIDDY_JUNCTION_LANE) {
    atomicSub(&junction_count[p.lane[c].link2], 1);
  }
}

// ------------------------------------------------------------------------------------------------
// Drawing read-back: (get-pos actor) for every slot (spatial.scm:125-149), which is all that `draw`
// reads of an actor (simulate.scm:155-160 -> draw.scm:74-88).  geom holds 8 doubles per item,
// computed by the host with the reference's own get-pos / get-vec / curve (include/griddy_cuda.h,
// griddy_upload_geometry) — like the lane lengths, because they go through Chickadee's vec2:
//   off road      (v-beg + v-offset) + pos-param * (get-vec segment)      spatial.scm:125-136
//   segment lane  (get-pos lane 'beg) + pos-param * (get-vec lane)         spatial.scm:142-144
//   junction lane bezier-curve-point-at curve pos-param                    spatial.scm:145-146
// one rounding per vec2 operation and component.  A slot without a living actor gets NaN, NaN.
__device__ __forceinline__ double bezier_1d(double w0, double w1, double w2, double w3, double a, double b, double c, double d) {
  return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(w0, a), __dmul_rn(w1, b)), __dmul_rn(w2, c)), __dmul_rn(w3, d));
}

// A segment lane's geometry record uses two of its four vec2; the device copy carries NaN in the third, so that the
// frame kernels tell a straight lane from a junction lane's curve without fetching the lane's record.
__global__ void __launch_bounds__(256) k_geom_mark(const LaneRec* __restrict__ lane, double2* __restrict__ geom, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n && meta_kind(lane[i].meta) == GRIDDY_SEGMENT_LANE) geom[4 * i + 2].x = __longlong_as_double(0x7ff8000000000000LL);
}

// (get-pos actor) of the living actor in slot a: st its flags, t its pos-param.
__device__ __forceinline__ double2 get_pos_xy(const Params& p, const double2* __restrict__ geom, int a, uint32_t st, double t) {
  const int32_t c = p.cold[a].container;
  const double2* g = geom + 4 * (int64_t)c;
  const double2 g0 = g[0], g1 = g[1], g2 = g[2];
  if (!(st & ST_ONROAD) || g2.x != g2.x) {
    const double2 o = (!(st & ST_ONROAD) && (st & ST_SIDE)) ? g2 : g0;
    const double2 v = g1;
    return make_double2(__dadd_rn(o.x, __dmul_rn(t, v.x)), __dadd_rn(o.y, __dmul_rn(t, v.y)));
  }
  // chickadee bezier-curve-point-at (not vendored; the form tests/golden/minischeme.py assumes)
  const double2 p0 = g0, p1 = g1, p2 = g2, p3 = g[3];
  const double u = __dsub_rn(1.0, t);
  const double w0 = __dmul_rn(__dmul_rn(u, u), u);
  const double w1 = __dmul_rn(__dmul_rn(__dmul_rn(3.0, u), u), t);
  const double w2 = __dmul_rn(__dmul_rn(__dmul_rn(3.0, u), t), t);
  const double w3 = __dmul_rn(__dmul_rn(t, t), t);
  return make_double2(bezier_1d(w0, w1, w2, w3, p0.x, p1.x, p2.x, p3.x), bezier_1d(w0, w1, w2, w3, p0.y, p1.y, p2.y, p3.y));
}

__global__ void __launch_bounds__(256) k_positions(Params p, int par, const double2* __restrict__ geom,
                                                    const uint8_t* __restrict__ ghost, double2* __restrict__ xy64,
                                                    float2* __restrict__ xy32, int32_t* __restrict__ gid, int n_bound) {
  const int a = blockIdx.x * 256 + threadIdx.x;
  if (a >= n_bound) return;   // the host's upper bound of the slots in use: what it will copy
  const bool used = a < p.scal[SC_NSLOTS];
  const uint32_t st = used ? p.hot[a].st : 0u;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  double2 xy = make_double2(nan, nan);
  int32_t id = -1;
  if (used && (st & ST_ALIVE) && !ghost[a]) {
    id = p.cold[a].gid;
    xy = get_pos_xy(p, geom, a, st, p.pos[par][a]);
  }
  if (xy64) xy64[a] = xy;
  if (xy32) xy32[a] = make_float2((float)xy.x, (float)xy.y);
  if (gid) gid[a] = id;
}

// ------------------------------------------------------------------------------------------------
// Frames for a `draw` loop (griddy_positions_begin / griddy_delta_begin): one kernel per frame, behind the tick.
// A block takes FRAME_TILE consecutive slots, FRAME_UNROLL per thread, strided so that every load is coalesced.
// The ghost rule is applied from the leader's side — a living on-road actor whose list follower (Hot::behind)
// holds an equal key is not part of the world — which needs no pass over the followers and no flag array.
//
// FULL frame: float x, y of every slot below n_bound, NaN where the slot holds no living actor.
// DELTA frame: only the slots whose drawn position differs from the last frame sent.  A position is a function of
// (container, road side, pos-param): the pos-param is compared with the one last sent (pos_drawn, DRAWN_NONE: the
// slot showed nothing), a change of container or side is flagged by k_slow in Hot::st (ST_REDRAW, cleared here).
// So an actor that stands still — waiting in a queue, asleep, idle — costs 16 bytes of HBM reads beyond its hot
// record and nothing on PCIe.  Output: one mask bit per slot (word j of a block covers its slots 32 j .. 32 j + 31),
// the changed slots' values packed in slot order per block, blocks in the order they reserved their room
// (blk_off[block] = first value of the block), and the total count — the last block to finish writes it where the
// host reads it (mapped pinned memory) and resets the device counters for the next frame.
// key != 0: the slot numbering changed (a reorder) or this is the first frame: every living slot is sent, every
// other slot below n_bound is reset to "shows nothing"; the host starts from an all-NaN frame.
constexpr int FRAME_THREADS = 256, FRAME_UNROLL = 4, FRAME_TILE = FRAME_THREADS * FRAME_UNROLL;
constexpr unsigned long long DRAWN_NONE = ~0ull;   // a NaN no pos-param ever is (all ones: set with cudaMemset 0xFF)

struct FrameOut {
  float2* xy32;                 // FULL: by slot
  unsigned long long* drawn;    // DELTA: by slot, bits of the pos-param last sent, or DRAWN_NONE
  uint32_t* mask;               //        FRAME_TILE / 32 words per block
  uint32_t* blk_off;            //        one per block
  float2* vals;                 //        packed values
  unsigned int* counter;        //        [0] values handed out, [1] blocks finished
  unsigned int* host_count;     //        mapped pinned memory: the frame's total
};

template <bool DELTA>
__global__ void __launch_bounds__(FRAME_THREADS) k_frame(Params p, int par, const double2* __restrict__ geom, FrameOut o,
                                                          int n_bound, int key) {
  __shared__ uint32_t s_cnt[FRAME_TILE / 32];
  __shared__ uint32_t s_base;
  __shared__ bool s_last;
  const int n_used = min(p.scal[SC_NSLOTS], n_bound);
  const double* __restrict__ pos = p.pos[par];
  const int4* __restrict__ hot4 = (const int4*)p.hot;
  const int tile0 = blockIdx.x * FRAME_TILE;
  const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t st[FRAME_UNROLL];
  int32_t bh[FRAME_UNROLL];
  double t[FRAME_UNROLL];
  unsigned long long drawn[FRAME_UNROLL];
#pragma unroll
  for (int k = 0; k < FRAME_UNROLL; ++k) {
    const int a = tile0 + k * FRAME_THREADS + threadIdx.x;
    const bool used = a < n_used;
    const int4 h0 = used ? hot4[2 * (size_t)a] : make_int4(0, -1, -1, -1);   // st, ahead, behind, tj
    st[k] = (uint32_t)h0.x;
    bh[k] = h0.z;
    t[k] = used ? pos[a] : 0.0;
    drawn[k] = (DELTA && a < n_bound) ? o.drawn[a] : DRAWN_NONE;
  }
  bool vis[FRAME_UNROLL], chg[FRAME_UNROLL];
#pragma unroll
  for (int k = 0; k < FRAME_UNROLL; ++k) {
    const int a = tile0 + k * FRAME_THREADS + threadIdx.x;
    const bool on_road = (st[k] & (ST_ALIVE | ST_ONROAD)) == (ST_ALIVE | ST_ONROAD);
    const int32_t b = on_road ? bh[k] : -1;
    const bool ghost = b >= 0 && pos[b] == t[k];   // (the follower is next in slot order: the same or a neighbouring line)
    vis[k] = (st[k] & ST_ALIVE) && !ghost;
    if (DELTA) {
      const unsigned long long now = (unsigned long long)__double_as_longlong(t[k]);
      chg[k] = key ? vis[k] : (vis[k] ? (now != drawn[k] || (st[k] & ST_REDRAW)) : drawn[k] != DRAWN_NONE);
      if (st[k] & ST_REDRAW) p.hot[a].st = st[k] & ~ST_REDRAW;   // (between ticks: nobody else touches the word)
      if (a < n_bound && (key || chg[k])) o.drawn[a] = vis[k] ? now : DRAWN_NONE;
    } else {
      chg[k] = a < n_bound;
    }
  }
  const float nanf_ = __int_as_float(0x7fc00000);
  float2 xy[FRAME_UNROLL];
#pragma unroll
  for (int k = 0; k < FRAME_UNROLL; ++k) {
    const int a = tile0 + k * FRAME_THREADS + threadIdx.x;
    xy[k] = make_float2(nanf_, nanf_);
    if (vis[k] && chg[k]) {
      const double2 v = get_pos_xy(p, geom, a, st[k], t[k]);
      xy[k] = make_float2((float)v.x, (float)v.y);
    }
    if (!DELTA && chg[k]) o.xy32[a] = xy[k];
  }
  if constexpr (DELTA) {
  // pack: rank of a changed slot = the block's room + changed slots before it in the block (slot order)
  uint32_t ballot[FRAME_UNROLL];
#pragma unroll
  for (int k = 0; k < FRAME_UNROLL; ++k) {
    ballot[k] = __ballot_sync(0xffffffffu, chg[k]);
    if (lane_id == 0) {
      const int word = k * (FRAME_THREADS / 32) + warp;
      s_cnt[word] = __popc(ballot[k]);
      o.mask[(size_t)blockIdx.x * (FRAME_TILE / 32) + word] = ballot[k];
    }
  }
  __syncthreads();
  if (warp == 0) {   // exclusive scan of the 32 word counts
    const uint32_t c = s_cnt[lane_id];
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane_id >= d) incl += up;
    }
    s_cnt[lane_id] = incl - c;
    if (lane_id == 31) {
      const uint32_t base = incl ? atomicAdd(&o.counter[0], incl) : 0u;
      s_base = base;
      o.blk_off[blockIdx.x] = base;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < FRAME_UNROLL; ++k)
    if (chg[k]) {
      const int word = k * (FRAME_THREADS / 32) + warp;
      o.vals[s_base + s_cnt[word] + __popc(ballot[k] & ((1u << lane_id) - 1u))] = xy[k];
    }
  // the last block to finish publishes the total and leaves the counters at zero for the next frame
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&o.counter[1], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    *o.host_count = atomicExch(&o.counter[0], 0u);
    o.counter[1] = 0u;
    __threadfence_system();
  }
  }   // DELTA
}

}  // namespace griddy
