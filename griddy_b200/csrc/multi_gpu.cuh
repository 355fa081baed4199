// multi_gpu.cuh — part of griddy_cuda.cu (one translation unit).  Device side of the partitioned world: halo and migrant kernels.
#pragma once

namespace griddy {

// ------------------------------------------------------------------------------------------------
// Multi-GPU: the world is partitioned spatially, one part per rank.  Every item has an owner; a
// segment and its lanes share one, a junction and its lanes share one, so an actor changes rank only
// while it turns from a segment lane onto a junction lane or back (never while it goes on or off the
// road).  An actor lives on the rank that owns its container.  What a turning actor reads of the OLD
// world on the other side — the target lane's rearmost positive pos-param (route.scm:118-121) and the
// target junction's occupancy (route.scm:98-108) — is the halo, exchanged before k_advance; the actors
// that crossed are exchanged as fixed-size records before k_relink, which then inserts them by key
// exactly as it inserts local movers (their old lane and old pos-param travel along, because
// processing order decides equal-key winners).  Results are identical to a single-GPU run.

struct MigRec {
  double pos_old, pos_new, delta, buf, arrive, land_pos, speed, speed_dt;
  int32_t gid, st, container, mv_old, hop, tj, popped, n_items, land_tick, land_lane, dest_lane, dest_from_j,
      dest_to_j, pad;
  int32_t item_sleep[MIG_ITEMS], item_seg[MIG_ITEMS];
  double item_pos[MIG_ITEMS];
  uint8_t item_kind[MIG_ITEMS], item_side[MIG_ITEMS];
};
constexpr size_t MIG_HEADER = 8;   // int32 count (reserved by remote atomics), int32 flag (tick + 1 once the sender is done)

__device__ __forceinline__ int32_t ld_acquire_sys(const int32_t* ptr) {
  int32_t v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int32_t* ptr, int32_t v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}

// Halo out: for every lane of mine that a neighbour's actors can enter, its smallest key in (0, ...]
// (the walk route.scm:118-121 would do), +inf if there is none; for every such junction, its occupancy.
// One launch for all neighbours: `items` and `out` are the neighbours' lists back to back (each
// neighbour's part of `out` is what is sent to it); a junction is listed as ~id.
__global__ void __launch_bounds__(128) k_halo_pack(Params p, int par, const int32_t* __restrict__ items, int n,
                                                    double* __restrict__ out) {
  grid_dependency_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t item = items[i];
  if (item < 0) { out[i] = (double)p.occ[~item]; return; }
  const double* __restrict__ pos = p.pos[par];
  int32_t x = p.lane[item].tail;
  while (x >= 0 && !(pos[x] > 0.0)) x = p.sa[x].ahead;
  out[i] = x >= 0 ? pos[x] : __longlong_as_double(0x7ff0000000000000LL);
}

// Halo in: occupancy of the neighbours' junctions (the lane values are read in place from halo_recv).
// One launch for all neighbours: junctions[i] is a junction another rank owns, where[i] its place in
// halo_recv.
__global__ void __launch_bounds__(128) k_halo_unpack(Params p, const int32_t* __restrict__ junctions,
                                                      const int32_t* __restrict__ where, int n) {
  grid_dependency_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p.occ[junctions[i]] = (int32_t)p.halo_recv[where[i]];
}

// Emigrants -> one record each, written straight into the receive buffer of the rank that owns their
// new lane (peer memory over NVLink: remote atomic for the slot, remote stores for the record), so only
// the actors that actually crossed travel, not a bound-sized message.  The block that finishes last
// raises the flag in every neighbour's buffer: no message follows the records.
__global__ void __launch_bounds__(128) k_emig_pack(Params p, int tick) {
  __shared__ bool s_last;
  grid_dependency_wait();
  const int par = tick & 1;
  const int32_t n = p.counters[CNT_STRIDE * par + CNT_EMIG];
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int a = p.emig[i];
    if (p.cold[a].dead_mark == tick + 1) continue;   // a ghost's move is void
    const int32_t lane = p.cold[a].container;
    const int32_t to = p.owner[lane];
    uint8_t* buf = p.mig_peer[to][par];
    const int32_t at = atomicAdd_system((int32_t*)buf, 1);
    if (at >= p.mig_cap[to]) { dev_error(p, DEV_ERR_MIG_OVERFLOW); continue; }
    MigRec r;
    r.pos_old = p.pos[par][a];
    r.pos_new = p.pos[par ^ 1][a];
    r.delta = p.db[a].x;
    r.buf = p.db[a].y;
    r.arrive = p.coldf[a].arrive;
    r.land_pos = p.coldf[a].land_pos;
    r.speed = p.coldf[a].speed;
    r.speed_dt = p.coldf[a].speed_dt;
    r.gid = p.cold[a].gid;
    // (k_unlink, which may be running beside this kernel, clears ALIVE | MOVED | EMIG of an emigrant's
    // slot; an emigrant had all three set)
    r.st = (int32_t)(((p.sa[a].st | ST_ALIVE | ST_MOVED) & ~ST_EMIG) | ST_IMMIG);
    r.container = lane;
    r.mv_old = p.cold[a].mv_old;
    r.hop = p.cold[a].hop;
    r.tj = p.tj[a] >= 0 ? (p.tj[a] & ~TJ_REMOTE) : -1;
    r.dest_lane = p.cold[a].dest_lane;
    r.dest_from_j = p.cold[a].dest_from_j;
    r.dest_to_j = p.cold[a].dest_to_j;
    r.pad = 0;
    const int32_t beg = p.coldf[a].agenda_beg, end = p.coldf[a].agenda_end;
    r.popped = p.cold[a].agenda_cur - beg;
    r.n_items = end - beg;
    r.land_tick = p.coldf[a].land_tick;
    r.land_lane = p.coldf[a].land_lane;
    if (r.n_items > MIG_ITEMS) { dev_error(p, DEV_ERR_MIG_AGENDA); r.n_items = MIG_ITEMS; }
    for (int k = 0; k < MIG_ITEMS; ++k) {
      const bool in = k < r.n_items;
      r.item_kind[k] = in ? p.item_kind[beg + k] : 0;
      r.item_side[k] = in ? p.item_side[beg + k] : 0;
      r.item_sleep[k] = in ? p.item_sleep[beg + k] : 0;
      r.item_seg[k] = in ? p.item_seg[beg + k] : -1;
      r.item_pos[k] = in ? p.item_pos[beg + k] : 0.0;
    }
    ((MigRec*)(buf + MIG_HEADER))[at] = r;
  }
  __threadfence_system();   // my records before my block's count
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(p.mig_done, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();   // every block's records before the flags
  if (threadIdx.x < MAX_RANKS && p.mig_peer[threadIdx.x][par]) st_release_sys((int32_t*)p.mig_peer[threadIdx.x][par] + 1, tick + 1);
  if (threadIdx.x == 0) *p.mig_done = 0;
}

// Immigrants: a fresh slot each, then onto their lane's inbox like any local mover.  One launch for all
// neighbours: blockIdx.y is the rank the records came from.
__global__ void __launch_bounds__(128) k_immig_unpack(Params p, int tick) {
  grid_dependency_wait();
  const int par = tick & 1;
  const uint8_t* buf = (const uint8_t*)p.mig_mine[blockIdx.y][par];
  if (!buf) return;
  if (threadIdx.x == 0) {   // the sender's flag: all its records of this tick are in my memory
    const long long t0 = clock64();
    while (ld_acquire_sys((const int32_t*)buf + 1) < tick + 1)
      if (clock64() - t0 > (20LL << 30)) { dev_error(p, DEV_ERR_MIG_TIMEOUT); break; }   // ~10 s: the neighbour is gone
  }
  __syncthreads();
  const int32_t n = min(ld_acquire_sys((const int32_t*)buf), p.mig_in[blockIdx.y]);
  const MigRec* recs = (const MigRec*)(buf + MIG_HEADER);
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const MigRec r = recs[i];
    const int32_t a = atomicAdd(&p.scal[SC_NSLOTS], 1);
    const int32_t it = atomicAdd(&p.scal[SC_NITEMS], r.n_items);
    if (a >= p.cap || it + r.n_items > p.icap) { dev_error(p, DEV_ERR_MIG_OVERFLOW); continue; }
    p.coldf[a].speed = r.speed;
    p.coldf[a].speed_dt = r.speed_dt;
    p.cold[a].gid = r.gid;
    p.coldf[a].agenda_beg = it;
    p.coldf[a].agenda_end = it + r.n_items;
    for (int k = 0; k < r.n_items; ++k) {
      p.item_kind[it + k] = r.item_kind[k];
      p.item_side[it + k] = r.item_side[k];
      p.item_sleep[it + k] = r.item_sleep[k];
      p.item_seg[it + k] = r.item_seg[k];
      p.item_pos[it + k] = r.item_pos[k];
    }
    p.sa[a] = SA{(uint32_t)r.st, -1};
    p.cold[a].behind = -1;
    p.pos[par][a] = r.pos_old;       // processing order among a lane's entrants looks at old lane and old key
    p.pos[par ^ 1][a] = r.pos_new;
    p.db[a] = make_double2(r.delta, r.buf);
    p.coldf[a].arrive = r.arrive;
    p.coldf[a].land_pos = r.land_pos;
    p.cold[a].container = r.container;
    p.cold[a].mv_old = r.mv_old;
    p.aux[a] = 0;
    p.cold[a].hop = r.hop;
    p.cold[a].dest_lane = r.dest_lane;
    p.cold[a].dest_from_j = r.dest_from_j;
    p.cold[a].dest_to_j = r.dest_to_j;
    p.tj[a] = tag_junction(p, r.tj);
    p.cold[a].agenda_cur = it + r.popped;
    p.cold[a].route_cur = 0;
    p.coldf[a].land_tick = r.land_tick;
    p.coldf[a].land_lane = r.land_lane;
    p.cold[a].dead_mark = 0;
    // onto the lane's inbox
    p.cold[a].inbox_next = atomicExch(&p.lane[r.container].inbox, a);
    if (atomicExch(&p.lane[r.container].mark_in, tick + 1) != tick + 1)
      p.touched[p.cap + atomicAdd(&p.counters[CNT_STRIDE * par + CNT_ENTERED], 1)] = r.container;
  }
}

}  // namespace griddy
