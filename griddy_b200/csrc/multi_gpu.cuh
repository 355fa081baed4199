// multi_gpu.cuh — part of griddy_cuda.cu (one translation unit).  Device side of the partitioned world: halo and migrant kernels.
#pragma once

namespace griddy {

// ------------------------------------------------------------------------------------------------
// Multi-GPU: the world is partitioned spatially, one part per rank.  Every item has an owner; a
// segment and its lanes share one, a junction and its lanes share one, so an actor changes rank only
// while it turns from a segment lane onto a junction lane or back (never while it goes on or off the
// road).  An actor lives on the rank that owns its container.  What a turning actor reads of the OLD
// world on the other side — the target lane's rearmost positive pos-param (route.scm:118-121) and the
// target junction's occupancy (route.scm:98-108) — is the halo; the actors that crossed are handed over
// as records before k_link, which then inserts them by key exactly as it inserts local movers (their old
// lane and old pos-param travel along, because processing order decides equal-key winners).  Results
// are identical to a single-GPU run.
//
// Transport: every rank owns a WINDOW of device memory that its neighbours' kernels write into directly
// (peer stores over NVLink 5 / NVSwitch; the window is shared through cudaIpc between processes).  Per
// neighbour and tick parity it holds
//     the halo values that neighbour exports to me, and a flag
//     [MigHeader][MigRec x bound][AgendaItem x bound]: the actors that crossed from it to me this tick.
// A sender writes its data, then raises the flag to tick + 1 with a system-scope release; the receiver's
// next kernel spins on the flag with a system-scope acquire.  No message, no collective, no host in the loop.
// A buffer of parity q is written at tick t and again at tick t + 2: in between, the receiver has read it
// (tick t), then sent its own halo of tick t + 1 — which the sender waits for before its tick t + 1 goes
// on — so the second write cannot overtake the first read.

struct __align__(16) MigHeader {
  int32_t count;     // records written this tick (published by the sender's last packing block)
  int32_t flag;      // tick + 1 once everything of that tick is in place
  int32_t n_items;   // agenda items written this tick
  int32_t pad;
};
struct __align__(16) MigRec {
  double pos_old, pos_new, delta, buf, arrive, land_pos, speed, speed_dt;
  int32_t gid, st, container, mv_old, hop, tj, popped, n_items;   // n_items: agenda items NOT yet popped, at item_off
  int32_t land_tick, land_lane, dest_lane, dest_from_j, dest_to_j, item_off, route_cur, pad;
};
static_assert(sizeof(MigHeader) == 16 && sizeof(MigRec) == 128, "window layout");

__device__ __forceinline__ int32_t ld_acquire_sys(const int32_t* ptr) {
  int32_t v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int32_t* ptr, int32_t v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}
// Spin until *flag >= want (thread 0 of a block), ~10 s at most: then the neighbour is gone.
__device__ __forceinline__ void wait_flag(const Params& p, const int32_t* flag, int32_t want) {
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) < want)
    if (clock64() - t0 > (20LL << 30)) { atomicOr(&p.scal[SC_ERR], DEV_ERR_MIG_TIMEOUT); break; }
}

// Halo out: for every lane of mine that a neighbour's actors can enter, its smallest key in (0, ...]
// (the walk route.scm:118-121 would do), +inf if there is none; for every such junction, its occupancy.
// Runs as the first blocks of k_advance (advance.cuh): both read the OLD world only, so the halo leaves at the very
// start of the tick and travels over NVLink while the other blocks stream the actors — no extra launch, no second
// stream.  Params::halo_items holds the neighbours' lists back to back (a junction is listed as ~id); every value is
// stored straight into the reading rank's window, and the block that finishes last raises the flag in every
// neighbour's window.
__device__ __forceinline__ void halo_pack_block(const Params& p, const int tick, const int n_blocks) {
  __shared__ bool s_last;
  TRACE(p, tick, TR_HALO_PACK);
  const int par = tick & 1;
  const int n = p.halo_total;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += n_blocks * blockDim.x) {
    const int32_t item = p.halo_items[i];
    double v;
    if (item < 0) {
      v = (double)p.occ[~item];
    } else {
      const double* __restrict__ pos = p.pos[par];
      int32_t x = p.lane[item].tail;
      while (x >= 0 && !(pos[x] > 0.0)) x = p.hot[x].ahead;
      v = x >= 0 ? pos[x] : __longlong_as_double(0x7ff0000000000000LL);
    }
    int r = 0;
    while (r < MAX_RANKS - 1 && !(i >= p.halo_first[r] && i < p.halo_first[r] + p.halo_n[r])) ++r;
    p.halo_peer[r][par][i - p.halo_first[r]] = v;
  }
  __threadfence_system();   // my values before my block's count
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(p.halo_done, 1) == n_blocks - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();   // every block's values before the flags
  if (threadIdx.x < MAX_RANKS && p.hflag_peer[threadIdx.x][par]) st_release_sys(p.hflag_peer[threadIdx.x][par], tick + 1);
  if (threadIdx.x == 0) *p.halo_done = 0;
}

// Halo in: wait for every neighbour's flag of this tick, then take over the occupancy of the neighbours'
// junctions (the lane values are read in place from halo_recv by turn_onto): Params::imp_junc[i] is a junction another
// rank owns, imp_where[i] its place in halo_recv.  k_slow, next on the stream, may then read the halo.  Nothing else
// reads those entries of the occupancy table (k_advance leaves an actor gated by a remote junction to k_slow), so with
// one process per GPU this runs as the LAST blocks of k_advance — the neighbours' halo left with the first blocks of
// theirs — and costs no launch; with all ranks on one stream-ordered GPU (tests) it is the kernel k_halo_unpack.
__device__ __forceinline__ void halo_unpack_block(const Params& p, const int tick, const int blk, const int n_blocks) {
  const int par = tick & 1;
  if (threadIdx.x < MAX_RANKS && p.hflag_mine[threadIdx.x][par]) wait_flag(p, p.hflag_mine[threadIdx.x][par], tick + 1);
  __syncthreads();
  for (int i = blk * blockDim.x + threadIdx.x; i < p.n_imp_junc; i += n_blocks * blockDim.x)
    p.occ[p.imp_junc[i]] = (int32_t)__ldcg(&p.halo_recv[par][p.imp_where[i]]);
}

__global__ void __launch_bounds__(128) k_halo_unpack(Params p, int tick) {
  grid_dependency_wait();
  TRACE(p, tick, TR_HALO_UNPACK);
  halo_unpack_block(p, tick, (int)blockIdx.x, (int)gridDim.x);
}

// Emigrants -> one record each plus their remaining agenda items, written by the k_slow thread that moves the actor
// straight into the window of the rank that owns its new lane (everything the record holds is in that thread's
// registers), so only the actors that actually crossed travel, not a bound-sized message, and nothing packs them
// in a separate pass.  Places are handed out by atomics on this rank's own counters.  The block of k_slow that
// finishes last publishes the counts and raises the flag in every neighbour's window: nothing follows the records.
__device__ void emigrate(const Params& p, int a, int tick, uint32_t st, const Cold& k, int32_t tj, double pos_old,
                         double pos_new, double delta, double buf) {
  const int par = tick & 1;
  const ColdF f = p.coldf[a];
  const int to = p.owner[k.container];
  MigRec r;
  r.n_items = f.agenda_end - k.agenda_cur;
  const int32_t at = atomicAdd(&p.mig_count[to], 1);
  r.item_off = atomicAdd(&p.mig_count[MAX_RANKS + to], r.n_items);
  if (at >= p.mig_cap[to] || r.item_off + r.n_items > p.mig_icap[to]) { dev_error(p, DEV_ERR_MIG_OVERFLOW); return; }
  r.pos_old = pos_old;
  r.pos_new = pos_new;
  r.delta = delta;
  r.buf = buf;
  r.arrive = f.arrive;
  r.land_pos = f.land_pos;
  r.speed = k.speed;
  r.speed_dt = k.speed_dt;
  r.gid = k.gid;
  r.st = (int32_t)(((st | ST_ALIVE | ST_MOVED) & ~(ST_EMIG | ST_GHOST)) | ST_IMMIG);
  r.container = k.container;
  r.mv_old = k.mv_old;
  r.hop = k.hop;
  r.tj = tj >= 0 ? (tj & ~TJ_REMOTE) : tj;   // (negative: none, or the junction of the junction lane it now stands on)
  r.popped = k.agenda_cur - f.agenda_beg;
  r.land_tick = f.land_tick;
  r.land_lane = f.land_lane;
  r.dest_lane = k.dest_lane;
  r.dest_from_j = k.dest_from_j;
  r.dest_to_j = k.dest_to_j;
  r.route_cur = k.route_cur;
  r.pad = 0;
  uint8_t* area = p.mig_peer[to][par];
  ((MigRec*)(area + sizeof(MigHeader)))[at] = r;
  AgendaItem* items = (AgendaItem*)(area + sizeof(MigHeader) + (size_t)p.mig_cap[to] * sizeof(MigRec)) + r.item_off;
  for (int32_t q = 0; q < r.n_items; ++q) items[q] = p.agenda[k.agenda_cur + q];
  if (!p.publish_late) __threadfence_system();   // my record before my block's count (emig_publish): only the threads that wrote one pay for it
}

// End of k_slow on a partitioned world, every block: my records before my block's count; the last block publishes.
__device__ void emig_publish(const Params& p, int tick) {
  __shared__ bool s_last;
  const int par = tick & 1;
  __syncthreads();   // (every thread that wrote a record has fenced it, see emigrate)
  if (threadIdx.x == 0) s_last = atomicAdd(p.mig_done, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();   // every block's records before the headers
  if (threadIdx.x < MAX_RANKS && p.mig_peer[threadIdx.x][par]) {
    MigHeader* h = (MigHeader*)p.mig_peer[threadIdx.x][par];
    h->count = min(atomicExch(&p.mig_count[threadIdx.x], 0), p.mig_cap[threadIdx.x]);
    h->n_items = min(atomicExch(&p.mig_count[MAX_RANKS + threadIdx.x], 0), p.mig_icap[threadIdx.x]);
    st_release_sys(&h->flag, tick + 1);
  }
  if (threadIdx.x == 0) *p.mig_done = 0;
#ifdef GRIDDY_TRACE
  if (p.trace && threadIdx.x == 0) {
    unsigned long long t_;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
    p.trace[(tick & 63) * TR_COUNT + TR_EMIG_FLAG] = t_;
  }
#endif
}

// The same from the kernel after k_slow (its first block, all ranks as processes of their own): k_slow is complete, so
// the counts are final and nothing in k_slow has to count blocks or fence — one system fence here orders every record
// written by the kernel before, then the headers, then the flags.
__device__ __forceinline__ void emig_publish_late(const Params& p, int tick) {
  const int par = tick & 1;
  if (threadIdx.x < MAX_RANKS && p.mig_peer[threadIdx.x][par]) {
    __threadfence_system();
    MigHeader* h = (MigHeader*)p.mig_peer[threadIdx.x][par];
    h->count = min(atomicExch(&p.mig_count[threadIdx.x], 0), p.mig_cap[threadIdx.x]);
    h->n_items = min(atomicExch(&p.mig_count[MAX_RANKS + threadIdx.x], 0), p.mig_icap[threadIdx.x]);
    st_release_sys(&h->flag, tick + 1);
  }
#ifdef GRIDDY_TRACE
  if (p.trace && threadIdx.x == 0) {
    unsigned long long t_;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
    p.trace[(tick & 63) * TR_COUNT + TR_EMIG_FLAG] = t_;
  }
#endif
}

// Immigrants: a fresh slot each, then onto their lane's inbox like any local mover; `src` is the rank the records came
// from, `blk` of `n_blocks` this block's share of them.  Nothing here meets what k_unlink touches (fresh slots are in no
// list; inbox heads and list tails are different words), so with one process per GPU these run as the last blocks of
// k_unlink and the wait for the neighbours' flags hides behind it; with all ranks on one GPU it is the kernel k_immig_unpack.
__device__ __forceinline__ void immig_unpack_block(const Params& p, const int tick, const int src, const int blk, const int n_blocks) {
  const int par = tick & 1;
  const uint8_t* area = (const uint8_t*)p.mig_mine[src][par];
  if (!area) return;
  const MigHeader* h = (const MigHeader*)area;
  if (threadIdx.x == 0) wait_flag(p, &h->flag, tick + 1);   // the sender's flag: all its records of this tick are in my memory
  __syncthreads();
#ifdef GRIDDY_TRACE
  if (p.trace && blk == 0 && threadIdx.x == 0) {   // the later of the neighbours' flags
    unsigned long long t_;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
    atomicMax(&p.trace[(tick & 63) * TR_COUNT + TR_IMMIG_GO], t_);
  }
#endif
  const int32_t n = min(ld_acquire_sys(&h->count), p.mig_in[src]);
  const MigRec* recs = (const MigRec*)(area + sizeof(MigHeader));
  const AgendaItem* payload = (const AgendaItem*)(area + sizeof(MigHeader) + (size_t)p.mig_in[src] * sizeof(MigRec));
  for (int32_t i = blk * blockDim.x + threadIdx.x; i < n; i += n_blocks * blockDim.x) {
    const MigRec r = recs[i];
    const int32_t a = atomicAdd(&p.scal[SC_NSLOTS], 1);
    const int32_t it = atomicAdd(&p.scal[SC_NITEMS], r.n_items);
    if (a >= p.cap) {   // no slot left: give the place back (the count stays a valid loop bound), report, drop the actor
      atomicSub(&p.scal[SC_NSLOTS], 1);
      dev_error(p, DEV_ERR_MIG_OVERFLOW);
      continue;
    }
    if (it + r.n_items > p.icap) {   // no room for its agenda: the slot stays counted, but dead
      p.hot[a].st = 0u;
      dev_error(p, DEV_ERR_MIG_OVERFLOW);
      continue;
    }
    Cold k;
    k.container = r.container;
    k.hop = r.hop;
    k.dest_lane = r.dest_lane;
    k.dest_from_j = r.dest_from_j;
    k.dest_to_j = r.dest_to_j;
    k.mv_old = r.mv_old;
    k.outbox_next = -1;
    k.speed = r.speed;
    k.speed_dt = r.speed_dt;
    k.agenda_cur = it;
    k.route_cur = r.route_cur;
    k.ghost_next = -1;
    k.gid = r.gid;
    ColdF f;
    f.arrive = r.arrive;
    f.land_pos = r.land_pos;
    f.agenda_beg = it - r.popped;   // only the items not yet popped travelled; the head is item `it`
    f.agenda_end = it + r.n_items;
    f.land_tick = r.land_tick;
    f.land_lane = r.land_lane;
    for (int q = 0; q < 8; ++q) f.pad[q] = 0;
    for (int q = 0; q < r.n_items; ++q) p.agenda[it + q] = payload[r.item_off + q];
    Hot h;
    h.st = (uint32_t)r.st;
    h.ahead = -1;
    h.behind = -1;
    h.tj = tag_junction(p, r.tj);
    h.delta = r.delta;
    h.buf = r.buf;
    p.hot[a] = h;
    p.pos[par][a] = r.pos_old;       // processing order among a lane's entrants looks at old lane and old key
    p.pos[par ^ 1][a] = r.pos_new;
    p.coldf[a] = f;
    // onto the lane's inbox
    LaneRec* lr = &p.lane[r.container];
    lr->e_pos = r.pos_new;
    k.inbox_next = atomicExch(&lr->inbox, a);
    p.cold[a] = k;
    if (k.inbox_next < 0) p.touched[atomicAdd(&p.counters[CNT_STRIDE * par + CNT_ENTERED], 1)] = r.container;
    else atomicOr(&lr->meta, META_MULTI);
  }
}

__global__ void __launch_bounds__(128) k_immig_unpack(Params p, int tick) {   // blockIdx.y: the rank they came from
  grid_dependency_wait();
  TRACE(p, tick, TR_IMMIG_UNPACK);
  immig_unpack_block(p, tick, (int)blockIdx.y, (int)blockIdx.x, (int)gridDim.x);
}

}  // namespace griddy
