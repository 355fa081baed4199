// advance.cuh — part of griddy_cuda.cu (one translation unit).  advance$ for every actor: k_advance (the common cases) and k_slow (everything else).
#pragma once

namespace griddy {

__device__ __forceinline__ void dev_error(const Params& p, int bit) { atomicOr(&p.scal[SC_ERR], bit); }

// location.scm:16-24 through the uploaded per-segment table
__device__ __forceinline__ int32_t outer_lane(const Params& p, int32_t seg, uint32_t side) {
  return side ? p.lane[seg].link2 : p.lane[seg].link;
}

// Dimension-ordered next hop on procedural grids (include/griddy_cuda.h, griddy_set_routing_grid):
// from the segment lane `it` towards the lane that leaves junction dest_from_j for junction dest_to_j.
__device__ __forceinline__ int32_t grid_next_hop(const Params& p, const LaneRec& it, int32_t dest_from_j, int32_t dest_to_j) {
  const int n = p.grid_n;
  const int32_t J = it.to_j, E = dest_from_j;
  int h;
  if (J == E) {
    const int32_t T = dest_to_j;
    h = (T / n != E / n) ? (T / n > E / n ? 0 : 1) : (T % n > E % n ? 2 : 3);
  } else if (J / n != E / n) {
    h = E / n > J / n ? 0 : 1;
  } else {
    h = E % n > J % n ? 2 : 3;
  }
  return h == 0 ? it.turn[0] : h == 1 ? it.turn[1] : h == 2 ? it.turn[2] : it.turn[3];   // (no dynamic index: the record stays in registers)
}

// Head of the route after the actor has reached `lane` (route.scm:48-51 evaluated lazily); `it` is
// lane's record, k holds the route's destination.
__device__ __forceinline__ int32_t route_head_on(const Params& p, int32_t lane, const LaneRec& it, const Cold& k, int32_t* rc) {
  if (p.route_off) {   // (k.dest_from_j: where the route ends in route_lane)
    const int32_t c = *rc;
    return c < k.dest_from_j ? p.route_lane[c] : HOP_ARRIVE;
  }
  if (lane == k.dest_lane) return HOP_ARRIVE;
  if (meta_kind(it.meta) == GRIDDY_JUNCTION_LANE) return it.link;
  const int32_t hop = grid_next_hop(p, it, k.dest_from_j, k.dest_to_j);
  if (hop < 0) dev_error(p, DEV_ERR_NO_ROUTE);
  return hop;
}

// The junction whose occupancy gates a (turn-onto lane) head (route.scm:105-108), or -1.  On a
// partitioned world a junction another rank owns carries TJ_REMOTE: its occupancy arrives with the
// halo, which k_advance does not wait for, so such actors are left to k_slow.
constexpr int32_t TJ_REMOTE = 1 << 30;
__device__ __forceinline__ int32_t tag_junction(const Params& p, int32_t j) {
  return (j >= 0 && p.owner && p.owner[j] != p.my_rank) ? (j | TJ_REMOTE) : j;
}
__device__ __forceinline__ int32_t junction_of_hop(const Params& p, int32_t hop) {
  if (hop < 0) return -1;
  const LaneRec& h = p.lane[hop];
  return tag_junction(p, meta_kind(h.meta) == GRIDDY_JUNCTION_LANE ? h.link2 : -1);
}

// The same for the head `nhop` of a route that has just reached lane `it`.  With the grid routing table a
// segment lane's next hop is a junction lane of the junction the lane reaches (LaneRec::to_j; checked by
// griddy_set_routing_grid) and a junction lane's next hop is a segment lane, so the hop's own record
// is not needed.
__device__ __forceinline__ int32_t junction_after(const Params& p, const LaneRec& it, int32_t nhop) {
  if (p.route_off) return junction_of_hop(p, nhop);
  if (!(nhop >= 0 && meta_kind(it.meta) == GRIDDY_SEGMENT_LANE) || it.to_j < 0) return -1;
  return (p.owner && (it.meta & META_TOJ_REMOTE)) ? (it.to_j | TJ_REMOTE) : it.to_j;   // (marked by griddy_mg_configure)
}

// New agenda head after a pop (actor.scm:28-31): its kind bits, and the sleep counter in *aux.
__device__ __forceinline__ uint32_t load_head(const Params& p, int32_t a, int32_t ac, int32_t* aux) {
  if (ac >= p.coldf[a].agenda_end) { *aux = 0; return HEAD_NONE; }
  const AgendaItem ai = p.agenda[ac];
  if (ai.kind == GRIDDY_SLEEP_FOR) { *aux = ai.sleep; return HEAD_SLEEP; }
  *aux = 0;
  return HEAD_TRAVEL;
}

// Work lists of the order maintenance, collected per block in shared memory and flushed with one global atomic per
// block (a same-address atomic per entry would serialise):
//   leavers  one record per actor that leaves its lane's list this tick — it changed lane, went off the road, jumped
//            back on its lane, or was found to be a ghost: {slot, the lane it leaves, that lane's junction if it is a
//            junction lane (else -1), 0}.  k_unlink takes them out, one thread per record.
//   entered  lanes that gained an actor: the first entrant — the one that finds the lane's inbox empty — lists the
//            lane, any other marks it (LaneRec::meta, META_MULTI), so k_link has one thread per lane and never
//            reads the chain through the cold records when there is one entrant.
struct TouchSink {
  int4* left;                     // shared: leaver records, LEAVER_STAGE entries; beyond that they go straight to the global list
  int32_t* n_left;
  int32_t *entered, *n_entered;   // shared: lanes that gained an actor, SLOW_THREADS entries
};
#ifdef SLOW_MIN_BLOCKS
#define SLOW_BOUNDS __launch_bounds__(SLOW_THREADS, SLOW_MIN_BLOCKS)
#else
#define SLOW_BOUNDS __launch_bounds__(SLOW_THREADS)
#endif
constexpr int SLOW_THREADS = 128;   // an actor enters at most one lane
constexpr int LEAVER_STAGE = 2 * SLOW_THREADS;   // an actor leaves at most one lane; a follower may find a chain of ghosts

__device__ __forceinline__ void push_leaver(const Params& p, int tick, int32_t z, int32_t lane, int32_t junction, const TouchSink& sink) {
  const int4 rec = make_int4(z, lane, junction, 0);
  const int32_t at = atomicAdd(sink.n_left, 1);
  if (at < LEAVER_STAGE) sink.left[at] = rec;
  else p.leavers[atomicAdd(&p.counters[CNT_STRIDE * (tick & 1) + CNT_LEFT], 1)] = rec;
}

// The junction whose occupancy counts the actors of `lane` (route.scm:98-104), or -1 if it is not a junction lane.
__device__ __forceinline__ int32_t junction_of_lane(const Params& p, int32_t lane) {
  const LaneRec& r = p.lane[lane];
  return meta_kind(r.meta) == GRIDDY_JUNCTION_LANE ? r.link2 : -1;
}
// An actor that stands on a junction lane has no gated route head (a junction lane leads onto a segment lane), so its
// Hot::tj is free to remember the junction it occupies: -2 - junction.  A lane change then needs nothing of the lane
// it leaves — not even its record.
__device__ __forceinline__ int32_t occupied_junction(int32_t tj) { return tj <= -2 ? -2 - tj : -1; }

// Returns the lane's previous inbox head (the caller stores it as the entrant's inbox_next).  `pos`: where it lands.
__device__ __forceinline__ int32_t enter_lane(const Params& p, int a, int32_t lane, double pos, const TouchSink& sink) {
  p.lane[lane].e_pos = pos;   // (several entrants overwrite each other: k_link then takes every one's own)
  const int32_t prev = atomicExch(&p.lane[lane].inbox, a);
  if (prev < 0) sink.entered[atomicAdd(sink.n_entered, 1)] = lane;   // the first entrant lists the lane
  else atomicOr(&p.lane[lane].meta, META_MULTI);                     // any other says that the inbox is a chain
  return prev;
}

// ------------------------------------------------------------------------------------------------
// route-step/turn-onto$ (route.scm:92-135) for an actor that reaches the end of its lane and is not
// held by a full junction: carry the rest of the tick over into the target lane, pop the route.
// Touches five 64-byte lines: the actor's cold record, its hot record, its new pos-param, the target lane's
// record and the pos-param of the target lane's rearmost actor.  Returns the new pos-param.
__device__ void emigrate(const Params& p, int a, int tick, uint32_t st, const Cold& k, int32_t tj, double pos_old,
                         double pos_new, double delta, double buf);   // multi_gpu.cuh
__device__ void emig_publish(const Params& p, int tick);
__device__ __forceinline__ void halo_pack_block(const Params& p, int tick, int n_blocks);
__device__ __forceinline__ void halo_unpack_block(const Params& p, int tick, int blk, int n_blocks);

__device__ __forceinline__ double turn_onto(const Params& p, const int a, const int tick, const uint32_t st,
                                            const double cur, const double* __restrict__ pos_old, const TouchSink& sink) {
  Cold k = p.cold[a];
  const int32_t tj_old = p.hot[a].tj;
  const int32_t target = k.hop, lane = k.container;
  const LaneRec it = p.lane[target];
  const double time_spent = __ddiv_rn(__dsub_rn(1.0, cur), k.speed);
  const double time_remaining = __dsub_rn(p.dt, time_spent);
  const double tinv = __ddiv_rn(1.0, it.len);
  const double tbuf = __ddiv_rn(p.two_size, it.len);
  double tnext = __dmul_rn(__dmul_rn(k.speed, time_remaining), tinv);   // (+ 0 delta)
  int32_t x = it.tail;
  const bool remote = x <= -2;   // multi-GPU: another rank owns the target lane
  if (remote) {                  // its smallest key in (0, ...] came with the halo (+inf: none)
    const double tlead = __ldcg(&p.halo_recv[tick & 1][-2 - x]);
    if (tlead <= __dadd_rn(tnext, tbuf)) tnext = __dsub_rn(tlead, tbuf);
  } else {
    while (x >= 0 && !(pos_old[x] > 0.0)) x = p.hot[x].ahead;   // smallest key in (0, ...]
    if (x >= 0) {
      const double tlead = pos_old[x];
      if (tlead <= __dadd_rn(tnext, tbuf)) tnext = __dsub_rn(tlead, tbuf);
    }
  }
  int32_t rc = p.route_off ? k.route_cur + 1 : 0;
  const int32_t nhop = route_head_on(p, target, it, k, &rc);
  // (ahead / behind stay: k_unlink reads them.  atomicOr: a follower may be flagging this actor as a ghost right now)
  const uint32_t more = ST_MOVED | ST_REDRAW | (nhop == HOP_ARRIVE ? ST_ARRIVE : 0u);
  const uint32_t st_before = atomicOr(&p.hot[a].st, more | (remote ? ST_EMIG : 0u));
  int32_t tjn = junction_after(p, it, nhop);
  if (meta_kind(it.meta) == GRIDDY_JUNCTION_LANE) tjn = -2 - it.link2;   // (no gated head on a junction lane: see occupied_junction)
  const double delta = __dmul_rn(k.speed_dt, tinv);
  p.hot[a].tj = tjn;
  p.hot[a].delta = delta;
  p.hot[a].buf = tbuf;
  k.hop = nhop;
  k.container = target;
  k.mv_old = lane;
  if (remote) {
    // The actor lives on another rank from now on: its record goes straight into that rank's window (k_unlink
    // then gives the slot up).  Not if it is a ghost — its follower holds its key, so its move is void
    // (file header): the follower's thread flags it, this one just keeps the record back.
    const int32_t b = p.hot[a].behind;
    if (!(b >= 0 && pos_old[b] == cur)) {
      if (p.route_off) k.route_cur = rc;
      emigrate(p, a, tick, st | more, k, tjn, cur, tnext, delta, tbuf);
    }
  } else {
    k.inbox_next = enter_lane(p, a, target, tnext, sink);
  }
  // (a follower that found this actor to be a ghost before the atomicOr above has listed it already)
  if (!(st_before & ST_GHOST)) push_leaver(p, tick, a, lane, occupied_junction(tj_old), sink);
  int4* kc = (int4*)&p.cold[a];   // the first sector, in two stores
  kc[0] = make_int4(k.container, k.hop, k.dest_lane, k.dest_from_j);
  kc[1] = make_int4(k.dest_to_j, k.mv_old, k.inbox_next, k.outbox_next);
  if (p.route_off) p.cold[a].route_cur = rc;
  l2_keep(&p.hot[a], 1);
  if (!remote) l2_keep(&p.lane[target], 4);
  return tnext;
}

// ------------------------------------------------------------------------------------------------
// advance$ for one actor (simulate.scm:93-111) with route.scm:53-141 inlined.  `a` is a slot; st,
// ah, cur, dlt, buf, lead were loaded by the caller (ah = -1 and lead = 0 when there is no leader).
// Returns the actor's new pos-param; everything else that changes is written here.
__device__ __forceinline__ double advance_actor(const Params& p, const int a, const int tick, const uint32_t st,
                                                int32_t ah, const double cur, const double dlt,
                                                const double buf, double lead,
                                                const double* __restrict__ pos_old, const TouchSink& sink) {
  const int head = (st >> ST_HEAD_SHIFT) & 3;
  const int rs = (st >> ST_RS_SHIFT) & 3;

  // nearest actor ahead on the lane (bbtree-find-min's only candidate, util.scm:152-162), skipping
  // ghosts: leaders that this actor replaced with an equal key at the end of the previous tick
  if (ah >= 0 && lead == cur) {
    const int32_t lane = p.cold[a].container;
    const int32_t junction = junction_of_lane(p, lane);
    do {
      // the first to notice hands it to k_unlink — unless the ghost changed lane itself this tick and is listed already
      if (!(atomicOr(&p.hot[ah].st, ST_GHOST) & (ST_GHOST | ST_MOVED))) push_leaver(p, tick, ah, lane, junction, sink);
      ah = p.hot[ah].ahead;
      if (ah >= 0) lead = pos_old[ah];
    } while (ah >= 0 && lead == cur);
  }

  if (rs == GRIDDY_ROUTE_SOME && head == HEAD_TRAVEL) {   // advance-on-route$  route.scm:137-141
    // get-pos-param-next on the current lane (route.scm:53-74)
    double next = __dadd_rn(cur, dlt);
    if (ah >= 0 && lead > cur && lead <= __dadd_rn(next, buf)) next = __dsub_rn(lead, buf);
    if (st & ST_ARRIVE) {   // route-step/arrive-at$  route.scm:76-90
      const double target = p.coldf[a].arrive;
      if (!(next >= target)) return next;
      const bool back = target < cur;
      // (atomicAnd / atomicOr: a follower may be flagging this actor as a ghost right now)
      atomicAnd(&p.hot[a].st, ~((3u << ST_RS_SHIFT) | ST_ARRIVE));
      const uint32_t st_before = atomicOr(&p.hot[a].st, (GRIDDY_ROUTE_DONE << ST_RS_SHIFT) | (back ? ST_MOVED : 0u));
      if (back) {   // jumped back past other residents: re-insert by key (the lane's occupancy does not change)
        const int32_t lane = p.cold[a].container;
        p.cold[a].mv_old = lane;
        if (!(st_before & ST_GHOST)) push_leaver(p, tick, a, lane, -1, sink);
        p.cold[a].inbox_next = enter_lane(p, a, lane, target, sink);
      }
      return target;
    }
    // route-step/turn-onto$  route.scm:92-135
    if (!(next >= 1.0)) return next;
    const int32_t tj = p.hot[a].tj;   // junction-full?  route.scm:98-108 (k_advance has usually answered this)
    if (tj >= 0 && p.occ[tj & ~TJ_REMOTE] > 0) return cur;   // waiting
    return turn_onto(p, a, tick, st, cur, pos_old, sink);
  }

  if (rs == GRIDDY_ROUTE_NONE && head == HEAD_NONE) return cur;   // do-nothing$  simulate.scm:70-72
  if (rs == GRIDDY_ROUTE_NONE && head == HEAD_SLEEP) {            // sleep$  simulate.scm:74-78
    const int32_t t = p.hot[a].tj;
    if (t == 0) {
      const int32_t ac = p.cold[a].agenda_cur + 1;
      int32_t aux;
      const uint32_t h = load_head(p, a, ac, &aux);
      p.cold[a].agenda_cur = ac;
      p.hot[a].tj = aux;
      atomicAnd(&p.hot[a].st, ~(3u << ST_HEAD_SHIFT));
      atomicOr(&p.hot[a].st, h << ST_HEAD_SHIFT);
    } else {
      p.hot[a].tj = t - 1;
    }
    return cur;
  }
  if (head != HEAD_TRAVEL) {   // simulate.scm:105-111 has no clause for these
    dev_error(p, DEV_ERR_NO_CLAUSE);
    return cur;
  }

  const int32_t item = p.cold[a].agenda_cur;
  if (rs == GRIDDY_ROUTE_NONE) {   // begin-route$  simulate.scm:80-85
    if (st & ST_ONROAD) { dev_error(p, DEV_ERR_BEGIN_ON_ROAD); return cur; }
    const int32_t seg = p.cold[a].container;
    const int32_t init_lane = outer_lane(p, seg, st & ST_SIDE);
    // the destination was resolved at upload time (off-road->on-road of the travel-to item, location.scm:49-57)
    const AgendaItem ai = p.agenda[item];
    const int32_t dest_lane = ai.dlane;
    if (init_lane < 0 || dest_lane < 0) { dev_error(p, DEV_ERR_NO_LANE); return cur; }
    const LaneRec init_it = p.lane[init_lane];
    // off-road->on-road  location.scm:49-57
    const double init_pos = meta_dir(init_it.meta) ? __dsub_rn(1.0, cur) : cur;
    const double dest_pos = ai.dpos;
    int32_t rc = p.route_off ? (int32_t)p.route_off[ai.rid] : 0;
    Cold k = p.cold[a];
    k.dest_lane = dest_lane;
    k.dest_from_j = p.route_off ? (int32_t)p.route_off[ai.rid + 1] : ai.dfrom;
    k.dest_to_j = ai.dto;
    const int32_t hop = route_head_on(p, init_lane, init_it, k, &rc);
    if (hop == -2) { dev_error(p, DEV_ERR_NO_ROUTE); return cur; }   // find-route found no path (a* throws 'no-path)
    const double len = init_it.len;
    p.coldf[a].arrive = dest_pos;
    p.hot[a].tj = junction_after(p, init_it, hop);
    p.hot[a].delta = __dmul_rn(k.speed_dt, __ddiv_rn(1.0, len));
    p.hot[a].buf = __ddiv_rn(p.two_size, len);
    p.hot[a].st = (st & ~((3u << ST_RS_SHIFT) | ST_SIDE | ST_ARRIVE)) | (GRIDDY_ROUTE_SOME << ST_RS_SHIFT) |
                  ST_ONROAD | ST_MOVED | ST_REDRAW | (hop == HOP_ARRIVE ? ST_ARRIVE : 0u);   // (off road: nobody flags it)
    k.hop = hop;
    k.container = init_lane;
    k.mv_old = ~seg;
    k.route_cur = rc;
    k.inbox_next = enter_lane(p, a, init_lane, init_pos, sink);
    p.cold[a] = k;
    return init_pos;
  }

  // end-route$  simulate.scm:87-91   (rs == GRIDDY_ROUTE_DONE)
  const int32_t lane = p.cold[a].container;
  const LaneRec lr = p.lane[lane];
  if (meta_kind(lr.meta) != GRIDDY_SEGMENT_LANE) { dev_error(p, DEV_ERR_END_ON_JLANE); return cur; }
  const int32_t seg = lr.link;
  const uint32_t dir = meta_dir(lr.meta);
  const int32_t ac = item + 1;
  int32_t aux;
  const uint32_t h = load_head(p, a, ac, &aux);
  p.cold[a].agenda_cur = ac;
  p.hot[a].tj = aux;
  p.cold[a].container = seg;
  p.cold[a].mv_old = lane;
  // landing stamp: where this actor sits in the segment's consed list (core.scm:169-171)
  p.coldf[a].land_tick = tick + 1;
  p.coldf[a].land_lane = lane;
  p.coldf[a].land_pos = cur;
  const uint32_t clear = (3u << ST_RS_SHIFT) | (3u << ST_HEAD_SHIFT) | ST_ONROAD | ST_SIDE | ST_CLASS_P | ST_ARRIVE;
  const uint32_t set = (h << ST_HEAD_SHIFT) | (dir ? ST_SIDE : 0u) | (lane > seg ? ST_CLASS_P : 0u) | ST_MOVED | ST_REDRAW;
  atomicAnd(&p.hot[a].st, ~clear);   // (a follower may be flagging this actor as a ghost right now)
  if (!(atomicOr(&p.hot[a].st, set) & ST_GHOST)) push_leaver(p, tick, a, lane, -1, sink);   // (a segment lane: no junction)
  return dir ? __dsub_rn(1.0, cur) : cur;   // on-road->off-road  location.scm:41-47
}

// k_advance: the common cases of advance$ for every slot; everything else is deferred to k_slow.
// ADV_UNROLL slots per thread, strided by the block so that every load is coalesced.  The loads of all
// ADV_UNROLL actors are issued before the first use (one slot per thread leaves too few bytes in
// flight to cover HBM latency): first the hot record (two 16-byte loads of one sector) and the pos-param,
// then the leaders' pos-params (a gather that the slot order keeps inside the same or a neighbouring line).
// Handled here: an actor on its route that has no (arrive-at _) head and either stays on its lane
// or waits at its end for a full junction (the junction-occupancy table is small enough to live in
// L2); a sleeper whose time has not run out; an idle actor.  These touch only the hot
// arrays.  Any other actor — a few percent — would drag its whole warp through chains of dependent
// gathers, so its slot is appended to the slow list instead (block-aggregated: one global atomic per
// block) and k_slow runs the full advance$ for it with one thread per listed actor.
// Measured on B200 (profiles/r2/variants.md): 2 slots per thread x 128 threads (40 registers) beats 256 and 512
// threads (the block's two barriers wait for fewer warps) and 1 or 4 slots per thread.
#ifndef ADV_UNROLL
#define ADV_UNROLL 2
#endif
#ifndef ADV_LD256   // the hot record as one 256-bit load (1), with L2 evict-first priority (2)
#define ADV_LD256 0
#endif
#ifndef ADV_THREADS_N
#define ADV_THREADS_N 128
#endif
constexpr int ADV_THREADS = ADV_THREADS_N;
#ifdef ADV_MIN_BLOCKS
#define ADV_BOUNDS __launch_bounds__(ADV_THREADS, ADV_MIN_BLOCKS)
#else
#define ADV_BOUNDS __launch_bounds__(ADV_THREADS)
#endif

// On a partitioned world the first halo_blocks blocks send the halo instead (multi_gpu.cuh, halo_pack_block) and, with
// one process per GPU, the last unpack_blocks blocks take the neighbours' halo in (halo_unpack_block).
__global__ void ADV_BOUNDS k_advance(Params p, int tick, int halo_blocks, int unpack_blocks) {
  __shared__ int4 s_list[ADV_THREADS * ADV_UNROLL];
  __shared__ double s_pos[ADV_THREADS * ADV_UNROLL];
  __shared__ int32_t s_n, s_base;
  grid_dependency_wait();
  TRACE(p, tick, TR_ADVANCE);
  const int par = tick & 1;
  int32_t* counters = p.counters + CNT_STRIDE * par;
  if (blockIdx.x == 0 && threadIdx.x < CNT_STRIDE) p.counters[CNT_STRIDE * (par ^ 1) + threadIdx.x] = 0;
  if ((int)blockIdx.x < halo_blocks) {
    halo_pack_block(p, tick, halo_blocks);
    return;
  }
  if ((int)blockIdx.x >= (int)gridDim.x - unpack_blocks) {
    halo_unpack_block(p, tick, (int)blockIdx.x - ((int)gridDim.x - unpack_blocks), unpack_blocks);
    return;
  }
  const int n = p.scal[SC_NSLOTS];
  const int base = ((int)blockIdx.x - halo_blocks) * (ADV_THREADS * ADV_UNROLL) + threadIdx.x;
  if (base - (int)threadIdx.x >= n) return;
  if (threadIdx.x == 0) s_n = 0;   // (the barrier after the first round of loads orders this before any append)
  const double* __restrict__ pos_old = p.pos[par];
  double* __restrict__ pos_new = p.pos[par ^ 1];
  const int4* __restrict__ hot4 = (const int4*)p.hot;

  uint32_t st[ADV_UNROLL];
  int32_t ah[ADV_UNROLL], tj[ADV_UNROLL], occ[ADV_UNROLL];
  double cur[ADV_UNROLL], dlt[ADV_UNROLL], buf[ADV_UNROLL], lead[ADV_UNROLL];
#pragma unroll
  for (int k = 0; k < ADV_UNROLL; ++k) {
    const int a = base + k * ADV_THREADS;
    const bool in = a < n;
    int4 h0 = make_int4(0, -1, -1, -1), h1 = make_int4(0, 0, 0, 0);   // st, ahead, behind, tj; delta, buf (one sector)
#if ADV_LD256 == 0
    if (in) { h0 = hot4[2 * (size_t)a]; h1 = hot4[2 * (size_t)a + 1]; }
#elif ADV_LD256 == 1
    if (in) asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(h0.x), "=r"(h0.y), "=r"(h0.z), "=r"(h0.w),
                         "=r"(h1.x), "=r"(h1.y), "=r"(h1.z), "=r"(h1.w) : "l"(hot4 + 2 * (size_t)a));
#else
    if (in) asm volatile("ld.global.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(h0.x), "=r"(h0.y), "=r"(h0.z), "=r"(h0.w),
                         "=r"(h1.x), "=r"(h1.y), "=r"(h1.z), "=r"(h1.w) : "l"(hot4 + 2 * (size_t)a));
#endif
    st[k] = (uint32_t)h0.x;
    ah[k] = h0.y;
    tj[k] = h0.w;
    dlt[k] = __hiloint2double(h1.y, h1.x);
    buf[k] = __hiloint2double(h1.w, h1.z);
    cur[k] = in ? pos_old[a] : 0.0;
    s_pos[k * ADV_THREADS + threadIdx.x] = cur[k];   // the block's pos-params: most leaders are among them
  }
  __syncthreads();
  // second round of loads, all independent of each other: the leader's pos-param — from shared memory when the
  // leader is one of this block's slots (after a reorder nearly always), else a gather — and, for an actor that can
  // reach its lane's end this tick, the occupancy of the junction ahead (clamping behind a leader only shortens the step)
  const int tile0 = base - (int)threadIdx.x;
#pragma unroll
  for (int k = 0; k < ADV_UNROLL; ++k) {
    const bool on_road = (st[k] & (ST_ALIVE | ST_ONROAD)) == (ST_ALIVE | ST_ONROAD);
    if (!on_road) ah[k] = -1;   // an off-road actor's ahead is stale
    const unsigned rel = (unsigned)(ah[k] - tile0);
    lead[k] = ah[k] < 0 ? 0.0 : (rel < (unsigned)(ADV_THREADS * ADV_UNROLL) ? s_pos[rel] : pos_old[ah[k]]);
    const bool may_turn = on_road && ((st[k] >> ST_RS_SHIFT) & 3) == GRIDDY_ROUTE_SOME && !(st[k] & ST_ARRIVE) &&
                          tj[k] >= 0 && tj[k] < TJ_REMOTE && __dadd_rn(cur[k], dlt[k]) >= 1.0;
    occ[k] = may_turn ? p.occ[tj[k]] : 0;
  }
#pragma unroll
  for (int k = 0; k < ADV_UNROLL; ++k) {
    const int a = base + k * ADV_THREADS;
    if (!(st[k] & ST_ALIVE)) continue;
    const int rs = (st[k] >> ST_RS_SHIFT) & 3, head = (st[k] >> ST_HEAD_SHIFT) & 3;
    bool slow = ah[k] >= 0 && lead[k] == cur[k];   // equal key ahead: a ghost to remove
    bool turn = false;                              // the only thing left to do is route-step/turn-onto$
    double np = cur[k];
    if (!slow) {
      if (rs == GRIDDY_ROUTE_SOME && head == HEAD_TRAVEL) {
        // get-pos-param-next on the current lane (route.scm:53-74)
        double next = __dadd_rn(cur[k], dlt[k]);
        if (ah[k] >= 0 && lead[k] > cur[k] && lead[k] <= __dadd_rn(next, buf[k])) next = __dsub_rn(lead[k], buf[k]);
        if (st[k] & ST_ARRIVE) slow = true;
        else if (!(next >= 1.0)) np = next;
        else if (tj[k] >= TJ_REMOTE) slow = true;                  // the junction's occupancy is still on its way (halo)
        else if (tj[k] >= 0 && occ[k] > 0) np = cur[k];         // waiting for a full junction  route.scm:98-111
        else slow = turn = true;                                   // turns onto the next lane
      } else if (rs == GRIDDY_ROUTE_NONE && head == HEAD_SLEEP) {   // sleep$, time left  simulate.scm:77
        if (tj[k] == 0) slow = true;
        else p.hot[a].tj = tj[k] - 1;
      } else if (!(rs == GRIDDY_ROUTE_NONE && head == HEAD_NONE)) {   // not do-nothing$ either
        slow = true;
      }
    }
    if (slow)   // ~slot: k_slow goes straight to turn_onto; st and pos-param ride along (two gathers less there)
      s_list[atomicAdd(&s_n, 1)] = make_int4(turn ? ~a : a, (int)st[k], __double2loint(cur[k]), __double2hiint(cur[k]));
    else pos_new[a] = np;
  }
  __syncthreads();
  const int ns = s_n;
  if (ns == 0) return;
  if (threadIdx.x == 0) s_base = atomicAdd(&counters[CNT_SLOW], ns);
  __syncthreads();
  for (int i = threadIdx.x; i < ns; i += ADV_THREADS) p.slow[s_base + i] = s_list[i];
}

// k_slow: the full advance$ for the actors k_advance deferred, one thread per actor, so that the
// chains of dependent gathers of 32 such actors overlap instead of stalling 31 idle lanes.
__global__ void SLOW_BOUNDS k_slow(Params p, int tick) {
  __shared__ int4 s_left[LEAVER_STAGE];
  __shared__ int32_t s_entered[SLOW_THREADS];
  __shared__ int32_t s_nl, s_ne, s_base_l, s_base_e;
  grid_dependency_wait();
  TRACE(p, tick, TR_SLOW);
  const int par = tick & 1;
  int32_t* counters = p.counters + CNT_STRIDE * par;
  const int32_t n_slow = counters[CNT_SLOW];
  const double* __restrict__ pos_old = p.pos[par];
  double* __restrict__ pos_new = p.pos[par ^ 1];
  const TouchSink sink{s_left, &s_nl, s_entered, &s_ne};
  // (handing the chunks out through a counter instead of this static split was measured 8 % slower)
  for (int32_t chunk = blockIdx.x * SLOW_THREADS; chunk < n_slow; chunk += gridDim.x * SLOW_THREADS) {
    if (threadIdx.x == 0) s_nl = s_ne = 0;
    __syncthreads();
    const int32_t i = chunk + threadIdx.x;
    if (i < n_slow) {
      const int4 ent = p.slow[i];
      const int32_t e = ent.x;
      const int a = e < 0 ? ~e : e;
      const uint32_t st = (uint32_t)ent.y;
      const double cur = __hiloint2double(ent.w, ent.z);
      if (e < 0) {
        pos_new[a] = turn_onto(p, a, tick, st, cur, pos_old, sink);
        l2_keep(&pos_new[a], 8);
      } else {
        const bool on_road = st & ST_ONROAD;
        const bool on_route = on_road && ((st >> ST_RS_SHIFT) & 3) == GRIDDY_ROUTE_SOME;
        const Hot h = p.hot[a];
        const int32_t ah = on_road ? h.ahead : -1;
        const double lead = ah >= 0 ? pos_old[ah] : 0.0;
        pos_new[a] = advance_actor(p, a, tick, st, ah, cur, on_route ? h.delta : 0.0, on_route ? h.buf : 0.0, lead, pos_old, sink);
      }
    }
    __syncthreads();
    const int nl = min(s_nl, LEAVER_STAGE), ne = s_ne;
    if (threadIdx.x == 0 && nl) s_base_l = atomicAdd(&counters[CNT_LEFT], nl);
    if (threadIdx.x == 1 && ne) s_base_e = atomicAdd(&counters[CNT_ENTERED], ne);
    __syncthreads();
    for (int k = threadIdx.x; k < nl; k += SLOW_THREADS) p.leavers[s_base_l + k] = s_left[k];
    if (threadIdx.x < ne) p.touched[s_base_e + threadIdx.x] = s_entered[threadIdx.x];
  }
  if (p.owner && !p.publish_late) emig_publish(p, tick);   // partitioned world: the block that finishes last raises the migrant flags
}

}  // namespace griddy
