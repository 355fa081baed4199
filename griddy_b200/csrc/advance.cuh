// advance.cuh — part of griddy_cuda.cu (one translation unit).  advance$ for every actor: k_advance (the common cases) and k_slow (everything else).
#pragma once

namespace griddy {

__device__ __forceinline__ void dev_error(const Params& p, int bit) { atomicOr(&p.scal[SC_ERR], bit); }

// location.scm:16-24 through the uploaded per-segment table
__device__ __forceinline__ int32_t outer_lane(const Params& p, int32_t seg, uint32_t side) {
  return side ? p.item[seg].link2 : p.item[seg].link;
}

// Dimension-ordered next hop on procedural grids (include/griddy_cuda.h, griddy_set_routing_grid):
// from segment lane `lane` towards the lane that leaves junction dest_from_j for junction dest_to_j.
__device__ int32_t grid_next_hop(const Params& p, int32_t lane, int32_t lane_to_j, int32_t dest_from_j, int32_t dest_to_j) {
  const int n = p.grid_n;
  const int32_t J = lane_to_j, E = dest_from_j;
  int h;
  if (J == E) {
    const int32_t T = dest_to_j;
    h = (T / n != E / n) ? (T / n > E / n ? 0 : 1) : (T % n > E % n ? 2 : 3);
  } else if (J / n != E / n) {
    h = E / n > J / n ? 0 : 1;
  } else {
    h = E % n > J % n ? 2 : 3;
  }
  return p.turn4[4 * (int64_t)lane + h];
}

// Head of the route after the actor has reached `lane` (route.scm:48-51 evaluated lazily); `it` is
// lane's record, k holds the route's destination.
__device__ int32_t route_head_on(const Params& p, int32_t lane, const Item& it, const Cold& k, int32_t item, int32_t* rc) {
  if (p.route_off) {
    const int64_t c = *rc;
    return c < p.route_off[item + 1] ? p.route_lane[c] : HOP_ARRIVE;
  }
  if (lane == k.dest_lane) return HOP_ARRIVE;
  if (it.kind == GRIDDY_JUNCTION_LANE) return it.link;
  const int32_t hop = grid_next_hop(p, lane, it.to_j, k.dest_from_j, k.dest_to_j);
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
  return tag_junction(p, (hop >= 0 && p.item[hop].kind == GRIDDY_JUNCTION_LANE) ? p.item[hop].link2 : -1);
}

// The same for the head `nhop` of a route that has just reached lane `it`.  With the grid routing table a
// segment lane's next hop is a junction lane of the junction the lane reaches (Item::to_j; checked by
// griddy_set_routing_grid) and a junction lane's next hop is a segment lane, so the hop's own record
// is not needed.
__device__ __forceinline__ int32_t junction_after(const Params& p, const Item& it, int32_t nhop) {
  if (p.route_off) return junction_of_hop(p, nhop);
  return tag_junction(p, (nhop >= 0 && it.kind == GRIDDY_SEGMENT_LANE) ? it.to_j : -1);
}

// New agenda head after a pop (actor.scm:28-31): its kind bits, and the sleep counter in *aux.
__device__ __forceinline__ uint32_t load_head(const Params& p, int32_t a, int32_t ac, int32_t* aux) {
  if (ac >= p.coldf[a].agenda_end) { *aux = 0; return HEAD_NONE; }
  const AgendaItem ai = p.agenda[ac];
  if (ai.kind == GRIDDY_SLEEP_FOR) { *aux = ai.sleep; return HEAD_SLEEP; }
  *aux = 0;
  return HEAD_TRAVEL;
}

// Lanes whose list must be rebuilt this tick are collected per block in shared memory and flushed to
// Params::touched with one global atomic per block (a same-address atomic per lane would serialise).
struct TouchSink {
  int32_t *left, *n_left;         // shared: lanes that lost an actor or hold a ghost (k_unlink), SLOW_THREADS entries
  int32_t *entered, *n_entered;   // shared: lanes that gained an actor (k_link), SLOW_THREADS entries
};
#ifdef SLOW_MIN_BLOCKS
#define SLOW_BOUNDS __launch_bounds__(SLOW_THREADS, SLOW_MIN_BLOCKS)
#else
#define SLOW_BOUNDS __launch_bounds__(SLOW_THREADS)
#endif
constexpr int SLOW_THREADS = 128;   // an actor leaves at most one lane (its own, ghost or not) and enters at most one

__device__ __forceinline__ void mark_left(const Params& p, int32_t lane, int tick, const TouchSink& sink) {
  if (atomicExch(&p.lane[lane].mark_out, tick + 1) != tick + 1) sink.left[atomicAdd(sink.n_left, 1)] = lane;
}

__device__ __forceinline__ void enter_lane(const Params& p, int a, int32_t lane, int tick, const TouchSink& sink) {
  p.cold[a].inbox_next = atomicExch(&p.lane[lane].inbox, a);
  if (atomicExch(&p.lane[lane].mark_in, tick + 1) != tick + 1) sink.entered[atomicAdd(sink.n_entered, 1)] = lane;
}

__device__ __forceinline__ void leave_lane(const Params& p, int a, int32_t lane, int tick, const TouchSink& sink) {
  p.cold[a].outbox_next = atomicExch(&p.lane[lane].outbox, a);
  mark_left(p, lane, tick, sink);
}

// ------------------------------------------------------------------------------------------------
// route-step/turn-onto$ (route.scm:92-135) for an actor that reaches the end of its lane and is not
// held by a full junction: carry the rest of the tick over into the target lane, pop the route.
// Reads one cold record of the actor and one record of the target lane; returns the new pos-param.
__device__ __forceinline__ double turn_onto(const Params& p, const int a, const int tick, const uint32_t st,
                                            const double cur, const double* __restrict__ pos_old, const TouchSink& sink) {
  Cold k = p.cold[a];
  const ColdF f = p.coldf[a];
  const int32_t target = k.hop, lane = k.container;
  const Item it = p.item[target];
  const double time_spent = __ddiv_rn(__dsub_rn(1.0, cur), f.speed);
  const double time_remaining = __dsub_rn(p.dt, time_spent);
  const double tinv = __ddiv_rn(1.0, it.len);
  const double tbuf = __ddiv_rn(p.two_size, it.len);
  double tnext = __dmul_rn(__dmul_rn(f.speed, time_remaining), tinv);   // (+ 0 delta)
  int32_t x = p.lane[target].tail;
  const bool remote = x <= -2;   // multi-GPU: another rank owns the target lane
  if (remote) {                  // its smallest key in (0, ...] came with the halo (+inf: none)
    const double tlead = __ldcg(&p.halo_recv[tick & 1][-2 - x]);
    if (tlead <= __dadd_rn(tnext, tbuf)) tnext = __dsub_rn(tlead, tbuf);
  } else {
    while (x >= 0 && !(pos_old[x] > 0.0)) x = p.sa[x].ahead;   // smallest key in (0, ...]
    if (x >= 0) {
      const double tlead = pos_old[x];
      if (tlead <= __dadd_rn(tnext, tbuf)) tnext = __dsub_rn(tlead, tbuf);
    }
  }
  int32_t rc = p.route_off ? k.route_cur + 1 : 0;
  const int32_t nhop = route_head_on(p, target, it, k, k.agenda_cur, &rc);
  p.cold[a].route_cur = rc;
  p.cold[a].hop = nhop;
  p.cold[a].container = target;
  p.cold[a].mv_old = lane;
  p.tj[a] = junction_after(p, it, nhop);
  p.db[a] = make_double2(__dmul_rn(f.speed_dt, tinv), tbuf);
  p.sa[a].st = st | ST_MOVED | (nhop == HOP_ARRIVE ? ST_ARRIVE : 0u) | (remote ? ST_EMIG : 0u);
  if (remote) p.emig[atomicAdd(&p.counters[CNT_STRIDE * (tick & 1) + CNT_EMIG], 1)] = a;   // a few thousand per tick
  else enter_lane(p, a, target, tick, sink);
  leave_lane(p, a, lane, tick, sink);
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
    do {
      if (atomicExch(&p.cold[ah].dead_mark, tick + 1) != tick + 1)   // the first to notice hands it to k_unlink
        p.cold[ah].ghost_next = atomicExch(&p.lane[lane].ghosts, ah);
      ah = p.sa[ah].ahead;
      if (ah >= 0) lead = pos_old[ah];
    } while (ah >= 0 && lead == cur);
    mark_left(p, lane, tick, sink);
  }

  if (rs == GRIDDY_ROUTE_SOME && head == HEAD_TRAVEL) {   // advance-on-route$  route.scm:137-141
    // get-pos-param-next on the current lane (route.scm:53-74)
    double next = __dadd_rn(cur, dlt);
    if (ah >= 0 && lead > cur && lead <= __dadd_rn(next, buf)) next = __dsub_rn(lead, buf);
    if (st & ST_ARRIVE) {   // route-step/arrive-at$  route.scm:76-90
      const double target = p.coldf[a].arrive;
      if (!(next >= target)) return next;
      const bool back = target < cur;
      p.sa[a].st = (st & ~((3u << ST_RS_SHIFT) | ST_ARRIVE)) | (GRIDDY_ROUTE_DONE << ST_RS_SHIFT) | (back ? ST_MOVED : 0u);
      if (back) {   // jumped back past other residents: re-insert by key
        const int32_t lane = p.cold[a].container;
        p.cold[a].mv_old = lane;
        leave_lane(p, a, lane, tick, sink);
        enter_lane(p, a, lane, tick, sink);
      }
      return target;
    }
    // route-step/turn-onto$  route.scm:92-135
    if (!(next >= 1.0)) return next;
    const int32_t tj = p.tj[a];   // junction-full?  route.scm:98-108 (k_advance has usually answered this)
    if (tj >= 0 && p.occ[tj & ~TJ_REMOTE] > 0) return cur;   // waiting
    return turn_onto(p, a, tick, st, cur, pos_old, sink);
  }

  if (rs == GRIDDY_ROUTE_NONE && head == HEAD_NONE) return cur;   // do-nothing$  simulate.scm:70-72
  if (rs == GRIDDY_ROUTE_NONE && head == HEAD_SLEEP) {            // sleep$  simulate.scm:74-78
    const int32_t t = p.aux[a];
    if (t == 0) {
      const int32_t ac = p.cold[a].agenda_cur + 1;
      int32_t aux;
      const uint32_t h = load_head(p, a, ac, &aux);
      p.cold[a].agenda_cur = ac;
      p.aux[a] = aux;
      p.sa[a].st = (st & ~(3u << ST_HEAD_SHIFT)) | (h << ST_HEAD_SHIFT);
    } else {
      p.aux[a] = t - 1;
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
    const Item init_it = p.item[init_lane];
    // off-road->on-road  location.scm:49-57
    const double init_pos = init_it.dir ? __dsub_rn(1.0, cur) : cur;
    const double dest_pos = ai.dpos;
    int32_t rc = p.route_off ? (int32_t)p.route_off[item] : 0;
    Cold k;
    k.dest_lane = dest_lane;
    k.dest_from_j = ai.dfrom;
    k.dest_to_j = ai.dto;
    const int32_t hop = route_head_on(p, init_lane, init_it, k, item, &rc);
    const double len = init_it.len;
    p.cold[a].route_cur = rc;
    p.cold[a].dest_lane = dest_lane;
    p.cold[a].dest_from_j = k.dest_from_j;
    p.cold[a].dest_to_j = k.dest_to_j;
    p.coldf[a].arrive = dest_pos;
    p.cold[a].hop = hop;
    p.tj[a] = junction_after(p, init_it, hop);
    p.db[a] = make_double2(__dmul_rn(p.coldf[a].speed_dt, __ddiv_rn(1.0, len)), __ddiv_rn(p.two_size, len));
    p.cold[a].container = init_lane;
    p.cold[a].mv_old = ~seg;
    p.sa[a].st = (st & ~((3u << ST_RS_SHIFT) | ST_SIDE | ST_ARRIVE)) | (GRIDDY_ROUTE_SOME << ST_RS_SHIFT) |
              ST_ONROAD | ST_MOVED | (hop == HOP_ARRIVE ? ST_ARRIVE : 0u);
    enter_lane(p, a, init_lane, tick, sink);
    return init_pos;
  }

  // end-route$  simulate.scm:87-91   (rs == GRIDDY_ROUTE_DONE)
  const int32_t lane = p.cold[a].container;
  if (p.item[lane].kind != GRIDDY_SEGMENT_LANE) { dev_error(p, DEV_ERR_END_ON_JLANE); return cur; }
  const int32_t seg = p.item[lane].link;
  const uint32_t dir = p.item[lane].dir;
  const int32_t ac = item + 1;
  int32_t aux;
  const uint32_t h = load_head(p, a, ac, &aux);
  p.cold[a].agenda_cur = ac;
  p.aux[a] = aux;
  p.cold[a].container = seg;
  p.cold[a].mv_old = lane;
  // landing stamp: where this actor sits in the segment's consed list (core.scm:169-171)
  p.coldf[a].land_tick = tick + 1;
  p.coldf[a].land_lane = lane;
  p.coldf[a].land_pos = cur;
  uint32_t ns = st & ~((3u << ST_RS_SHIFT) | (3u << ST_HEAD_SHIFT) | ST_ONROAD | ST_SIDE | ST_CLASS_P | ST_ARRIVE);
  ns |= (h << ST_HEAD_SHIFT) | (dir ? ST_SIDE : 0u) | (lane > seg ? ST_CLASS_P : 0u) | ST_MOVED;
  p.sa[a].st = ns;
  leave_lane(p, a, lane, tick, sink);
  return dir ? __dsub_rn(1.0, cur) : cur;   // on-road->off-road  location.scm:41-47
}

// k_advance: the common cases of advance$ for every slot; everything else is deferred to k_slow.
// ADV_UNROLL slots per thread, strided by the block so that every load is coalesced.  The loads of all
// ADV_UNROLL actors are issued before the first use (one slot per thread leaves too few bytes in
// flight to cover HBM latency): first st / ahead / pos, then the leaders' pos-params (a gather that
// the slot order keeps inside the same or a neighbouring cache line) with delta / buf.
// Handled here: an actor on its route that has no (arrive-at _) head and either stays on its lane
// or waits at its end for a full junction (the junction-occupancy table is small enough to live in
// L2); a sleeper whose time has not run out; an idle actor.  These touch only the hot
// arrays.  Any other actor — a few percent — would drag its whole warp through chains of dependent
// gathers, so its slot is appended to the slow list instead (block-aggregated: one global atomic per
// block) and k_slow runs the full advance$ for it with one thread per listed actor.
// Measured on B200 (profiles/r1_variants.md): 2 slots per thread = 40 registers = 6 blocks per SM beats
// 3 slots (58 registers with the slow list's payload, 4 blocks) and 3 slots capped at 48 registers.
#ifndef ADV_UNROLL
#define ADV_UNROLL 2
#endif
#ifndef ADV_THREADS_N
#define ADV_THREADS_N 256
#endif
constexpr int ADV_THREADS = ADV_THREADS_N;
#ifndef SLOW_PAYLOAD
#define SLOW_PAYLOAD 1
#endif
#ifdef ADV_MIN_BLOCKS
#define ADV_BOUNDS __launch_bounds__(ADV_THREADS, ADV_MIN_BLOCKS)
#else
#define ADV_BOUNDS __launch_bounds__(ADV_THREADS)
#endif

__global__ void ADV_BOUNDS k_advance(Params p, int tick) {
  __shared__ int4 s_list[ADV_THREADS * ADV_UNROLL];
  __shared__ int32_t s_n, s_base;
  grid_dependency_wait();
  TRACE(p, tick, TR_ADVANCE);
  const int par = tick & 1;
  int32_t* counters = p.counters + CNT_STRIDE * par;
  if (blockIdx.x == 0 && threadIdx.x < CNT_STRIDE) p.counters[CNT_STRIDE * (par ^ 1) + threadIdx.x] = 0;
  const int n = p.scal[SC_NSLOTS];
  const int base = blockIdx.x * (ADV_THREADS * ADV_UNROLL) + threadIdx.x;
  if (base - (int)threadIdx.x >= n) return;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const double* __restrict__ pos_old = p.pos[par];
  double* __restrict__ pos_new = p.pos[par ^ 1];

  uint32_t st[ADV_UNROLL];
  int32_t ah[ADV_UNROLL], slp[ADV_UNROLL], tj[ADV_UNROLL];
  double cur[ADV_UNROLL], dlt[ADV_UNROLL], buf[ADV_UNROLL], lead[ADV_UNROLL];
#pragma unroll
  for (int k = 0; k < ADV_UNROLL; ++k) {
    const int a = base + k * ADV_THREADS;
    const bool in = a < n;
    const SA sa = in ? p.sa[a] : SA{0u, -1};
    st[k] = sa.st;
    ah[k] = sa.ahead;
    cur[k] = in ? pos_old[a] : 0.0;
  }
#pragma unroll
  for (int k = 0; k < ADV_UNROLL; ++k) {
    const int a = base + k * ADV_THREADS;
    const bool alive = st[k] & ST_ALIVE;
    const bool on_road = alive && (st[k] & ST_ONROAD);
    if (!on_road) ah[k] = -1;   // an off-road actor's ahead is stale
    const int rs = (st[k] >> ST_RS_SHIFT) & 3, head = (st[k] >> ST_HEAD_SHIFT) & 3;
    const bool on_route = on_road && rs == GRIDDY_ROUTE_SOME;
    lead[k] = ah[k] >= 0 ? pos_old[ah[k]] : 0.0;
    const double2 db = on_route ? p.db[a] : make_double2(0.0, 0.0);
    dlt[k] = db.x;
    buf[k] = db.y;
    // (loading tj only at the lane's end was measured slower: on a congested grid every queue head
    // gets there every tick, so most sectors are fetched anyway, now behind a dependent load)
    tj[k] = (on_route && !(st[k] & ST_ARRIVE)) ? p.tj[a] : -1;
    slp[k] = (alive && rs == GRIDDY_ROUTE_NONE && head == HEAD_SLEEP) ? p.aux[a] : 0;
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
        else if (tj[k] >= 0 && p.occ[tj[k]] > 0) np = cur[k];   // waiting for a full junction  route.scm:98-111
        else slow = turn = true;                                   // turns onto the next lane
      } else if (rs == GRIDDY_ROUTE_NONE && head == HEAD_SLEEP) {   // sleep$, time left  simulate.scm:77
        if (slp[k] == 0) slow = true;
        else p.aux[a] = slp[k] - 1;
      } else if (!(rs == GRIDDY_ROUTE_NONE && head == HEAD_NONE)) {   // not do-nothing$ either
        slow = true;
      }
    }
    if (slow)   // ~slot: k_slow goes straight to turn_onto; st and pos-param ride along (two gathers less there)
#if SLOW_PAYLOAD
      s_list[atomicAdd(&s_n, 1)] = make_int4(turn ? ~a : a, (int)st[k], __double2loint(cur[k]), __double2hiint(cur[k]));
#else
      s_list[atomicAdd(&s_n, 1)] = make_int4(turn ? ~a : a, 0, 0, 0);
#endif
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
  __shared__ int32_t s_left[SLOW_THREADS], s_entered[SLOW_THREADS];
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
#if SLOW_PAYLOAD
      const uint32_t st = (uint32_t)ent.y;
      const double cur = __hiloint2double(ent.w, ent.z);
#else
      const uint32_t st = p.sa[a].st;
      const double cur = pos_old[a];
#endif
      if (e < 0) {
        pos_new[a] = turn_onto(p, a, tick, st, cur, pos_old, sink);
      } else {
      const bool on_road = st & ST_ONROAD;
      const bool on_route = on_road && ((st >> ST_RS_SHIFT) & 3) == GRIDDY_ROUTE_SOME;
      const int32_t ah = on_road ? p.sa[a].ahead : -1;
      const double lead = ah >= 0 ? pos_old[ah] : 0.0;
      const double2 db = on_route ? p.db[a] : make_double2(0.0, 0.0);
      pos_new[a] = advance_actor(p, a, tick, st, ah, cur, db.x, db.y, lead, pos_old, sink);
      }
    }
    __syncthreads();
    const int nl = s_nl, ne = s_ne;
    if (threadIdx.x == 0 && nl) s_base_l = atomicAdd(&counters[CNT_LEFT], nl);
    if (threadIdx.x == 1 && ne) s_base_e = atomicAdd(&counters[CNT_ENTERED], ne);
    __syncthreads();
    if (threadIdx.x < nl) p.touched[s_base_l + threadIdx.x] = s_left[threadIdx.x];
    if (threadIdx.x < ne) p.touched[p.cap + s_base_e + threadIdx.x] = s_entered[threadIdx.x];
  }
}

}  // namespace griddy
