// griddy_cuda.cu — the per-tick actor update of griddy/simulate as sm_100a CUDA, behind the C ABI
// of include/griddy_cuda.h.  No CPU fallback, no torch types, no third-party kernels.
//
// What one tick computes (reference simulate.scm:141-153): every actor of the OLD world is mapped
// to an actor of a NEW world by `advance$` (simulate.scm:93-111).  An actor reads only the old
// world — its own state, the nearest actor ahead on its lane, and on a lane change the target
// lane's rearmost positive pos-param and the target junction's occupancy (route.scm:53-135) — so
// the update is data-parallel.  Processing order matters in exactly two places: an equal-key insert
// on a lane replaces the earlier actor (core.scm:173-178), and a segment's off-road list is consed
// (core.scm:169-171).  Both are reproduced from local data, without a global sort (see below).
//
// Data layout (HBM, structure of arrays, index = actor id; items by registration index)
//   st         u32  alive | on_road | side | route status | agenda head kind | moved | stamp class
//   container  i32  lane (on road) or segment (off road)
//   pos[2]     f64  pos-param, ping-pong by tick parity: followers read the old buffer
//   ahead      i32  next actor ahead on the same lane (ascending pos-param), -1 at the front
//   delta,buf  f64  per-tick advance and tailgate buffer on the current lane, cached at lane entry
//                   (route.scm:62-66: fl(speed*dt)*fl(1/len) and 10/len)
//   aux        i32  sleeping: remaining time; travelling: head route step (lane, or -1 arrive-at)
//   lane_tail  i32  per lane: rearmost actor;  occ i32 per junction: actors on its junction lanes
// Per-lane order is kept as linked lists that change only where an actor entered or left a lane
// this tick; nothing is O(lanes) per tick (lanes outnumber actors 1.5:1 to 4:1 on the grids).
//
// Kernels per tick
//   k_advance  one thread per actor: agenda dispatch + motion, all handlers fused.  Coalesced SoA
//              reads, one gather for the leader's pos-param.  Actors that change lane or
//              road-side are appended to the movers list and pushed on their target lane's inbox
//              (warp-aggregated atomics); their lanes are marked touched.
//   k_relink   one thread per touched lane: merges the lane's surviving residents (already in
//              order: residents never overtake) with its sorted entrants, drops equal-key losers
//              exactly as bbtree-set would, rewrites the ahead pointers.
//   k_settle   one thread per mover: publishes its new ahead pointer, applies junction occupancy
//              deltas.
//
// Processing order without a global sort.  "a is processed before b" (world.scm:24-30) is decided
// from their OLD locations: different containers -> higher registration index first; same lane ->
// higher pos-param first; same segment -> list order, which is encoded per actor by a landing stamp
// (tick, lane, pos-param, before/after class) because a segment's list reverses every tick.
//
// Equal keys among residents.  Residents of a lane keep their relative order (a follower is clamped
// to leader.old - buf, route.scm:68-74), but only weakly: in fp64 two residents an ulp apart can
// round to the same new key, and bbtree-set then drops the one visited first — the leader, since a
// lane is walked in descending pos-param.  Such ties are always list-adjacent.  Lanes that were
// touched anyway resolve them in k_relink; elsewhere they are resolved one tick late at no cost: an
// actor whose list follower holds an equal key is a GHOST (it is not part of the world any more).
// Every on-road actor reads its leader's key each tick, sees the tie, marks the ghost dead and has
// the lane relinked; whatever the ghost itself did this tick is discarded by k_relink / k_settle.
// Nothing else can observe a ghost: queries reach the follower first and see the same key, and the
// junction it may stand on is occupied by the follower as well.  Downloads apply the same rule.
//
// fp64 parity: every operation of route.scm:53-135 is issued with round-to-nearest intrinsics
// (__dadd_rn / __dmul_rn / __ddiv_rn), which the compiler never contracts into FMAs.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cooperative_groups.h>

#include "griddy_cuda.h"

namespace cg = cooperative_groups;

namespace {

constexpr uint32_t ST_ALIVE = 1u, ST_ONROAD = 2u, ST_SIDE = 4u;
constexpr int ST_RS_SHIFT = 3;    // 2 bits: GRIDDY_ROUTE_*
constexpr int ST_HEAD_SHIFT = 5;  // 2 bits: 0 no head, 1 sleep-for, 2 travel-to
constexpr uint32_t ST_MOVED = 128u;    // changed lane / road side this tick (cleared by k_settle)
constexpr uint32_t ST_CLASS_P = 256u;  // landing stamp: came from a lane visited before the segment
constexpr int HEAD_NONE = 0, HEAD_SLEEP = 1, HEAD_TRAVEL = 2;
constexpr int HOP_ARRIVE = -1;

constexpr int DEV_ERR_BEGIN_ON_ROAD = 1, DEV_ERR_NO_CLAUSE = 2, DEV_ERR_END_ON_JLANE = 4,
              DEV_ERR_NO_LANE = 8, DEV_ERR_NO_ROUTE = 16;

struct Params {
  // network
  int64_t n_items;
  const uint8_t* kind;
  const double* lane_len;
  const uint8_t* lane_dir;
  const int32_t *lane_segment, *lane_out, *lane_junction, *outer_forw, *outer_back;
  int32_t grid_n;
  const int32_t *lane_from_j, *lane_to_j, *turn4;
  int32_t *lane_tail, *occ, *lane_mark, *inbox_head;
  // actors
  int64_t n;
  uint32_t* st;
  int32_t* container;
  double* pos[2];
  int32_t* ahead;
  double *delta, *buf;
  int32_t *aux, *agenda_cur, *route_cur;
  double* arrive;
  const double *speed, *speed_dt;
  const int64_t* agenda_off;
  const uint8_t* item_kind;
  const int32_t *item_sleep, *item_seg;
  const uint8_t* item_side;
  const double* item_pos;
  const int64_t* route_off;   // null: grid routing table
  const int32_t* route_lane;
  int32_t *land_tick, *land_lane;
  double* land_pos;
  // per-tick scratch
  int32_t *mv_old, *ahead_new, *inbox_next;
  int32_t* dead_mark;   // tick+1 when the actor was found to be a ghost during `tick`
  int32_t *movers, *touched;
  int32_t* counters;   // [2 parities][2]: movers, touched lanes
  int32_t* err;
  double dt, two_size;
};

// Claim-then-broadcast: one atomic per converged group instead of one per thread.
__device__ __forceinline__ int32_t agg_append(int32_t* counter) {
  cg::coalesced_group g = cg::coalesced_threads();
  int32_t base = 0;
  if (g.thread_rank() == 0) base = atomicAdd(counter, (int32_t)g.size());
  return g.shfl(base, 0) + (int32_t)g.thread_rank();
}

// location.scm:16-24 through the uploaded per-segment table
__device__ __forceinline__ int32_t outer_lane(const Params& p, int32_t seg, uint32_t side) {
  return side ? p.outer_back[seg] : p.outer_forw[seg];
}

// Dimension-ordered next hop on procedural grids (include/griddy_cuda.h, griddy_set_routing_grid).
__device__ int32_t grid_next_hop(const Params& p, int32_t lane, int32_t dest) {
  const int n = p.grid_n;
  const int32_t J = p.lane_to_j[lane], E = p.lane_from_j[dest];
  int h;
  if (J == E) {
    const int32_t T = p.lane_to_j[dest];
    h = (T / n != E / n) ? (T / n > E / n ? 0 : 1) : (T % n > E % n ? 2 : 3);
  } else if (J / n != E / n) {
    h = E / n > J / n ? 0 : 1;
  } else {
    h = E % n > J % n ? 2 : 3;
  }
  return p.turn4[4 * (int64_t)lane + h];
}

// Head of the route after the actor has reached `lane` (route.scm:48-51 evaluated lazily).
__device__ int32_t route_head_on(const Params& p, int a, int32_t lane, int32_t item, int32_t* rc) {
  if (p.route_off) {
    const int64_t c = *rc;
    return c < p.route_off[item + 1] ? p.route_lane[c] : HOP_ARRIVE;
  }
  const int32_t dest = outer_lane(p, p.item_seg[item], p.item_side[item]);
  if (lane == dest) return HOP_ARRIVE;
  if (p.kind[lane] == GRIDDY_JUNCTION_LANE) return p.lane_out[lane];
  const int32_t hop = grid_next_hop(p, lane, dest);
  if (hop < 0) atomicOr(p.err, DEV_ERR_NO_ROUTE);
  return hop;
}

// New agenda head after a pop (actor.scm:28-31): its kind bits, and the sleep counter in *aux.
__device__ __forceinline__ uint32_t load_head(const Params& p, int a, int32_t ac, int32_t* aux) {
  if (ac >= p.agenda_off[a + 1]) { *aux = 0; return HEAD_NONE; }
  if (p.item_kind[ac] == GRIDDY_SLEEP_FOR) { *aux = p.item_sleep[ac]; return HEAD_SLEEP; }
  *aux = 0;
  return HEAD_TRAVEL;
}

__device__ __forceinline__ void mark_touched(const Params& p, int32_t lane, int tick, int32_t* n_touched) {
  if (atomicExch(&p.lane_mark[lane], tick + 1) != tick + 1) p.touched[agg_append(n_touched)] = lane;
}

__device__ __forceinline__ void enter_lane(const Params& p, int a, int32_t lane, int tick, int32_t* n_touched) {
  p.inbox_next[a] = atomicExch(&p.inbox_head[lane], a);
  mark_touched(p, lane, tick, n_touched);
}

// ------------------------------------------------------------------------------------------------
// k_advance: advance$ for every actor (simulate.scm:93-111) with route.scm:53-141 inlined.
__global__ void __launch_bounds__(256) k_advance(Params p, int tick) {
  const int par = tick & 1;
  int32_t* counters = p.counters + 2 * par;
  if (blockIdx.x == 0 && threadIdx.x < 2) p.counters[2 * (par ^ 1) + threadIdx.x] = 0;
  const int64_t a64 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a64 >= p.n) return;
  const int a = (int)a64;
  const uint32_t st = p.st[a];
  if (!(st & ST_ALIVE)) return;
  const double* __restrict__ pos_old = p.pos[par];
  double* __restrict__ pos_new = p.pos[par ^ 1];
  const double cur = pos_old[a];
  const int head = (st >> ST_HEAD_SHIFT) & 3;
  const int rs = (st >> ST_RS_SHIFT) & 3;

  // nearest actor ahead on the lane (bbtree-find-min's only candidate, util.scm:152-162), skipping
  // ghosts: leaders that this actor replaced with an equal key at the end of the previous tick
  int32_t ah = -1;
  double lead = 0.0;
  if (st & ST_ONROAD) {
    ah = p.ahead[a];
    if (ah >= 0) {
      lead = pos_old[ah];
      if (lead == cur) {
        do {
          p.dead_mark[ah] = tick + 1;
          ah = p.ahead[ah];
          if (ah >= 0) lead = pos_old[ah];
        } while (ah >= 0 && lead == cur);
        mark_touched(p, p.container[a], tick, &counters[1]);
      }
    }
  }

  if (rs == GRIDDY_ROUTE_NONE && head == HEAD_NONE) {   // do-nothing$  simulate.scm:70-72
    pos_new[a] = cur;
    return;
  }
  if (rs == GRIDDY_ROUTE_NONE && head == HEAD_SLEEP) {   // sleep$  simulate.scm:74-78
    const int32_t t = p.aux[a];
    if (t == 0) {
      const int32_t ac = p.agenda_cur[a] + 1;
      int32_t aux;
      const uint32_t h = load_head(p, a, ac, &aux);
      p.agenda_cur[a] = ac;
      p.aux[a] = aux;
      p.st[a] = (st & ~(3u << ST_HEAD_SHIFT)) | (h << ST_HEAD_SHIFT);
    } else {
      p.aux[a] = t - 1;
    }
    pos_new[a] = cur;
    return;
  }
  if (head != HEAD_TRAVEL) {   // simulate.scm:105-111 has no clause for these
    atomicOr(p.err, DEV_ERR_NO_CLAUSE);
    pos_new[a] = cur;
    return;
  }
  const int32_t item = p.agenda_cur[a];

  if (rs == GRIDDY_ROUTE_NONE) {   // begin-route$  simulate.scm:80-85
    if (st & ST_ONROAD) { atomicOr(p.err, DEV_ERR_BEGIN_ON_ROAD); pos_new[a] = cur; return; }
    const int32_t seg = p.container[a];
    const int32_t init_lane = outer_lane(p, seg, st & ST_SIDE);
    const int32_t dest_lane = outer_lane(p, p.item_seg[item], p.item_side[item]);
    if (init_lane < 0 || dest_lane < 0) { atomicOr(p.err, DEV_ERR_NO_LANE); pos_new[a] = cur; return; }
    // off-road->on-road  location.scm:49-57
    const double init_pos = p.lane_dir[init_lane] ? __dsub_rn(1.0, cur) : cur;
    const double ipos = p.item_pos[item];
    const double dest_pos = p.lane_dir[dest_lane] ? __dsub_rn(1.0, ipos) : ipos;
    int32_t rc = p.route_off ? (int32_t)p.route_off[item] : 0;
    const int32_t hop = route_head_on(p, a, init_lane, item, &rc);
    const double len = p.lane_len[init_lane];
    p.route_cur[a] = rc;
    p.arrive[a] = dest_pos;
    p.aux[a] = hop;
    p.delta[a] = __dmul_rn(p.speed_dt[a], __ddiv_rn(1.0, len));
    p.buf[a] = __ddiv_rn(p.two_size, len);
    p.container[a] = init_lane;
    p.mv_old[a] = ~seg;
    p.st[a] = (st & ~((3u << ST_RS_SHIFT) | ST_SIDE)) | (GRIDDY_ROUTE_SOME << ST_RS_SHIFT) | ST_ONROAD | ST_MOVED;
    pos_new[a] = init_pos;
    p.movers[agg_append(&counters[0])] = a;
    enter_lane(p, a, init_lane, tick, &counters[1]);
    return;
  }

  const int32_t lane = p.container[a];
  if (rs == GRIDDY_ROUTE_DONE) {   // end-route$  simulate.scm:87-91
    if (p.kind[lane] != GRIDDY_SEGMENT_LANE) { atomicOr(p.err, DEV_ERR_END_ON_JLANE); pos_new[a] = cur; return; }
    const int32_t seg = p.lane_segment[lane];
    const uint32_t dir = p.lane_dir[lane];
    const int32_t ac = item + 1;
    int32_t aux;
    const uint32_t h = load_head(p, a, ac, &aux);
    p.agenda_cur[a] = ac;
    p.aux[a] = aux;
    p.container[a] = seg;
    p.mv_old[a] = lane;
    // landing stamp: where this actor sits in the segment's consed list (core.scm:169-171)
    p.land_tick[a] = tick + 1;
    p.land_lane[a] = lane;
    p.land_pos[a] = cur;
    uint32_t ns = st & ~((3u << ST_RS_SHIFT) | (3u << ST_HEAD_SHIFT) | ST_ONROAD | ST_SIDE | ST_CLASS_P);
    ns |= (h << ST_HEAD_SHIFT) | (dir ? ST_SIDE : 0u) | (lane > seg ? ST_CLASS_P : 0u) | ST_MOVED;
    p.st[a] = ns;
    pos_new[a] = dir ? __dsub_rn(1.0, cur) : cur;   // on-road->off-road  location.scm:41-47
    p.movers[agg_append(&counters[0])] = a;
    mark_touched(p, lane, tick, &counters[1]);
    return;
  }

  // advance-on-route$  route.scm:137-141.  get-pos-param-next on the current lane (route.scm:53-74)
  const double buf = p.buf[a];
  double next = __dadd_rn(cur, p.delta[a]);
  if (ah >= 0 && lead > cur && lead <= __dadd_rn(next, buf)) next = __dsub_rn(lead, buf);
  const int32_t hop = p.aux[a];
  if (hop == HOP_ARRIVE) {   // route-step/arrive-at$  route.scm:76-90
    const double target = p.arrive[a];
    if (next >= target) {
      p.st[a] = (st & ~(3u << ST_RS_SHIFT)) | (GRIDDY_ROUTE_DONE << ST_RS_SHIFT) | (target < cur ? ST_MOVED : 0u);
      pos_new[a] = target;
      if (target < cur) {   // jumped back past other residents: re-insert by key
        p.mv_old[a] = lane;
        p.movers[agg_append(&counters[0])] = a;
        enter_lane(p, a, lane, tick, &counters[1]);
      }
    } else {
      pos_new[a] = next;
    }
    return;
  }
  // route-step/turn-onto$  route.scm:92-135
  if (!(next >= 1.0)) { pos_new[a] = next; return; }
  const int32_t target = hop;
  if (p.kind[target] == GRIDDY_JUNCTION_LANE && p.occ[p.lane_junction[target]] > 0) {   // waiting
    pos_new[a] = cur;
    return;
  }
  const double speed = p.speed[a];
  const double time_spent = __ddiv_rn(__dsub_rn(1.0, cur), speed);
  const double time_remaining = __dsub_rn(p.dt, time_spent);
  const double tlen = p.lane_len[target];
  const double tinv = __ddiv_rn(1.0, tlen);
  const double tbuf = __ddiv_rn(p.two_size, tlen);
  double tnext = __dmul_rn(__dmul_rn(speed, time_remaining), tinv);   // (+ 0 delta)
  int32_t x = p.lane_tail[target];
  while (x >= 0 && !(pos_old[x] > 0.0)) x = p.ahead[x];   // smallest key in (0, ...]
  if (x >= 0) {
    const double lead = pos_old[x];
    if (lead <= __dadd_rn(tnext, tbuf)) tnext = __dsub_rn(lead, tbuf);
  }
  int32_t rc = p.route_cur[a] + 1;
  const int32_t nhop = route_head_on(p, a, target, item, &rc);
  p.route_cur[a] = rc;
  p.aux[a] = nhop;
  p.delta[a] = __dmul_rn(p.speed_dt[a], tinv);
  p.buf[a] = tbuf;
  p.container[a] = target;
  p.mv_old[a] = lane;
  p.st[a] = st | ST_MOVED;
  pos_new[a] = tnext;
  p.movers[agg_append(&counters[0])] = a;
  enter_lane(p, a, target, tick, &counters[1]);
  mark_touched(p, lane, tick, &counters[1]);
}

// ------------------------------------------------------------------------------------------------
// Processing order (world.scm:24-30) from old locations.

// earlier(a,b) among actors that landed on the same segment in the same tick
__device__ __forceinline__ bool landed_earlier(const Params& p, int a, int b, int32_t t_land) {
  const int32_t la = p.land_lane[a], lb = p.land_lane[b];
  if (t_land == 0) return la < lb;   // initial placement: land_lane holds the link! order
  if (la != lb) return la > lb;      // higher registration index comes first in static-items
  return p.land_pos[a] > p.land_pos[b];   // a lane yields its actors in descending pos-param
}

__device__ __forceinline__ bool stamp_less(const Params& p, int a, int b, uint32_t sta) {
  const int32_t ta = p.land_tick[a], tb = p.land_tick[b];
  if (ta != tb) return ta < tb;
  return (sta & ST_CLASS_P) ? landed_earlier(p, b, a, ta) : landed_earlier(p, a, b, ta);
}

// Is a before b in their segment's list at `tick`?  The list reverses every tick (core.scm:169-171
// conses while get-actors walks head to tail), so a stamp's sign alternates with tick parity.
__device__ bool offroad_before(const Params& p, int a, int b, int tick) {
  const uint32_t sta = p.st[a], stb = p.st[b];
  const bool nega = (((tick - p.land_tick[a]) & 1) != 0) == ((sta & ST_CLASS_P) != 0);
  const bool negb = (((tick - p.land_tick[b]) & 1) != 0) == ((stb & ST_CLASS_P) != 0);
  if (nega != negb) return nega;
  return nega ? stamp_less(p, b, a, stb) : stamp_less(p, a, b, sta);
}

// Of two actors landing on the same (lane, pos-param) this tick: the one visited later, which is
// the one bbtree-set keeps (core.scm:173-178).
__device__ int later_processed(const Params& p, int x, int y, int32_t lane, const double* pos_old, int tick) {
  const uint32_t sx = p.st[x], sy = p.st[y];
  const int32_t mx = (sx & ST_MOVED) ? p.mv_old[x] : lane, my = (sy & ST_MOVED) ? p.mv_old[y] : lane;
  const int32_t cx = mx < 0 ? ~mx : mx, cy = my < 0 ? ~my : my;
  if (cx != cy) return cx < cy ? x : y;
  if (mx < 0) return offroad_before(p, x, y, tick) ? y : x;
  return pos_old[x] < pos_old[y] ? x : y;
}

// ------------------------------------------------------------------------------------------------
// k_relink: rebuild the ascending-pos list of every lane an actor entered or left this tick.
__global__ void __launch_bounds__(128) k_relink(Params p, int tick) {
  const int par = tick & 1;
  const int32_t n_touched = p.counters[2 * par + 1];
  const double* __restrict__ pos_old = p.pos[par];
  const double* __restrict__ pos_new = p.pos[par ^ 1];
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_touched; i += gridDim.x * blockDim.x) {
    const int32_t lane = p.touched[i];
    const int32_t inbox = p.inbox_head[lane];
    p.inbox_head[lane] = -1;
    const bool jl = p.kind[lane] == GRIDDY_JUNCTION_LANE;

    // Residents that stay, in list order (they keep their relative order: route.scm:68-74).
    // Ghosts found this tick are unlinked here; one that also moved is finalised by k_settle.
    auto next_resident = [&](int32_t z) {
      while (z >= 0) {
        const uint32_t sz = p.st[z];
        const bool ghost = p.dead_mark[z] == tick + 1;
        if (!(sz & ST_MOVED) && !ghost) break;
        if (ghost && !(sz & ST_MOVED)) {
          p.st[z] = sz & ~ST_ALIVE;
          if (jl) atomicSub(&p.occ[p.lane_junction[lane]], 1);
        }
        z = p.ahead[z];
      }
      return z;
    };
    int32_t x = next_resident(p.lane_tail[lane]);
    // entrants by ascending (pos-param, id): selection over the inbox, which is short
    double e_pos = 0.0;
    int32_t e = -1;
    auto next_entrant = [&](bool first) {
      int32_t best = -1;
      double best_pos = 0.0;
      for (int32_t y = inbox; y >= 0; y = p.inbox_next[y]) {
        if (p.dead_mark[y] == tick + 1) continue;   // a ghost's move is void
        const double py = pos_new[y];
        if (!first && !(py > e_pos || (py == e_pos && y > e))) continue;
        if (best < 0 || py < best_pos || (py == best_pos && y < best)) { best = y; best_pos = py; }
      }
      e = best;
      e_pos = best_pos;
    };
    next_entrant(true);

    int32_t prev = -1, pending = -1;
    auto commit = [&](int32_t z) {
      if (prev < 0) p.lane_tail[lane] = z;
      else if (p.st[prev] & ST_MOVED) p.ahead_new[prev] = z;
      else p.ahead[prev] = z;
      prev = z;
    };
    while (x >= 0 || e >= 0) {
      int32_t cand;
      if (e < 0 || (x >= 0 && pos_new[x] <= e_pos)) {
        cand = x;
        x = next_resident(p.ahead[x]);
      } else {
        cand = e;
        next_entrant(false);
      }
      if (pending < 0) {
        pending = cand;
      } else if (pos_new[cand] == pos_new[pending]) {   // equal key: bbtree-set keeps the later one
        const int32_t win = later_processed(p, cand, pending, lane, pos_old, tick);
        const int32_t lose = win == cand ? pending : cand;
        const uint32_t sl = p.st[lose];
        p.st[lose] = sl & ~ST_ALIVE;
        if (jl && !(sl & ST_MOVED)) atomicSub(&p.occ[p.lane_junction[lane]], 1);
        pending = win;
      } else {
        commit(pending);
        pending = cand;
      }
    }
    if (pending >= 0) commit(pending);
    if (prev < 0) p.lane_tail[lane] = -1;
    else if (p.st[prev] & ST_MOVED) p.ahead_new[prev] = -1;
    else p.ahead[prev] = -1;
  }
}

// k_settle: per mover, publish the new ahead pointer and apply junction occupancy deltas.
__global__ void __launch_bounds__(256) k_settle(Params p, int tick) {
  const int par = tick & 1;
  const int32_t n_movers = p.counters[2 * par];
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_movers; i += gridDim.x * blockDim.x) {
    const int a = p.movers[i];
    const uint32_t st = p.st[a];
    const int32_t old = p.mv_old[a];
    if (old >= 0 && p.kind[old] == GRIDDY_JUNCTION_LANE) atomicSub(&p.occ[p.lane_junction[old]], 1);
    if (p.dead_mark[a] == tick + 1) {   // it was a ghost when it moved: the move is void
      p.st[a] = st & ~(ST_MOVED | ST_ALIVE);
      continue;
    }
    if ((st & ST_ALIVE) && (st & ST_ONROAD)) {
      p.ahead[a] = p.ahead_new[a];
      const int32_t lane = p.container[a];
      if (p.kind[lane] == GRIDDY_JUNCTION_LANE) atomicAdd(&p.occ[p.lane_junction[lane]], 1);
    }
    p.st[a] = st & ~ST_MOVED;
  }
}

// ------------------------------------------------------------------------------------------------
// Read-back: format the world on the device, then one copy per output array.

// Ghost rule for downloads: flag every leader that shares its follower's key.
__global__ void __launch_bounds__(256) k_flag_ghosts(Params p, int par, uint8_t* ghost) {
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= p.n) return;
  const uint32_t st = p.st[a];
  if ((st & (ST_ALIVE | ST_ONROAD)) != (ST_ALIVE | ST_ONROAD)) return;
  const double* pos = p.pos[par];
  const double cur = pos[a];
  for (int32_t x = p.ahead[a]; x >= 0 && pos[x] == cur; x = p.ahead[x]) ghost[x] = 1;
}

struct PackOut {
  uint8_t *alive, *loc_kind, *side, *route_status;
  int32_t *container, *agenda_cur, *sleep_left, *route_head;
  double* pos;
};

__global__ void __launch_bounds__(256) k_pack_state(Params p, int par, const uint8_t* ghost, PackOut o) {
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= p.n) return;
  const uint32_t st = p.st[a];
  const int head = (st >> ST_HEAD_SHIFT) & 3, rs = (st >> ST_RS_SHIFT) & 3;
  const int32_t aux = p.aux[a];
  if (o.alive) o.alive[a] = ((st & ST_ALIVE) && !ghost[a]) ? 1 : 0;
  if (o.loc_kind) o.loc_kind[a] = (st & ST_ONROAD) ? 1 : 0;
  if (o.container) o.container[a] = p.container[a];
  if (o.side) o.side[a] = (!(st & ST_ONROAD) && (st & ST_SIDE)) ? 1 : 0;
  if (o.pos) o.pos[a] = p.pos[par][a];
  if (o.agenda_cur) o.agenda_cur[a] = (int32_t)(p.agenda_cur[a] - p.agenda_off[a]);
  if (o.sleep_left) o.sleep_left[a] = head == HEAD_SLEEP ? aux : 0;
  if (o.route_status) o.route_status[a] = (uint8_t)rs;
  if (o.route_head) o.route_head[a] = rs == GRIDDY_ROUTE_SOME ? aux : -2;
}

// ================================================================================================
// Host side

struct Ctx {
  int device = 0;
  std::string err;
  std::vector<void*> net_allocs, actor_allocs;
  uint8_t* d_ghost = nullptr;   // read-back staging (actor pool)
  uint8_t* d_pack = nullptr;
  Params p{};
  bool have_net = false, have_actors = false;
  int64_t tick = 0, launches = 0;
  double last_ms = 0.0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // host copies needed to format downloads
  std::vector<uint8_t> h_kind;
  std::vector<int32_t> h_lane_junction;
  std::vector<int64_t> h_agenda_off;
  std::vector<uint8_t> h_item_kind;
};

int fail(Ctx* c, int code, const std::string& msg) {
  c->err = msg;
  return code;
}

#define CUDA_TRY(c, expr)                                                              \
  do {                                                                                  \
    cudaError_t e_ = (expr);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return fail(c, e_ == cudaErrorMemoryAllocation ? GRIDDY_ERR_OOM : GRIDDY_ERR_CUDA, \
                  std::string(#expr) + ": " + cudaGetErrorString(e_));                  \
  } while (0)

template <typename T>
int dev_alloc(Ctx* c, std::vector<void*>& pool, T** out, int64_t count) {
  void* ptr = nullptr;
  CUDA_TRY(c, cudaMalloc(&ptr, (size_t)std::max<int64_t>(count, 1) * sizeof(T)));
  pool.push_back(ptr);
  *out = (T*)ptr;
  return GRIDDY_OK;
}

template <typename T>
int dev_upload(Ctx* c, std::vector<void*>& pool, const T** out, const T* host, int64_t count) {
  T* d = nullptr;
  int rc = dev_alloc(c, pool, &d, count);
  if (rc) return rc;
  if (count > 0) CUDA_TRY(c, cudaMemcpy(d, host, (size_t)count * sizeof(T), cudaMemcpyHostToDevice));
  *out = d;
  return GRIDDY_OK;
}

void free_pool(std::vector<void*>& pool) {
  for (void* ptr : pool) cudaFree(ptr);
  pool.clear();
}

#define TRY(expr)            \
  do {                       \
    int rc_ = (expr);        \
    if (rc_) return rc_;     \
  } while (0)

}  // namespace

extern "C" {

int griddy_abi_version(void) { return GRIDDY_ABI_VERSION; }

int griddy_create(void** ctx, int device) {
  if (!ctx) return GRIDDY_ERR_BAD_ARG;
  *ctx = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return GRIDDY_ERR_CUDA;
  Ctx* c = new Ctx();
  if (device < 0) cudaGetDevice(&device);
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreate(&c->stream) != cudaSuccess ||
      cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) {
    delete c;
    return GRIDDY_ERR_CUDA;
  }
  c->p.dt = 1.0 / 15.0;
  c->p.two_size = 10.0;
  *ctx = c;
  return GRIDDY_OK;
}

void griddy_destroy(void* ctx) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return;
  cudaSetDevice(c->device);
  free_pool(c->net_allocs);
  free_pool(c->actor_allocs);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char* griddy_last_error(void* ctx) { return ctx ? ((Ctx*)ctx)->err.c_str() : "null context"; }

int griddy_set_constants(void* ctx, double dt, double two_size) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!(dt > 0.0) || !(two_size >= 0.0)) return fail(c, GRIDDY_ERR_BAD_ARG, "dt must be > 0 and two_size >= 0");
  c->p.dt = dt;
  c->p.two_size = two_size;
  return GRIDDY_OK;
}

int griddy_upload_network(void* ctx, int64_t n_items, const uint8_t* kind, const double* lane_len,
                          const uint8_t* lane_dir, const int32_t* lane_segment,
                          const int32_t* lane_out, const int32_t* lane_junction,
                          const int32_t* seg_outer_forw, const int32_t* seg_outer_back) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (n_items < 0 || n_items > INT32_MAX || !kind || !lane_len || !lane_dir || !lane_segment || !lane_out ||
      !lane_junction || !seg_outer_forw || !seg_outer_back)
    return fail(c, GRIDDY_ERR_BAD_ARG, "upload_network: null array or bad n_items");
  auto in_range = [&](int32_t v, int want_kind) { return v >= 0 && v < n_items && kind[v] == want_kind; };
  for (int64_t r = 0; r < n_items; ++r) {
    const bool lane = kind[r] == GRIDDY_SEGMENT_LANE || kind[r] == GRIDDY_JUNCTION_LANE;
    if (lane && !(lane_len[r] > 0.0)) return fail(c, GRIDDY_ERR_BAD_ARG, "lane " + std::to_string(r) + " has no positive length");
    if (kind[r] == GRIDDY_SEGMENT_LANE && !in_range(lane_segment[r], GRIDDY_SEGMENT))
      return fail(c, GRIDDY_ERR_BAD_ARG, "segment lane " + std::to_string(r) + " has no segment");
    if (kind[r] == GRIDDY_JUNCTION_LANE &&
        (!in_range(lane_out[r], GRIDDY_SEGMENT_LANE) || !in_range(lane_junction[r], GRIDDY_JUNCTION)))
      return fail(c, GRIDDY_ERR_BAD_ARG, "junction lane " + std::to_string(r) + " is not linked");
    if (kind[r] == GRIDDY_SEGMENT)
      for (int32_t o : {seg_outer_forw[r], seg_outer_back[r]})
        if (o != -1 && !in_range(o, GRIDDY_SEGMENT_LANE))
          return fail(c, GRIDDY_ERR_BAD_ARG, "segment " + std::to_string(r) + " has a bad outer lane");
  }
  CUDA_TRY(c, cudaSetDevice(c->device));
  free_pool(c->net_allocs);
  free_pool(c->actor_allocs);
  c->have_net = c->have_actors = false;
  Params& p = c->p;
  p.n_items = n_items;
  p.grid_n = 0;
  p.lane_from_j = p.lane_to_j = p.turn4 = nullptr;
  TRY(dev_upload(c, c->net_allocs, &p.kind, kind, n_items));
  TRY(dev_upload(c, c->net_allocs, &p.lane_len, lane_len, n_items));
  TRY(dev_upload(c, c->net_allocs, &p.lane_dir, lane_dir, n_items));
  TRY(dev_upload(c, c->net_allocs, &p.lane_segment, lane_segment, n_items));
  TRY(dev_upload(c, c->net_allocs, &p.lane_out, lane_out, n_items));
  TRY(dev_upload(c, c->net_allocs, &p.lane_junction, lane_junction, n_items));
  TRY(dev_upload(c, c->net_allocs, &p.outer_forw, seg_outer_forw, n_items));
  TRY(dev_upload(c, c->net_allocs, &p.outer_back, seg_outer_back, n_items));
  TRY(dev_alloc(c, c->net_allocs, &p.lane_tail, n_items));
  TRY(dev_alloc(c, c->net_allocs, &p.occ, n_items));
  TRY(dev_alloc(c, c->net_allocs, &p.lane_mark, n_items));
  TRY(dev_alloc(c, c->net_allocs, &p.inbox_head, n_items));
  c->h_kind.assign(kind, kind + n_items);
  c->h_lane_junction.assign(lane_junction, lane_junction + n_items);
  c->have_net = true;
  return GRIDDY_OK;
}

int griddy_set_routing_grid(void* ctx, int32_t n, const int32_t* lane_from_j, const int32_t* lane_to_j,
                            const int32_t* turn4) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_net) return fail(c, GRIDDY_ERR_BAD_ARG, "set_routing_grid before upload_network");
  if (n < 1 || (int64_t)n * n > c->p.n_items || !lane_from_j || !lane_to_j || !turn4)
    return fail(c, GRIDDY_ERR_BAD_ARG, "set_routing_grid: bad arguments");
  CUDA_TRY(c, cudaSetDevice(c->device));
  Params& p = c->p;
  TRY(dev_upload(c, c->net_allocs, &p.lane_from_j, lane_from_j, p.n_items));
  TRY(dev_upload(c, c->net_allocs, &p.lane_to_j, lane_to_j, p.n_items));
  TRY(dev_upload(c, c->net_allocs, &p.turn4, turn4, 4 * p.n_items));
  p.grid_n = n;
  return GRIDDY_OK;
}

int griddy_upload_actors(void* ctx, int64_t n, const uint8_t* loc_kind, const int32_t* container,
                         const double* pos, const uint8_t* side, const double* speed,
                         const double* speed_dt, const int64_t* agenda_off, const uint8_t* item_kind,
                         const int32_t* item_sleep, const int32_t* item_seg, const uint8_t* item_side,
                         const double* item_pos, const int64_t* route_off, const int32_t* route_lane) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_net) return fail(c, GRIDDY_ERR_BAD_ARG, "upload_actors before upload_network");
  if (n < 0 || n > INT32_MAX - 1 || !loc_kind || !container || !pos || !side || !speed || !speed_dt || !agenda_off)
    return fail(c, GRIDDY_ERR_BAD_ARG, "upload_actors: null array or bad n");
  const int64_t n_agenda = agenda_off[n];
  if (n_agenda < 0 || n_agenda > INT32_MAX - 1) return fail(c, GRIDDY_ERR_BAD_ARG, "too many agenda items");
  if (n_agenda > 0 && (!item_kind || !item_sleep || !item_seg || !item_side || !item_pos))
    return fail(c, GRIDDY_ERR_BAD_ARG, "upload_actors: null agenda arrays");
  if (!route_off && c->p.grid_n == 0)
    return fail(c, GRIDDY_ERR_BAD_ARG, "no routes uploaded and no grid routing table set");
  const Params& q = c->p;
  for (int64_t a = 0; a < n; ++a) {
    const int32_t ct = container[a];
    if (ct < 0 || ct >= q.n_items) return fail(c, GRIDDY_ERR_BAD_ARG, "actor " + std::to_string(a) + ": container out of range");
    const uint8_t k = c->h_kind[ct];
    if (loc_kind[a] ? (k != GRIDDY_SEGMENT_LANE && k != GRIDDY_JUNCTION_LANE) : (k != GRIDDY_SEGMENT))
      return fail(c, GRIDDY_ERR_BAD_ARG, "actor " + std::to_string(a) + ": location kind does not match its container");
    if (!(speed[a] > 0.0)) return fail(c, GRIDDY_ERR_BAD_ARG, "actor " + std::to_string(a) + ": max-speed must be positive");
    if (agenda_off[a + 1] < agenda_off[a]) return fail(c, GRIDDY_ERR_BAD_ARG, "agenda_off not monotone");
  }
  for (int64_t k = 0; k < n_agenda; ++k) {
    if (item_kind[k] == GRIDDY_TRAVEL_TO) {
      const int32_t s = item_seg[k];
      if (s < 0 || s >= q.n_items || c->h_kind[s] != GRIDDY_SEGMENT)
        return fail(c, GRIDDY_ERR_BAD_ARG, "agenda item " + std::to_string(k) + ": destination is not a segment");
    } else if (item_kind[k] != GRIDDY_SLEEP_FOR) {
      return fail(c, GRIDDY_ERR_BAD_ARG, "agenda item " + std::to_string(k) + ": unknown kind");
    }
  }
  if (route_off) {
    if (route_off[n_agenda] > 0 && !route_lane) return fail(c, GRIDDY_ERR_BAD_ARG, "route_lane is null");
    if (route_off[n_agenda] > INT32_MAX - 1) return fail(c, GRIDDY_ERR_BAD_ARG, "too many route steps");
    for (int64_t k = 0; k < route_off[n_agenda]; ++k) {
      const int32_t l = route_lane[k];
      if (l < 0 || l >= q.n_items || (c->h_kind[l] != GRIDDY_SEGMENT_LANE && c->h_kind[l] != GRIDDY_JUNCTION_LANE))
        return fail(c, GRIDDY_ERR_BAD_ARG, "route step " + std::to_string(k) + " is not a lane");
    }
  }

  CUDA_TRY(c, cudaSetDevice(c->device));
  free_pool(c->actor_allocs);
  c->have_actors = false;
  Params& p = c->p;
  p.n = n;
  auto& pool = c->actor_allocs;

  // Initial world: link! every actor in id order (core.scm:169-178).
  std::vector<uint32_t> st(n);
  std::vector<int32_t> aux(n, 0), acur(n), ahead(n, -1), land_lane(n);
  std::vector<int32_t> lane_tail(p.n_items, -1), occ(p.n_items, 0);
  std::vector<int32_t> on_road;
  for (int64_t a = 0; a < n; ++a) {
    uint32_t s = ST_ALIVE;
    acur[a] = (int32_t)agenda_off[a];
    if (agenda_off[a] < agenda_off[a + 1]) {
      if (item_kind[agenda_off[a]] == GRIDDY_SLEEP_FOR) { s |= HEAD_SLEEP << ST_HEAD_SHIFT; aux[a] = item_sleep[agenda_off[a]]; }
      else s |= HEAD_TRAVEL << ST_HEAD_SHIFT;
    }
    if (loc_kind[a]) { s |= ST_ONROAD; on_road.push_back((int32_t)a); }
    else if (side[a]) s |= ST_SIDE;
    st[a] = s;
    land_lane[a] = (int32_t)a;   // landing stamp of the initial placement: tick 0, link! order
  }
  // lanes: ascending pos-param; an equal key replaces the earlier actor (bbtree-set)
  std::sort(on_road.begin(), on_road.end(), [&](int32_t x, int32_t y) {
    if (container[x] != container[y]) return container[x] < container[y];
    if (pos[x] != pos[y]) return pos[x] < pos[y];
    return x < y;
  });
  int32_t prev = -1;
  for (size_t k = 0; k < on_road.size(); ++k) {
    const int32_t a = on_road[k];
    const bool replaced = k + 1 < on_road.size() && container[on_road[k + 1]] == container[a] && pos[on_road[k + 1]] == pos[a];
    if (replaced) { st[a] &= ~ST_ALIVE; continue; }
    const int32_t lane = container[a];
    if (prev >= 0 && container[prev] == lane) ahead[prev] = a; else lane_tail[lane] = a;
    if (c->h_kind[lane] == GRIDDY_JUNCTION_LANE) ++occ[c->h_lane_junction[lane]];
    prev = a;
  }

  const uint32_t* d_st; const int32_t* d_i32; const double* d_f64;
  TRY(dev_upload(c, pool, &d_st, st.data(), n)); p.st = (uint32_t*)d_st;
  TRY(dev_upload(c, pool, &d_i32, container, n)); p.container = (int32_t*)d_i32;
  TRY(dev_upload(c, pool, &d_f64, pos, n)); p.pos[0] = (double*)d_f64;
  TRY(dev_upload(c, pool, &d_f64, pos, n)); p.pos[1] = (double*)d_f64;
  TRY(dev_upload(c, pool, &d_i32, ahead.data(), n)); p.ahead = (int32_t*)d_i32;
  TRY(dev_upload(c, pool, &d_i32, aux.data(), n)); p.aux = (int32_t*)d_i32;
  TRY(dev_upload(c, pool, &d_i32, acur.data(), n)); p.agenda_cur = (int32_t*)d_i32;
  TRY(dev_upload(c, pool, &d_i32, land_lane.data(), n)); p.land_lane = (int32_t*)d_i32;
  TRY(dev_alloc(c, pool, &p.delta, n));
  TRY(dev_alloc(c, pool, &p.buf, n));
  TRY(dev_alloc(c, pool, &p.arrive, n));
  TRY(dev_alloc(c, pool, &p.land_pos, n));
  TRY(dev_alloc(c, pool, &p.route_cur, n));
  TRY(dev_alloc(c, pool, &p.land_tick, n));
  TRY(dev_alloc(c, pool, &p.mv_old, n));
  TRY(dev_alloc(c, pool, &p.ahead_new, n));
  TRY(dev_alloc(c, pool, &p.inbox_next, n));
  TRY(dev_alloc(c, pool, &p.dead_mark, n));
  TRY(dev_alloc(c, pool, &c->d_ghost, n));
  TRY(dev_alloc(c, pool, &c->d_pack, 32 * n));   // 4 u8 + 4 i32 + 1 f64 per actor, 8-aligned runs
  TRY(dev_alloc(c, pool, &p.movers, n));
  TRY(dev_alloc(c, pool, &p.touched, 2 * n));
  TRY(dev_alloc(c, pool, &p.counters, 4));
  TRY(dev_alloc(c, pool, &p.err, 1));
  for (void* z : {(void*)p.delta, (void*)p.buf, (void*)p.arrive, (void*)p.land_pos})
    CUDA_TRY(c, cudaMemset(z, 0, (size_t)std::max<int64_t>(n, 1) * sizeof(double)));
  for (void* z : {(void*)p.route_cur, (void*)p.land_tick, (void*)p.mv_old, (void*)p.ahead_new, (void*)p.inbox_next, (void*)p.dead_mark})
    CUDA_TRY(c, cudaMemset(z, 0, (size_t)std::max<int64_t>(n, 1) * sizeof(int32_t)));
  CUDA_TRY(c, cudaMemset(p.counters, 0, 4 * sizeof(int32_t)));
  CUDA_TRY(c, cudaMemset(p.err, 0, sizeof(int32_t)));
  TRY(dev_upload(c, pool, &p.speed, speed, n));
  TRY(dev_upload(c, pool, &p.speed_dt, speed_dt, n));
  TRY(dev_upload(c, pool, &p.agenda_off, agenda_off, n + 1));
  TRY(dev_upload(c, pool, &p.item_kind, item_kind, n_agenda));
  TRY(dev_upload(c, pool, &p.item_sleep, item_sleep, n_agenda));
  TRY(dev_upload(c, pool, &p.item_seg, item_seg, n_agenda));
  TRY(dev_upload(c, pool, &p.item_side, item_side, n_agenda));
  TRY(dev_upload(c, pool, &p.item_pos, item_pos, n_agenda));
  if (route_off) {
    TRY(dev_upload(c, pool, &p.route_off, route_off, n_agenda + 1));
    TRY(dev_upload(c, pool, &p.route_lane, route_lane, route_off[n_agenda]));
  } else {
    p.route_off = nullptr;
    p.route_lane = nullptr;
  }
  // per-lane state of the network belongs to this world
  CUDA_TRY(c, cudaMemcpy(p.lane_tail, lane_tail.data(), (size_t)p.n_items * sizeof(int32_t), cudaMemcpyHostToDevice));
  CUDA_TRY(c, cudaMemcpy(p.occ, occ.data(), (size_t)p.n_items * sizeof(int32_t), cudaMemcpyHostToDevice));
  CUDA_TRY(c, cudaMemset(p.lane_mark, 0, (size_t)std::max<int64_t>(p.n_items, 1) * sizeof(int32_t)));
  CUDA_TRY(c, cudaMemset(p.inbox_head, 0xFF, (size_t)std::max<int64_t>(p.n_items, 1) * sizeof(int32_t)));
  c->h_agenda_off.assign(agenda_off, agenda_off + n + 1);
  c->h_item_kind.assign(item_kind, item_kind + n_agenda);
  c->tick = 0;
  c->have_actors = true;
  return GRIDDY_OK;
}

namespace {

int check_device_error(Ctx* c) {
  int32_t derr = 0;
  CUDA_TRY(c, cudaMemcpy(&derr, c->p.err, sizeof(derr), cudaMemcpyDeviceToHost));
  if (!derr) return GRIDDY_OK;
  std::string m = "advance$:";
  if (derr & DEV_ERR_BEGIN_ON_ROAD) m += " begin-route$ on an on-road actor (no applicable method);";
  if (derr & DEV_ERR_NO_CLAUSE) m += " no matching clause for (agenda, route);";
  if (derr & DEV_ERR_END_ON_JLANE) m += " end-route$ on a junction lane;";
  if (derr & DEV_ERR_NO_LANE) m += " road-has-no-lanes;";
  if (derr & DEV_ERR_NO_ROUTE) m += " routing table has no next hop;";
  return fail(c, GRIDDY_ERR_BAD_STATE, m);
}

// One tick = three launches.  ev (optional) receives 4 events bracketing them.
void launch_tick(Ctx* c, int tick, cudaEvent_t* ev) {
  const Params& p = c->p;
  const unsigned adv_blocks = (unsigned)((p.n + 255) / 256);
  const unsigned aux_blocks = std::min<unsigned>(adv_blocks, 148u * 8u);
  if (ev) cudaEventRecord(ev[0], c->stream);
  k_advance<<<adv_blocks, 256, 0, c->stream>>>(p, tick);
  if (ev) cudaEventRecord(ev[1], c->stream);
  k_relink<<<aux_blocks * 2, 128, 0, c->stream>>>(p, tick);
  if (ev) cudaEventRecord(ev[2], c->stream);
  k_settle<<<aux_blocks, 256, 0, c->stream>>>(p, tick);
  if (ev) cudaEventRecord(ev[3], c->stream);
  c->launches += 3;
}

}  // namespace

int griddy_step(void* ctx, int64_t n_ticks) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_actors) return fail(c, GRIDDY_ERR_BAD_ARG, "step before upload_actors");
  if (n_ticks < 0 || c->tick + n_ticks > INT32_MAX - 2) return fail(c, GRIDDY_ERR_BAD_ARG, "bad n_ticks");
  CUDA_TRY(c, cudaSetDevice(c->device));
  CUDA_TRY(c, cudaEventRecord(c->ev0, c->stream));
  if (c->p.n > 0)
    for (int64_t t = 0; t < n_ticks; ++t) launch_tick(c, (int)(c->tick + t), nullptr);
  CUDA_TRY(c, cudaEventRecord(c->ev1, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  CUDA_TRY(c, cudaGetLastError());
  float ms = 0.f;
  CUDA_TRY(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->last_ms = ms;
  c->tick += n_ticks;
  return check_device_error(c);
}

int griddy_profile_step(void* ctx, int64_t n_ticks, double* kernel_ms) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_actors || !kernel_ms) return fail(c, GRIDDY_ERR_BAD_ARG, "profile_step: no actors or null output");
  if (n_ticks < 0 || c->tick + n_ticks > INT32_MAX - 2) return fail(c, GRIDDY_ERR_BAD_ARG, "bad n_ticks");
  CUDA_TRY(c, cudaSetDevice(c->device));
  cudaEvent_t ev[4];
  for (auto& e : ev) CUDA_TRY(c, cudaEventCreate(&e));
  double total = 0.0;
  for (int64_t t = 0; t < n_ticks && c->p.n > 0; ++t) {
    launch_tick(c, (int)(c->tick + t), ev);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 3; ++k) {
      float ms = 0.f;
      CUDA_TRY(c, cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
      kernel_ms[k] += ms;
      total += ms;
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  CUDA_TRY(c, cudaGetLastError());
  c->last_ms = total;
  c->tick += n_ticks;
  return check_device_error(c);
}

namespace {

// Ghost rule (see the header comment): an on-road actor whose list follower holds an equal key was
// replaced at the end of the last tick and is not part of the world.  Clears its alive bit in the
// host copy of `st`.
int drop_ghosts(Ctx* c, std::vector<uint32_t>& st, const std::vector<double>& ps) {
  const Params& p = c->p;
  std::vector<int32_t> ahead(p.n);
  if (p.n) CUDA_TRY(c, cudaMemcpy(ahead.data(), p.ahead, p.n * sizeof(int32_t), cudaMemcpyDeviceToHost));
  std::vector<uint8_t> ghost(p.n, 0);
  for (int64_t a = 0; a < p.n; ++a) {
    if ((st[a] & (ST_ALIVE | ST_ONROAD)) != (ST_ALIVE | ST_ONROAD)) continue;
    for (int32_t x = ahead[a]; x >= 0 && ps[x] == ps[a]; x = ahead[x]) ghost[x] = 1;
  }
  for (int64_t a = 0; a < p.n; ++a)
    if (ghost[a]) st[a] &= ~ST_ALIVE;
  return GRIDDY_OK;
}

}  // namespace

int griddy_download_actors(void* ctx, uint8_t* alive, uint8_t* loc_kind, int32_t* container,
                           uint8_t* side, double* pos, int32_t* agenda_cur, int32_t* sleep_left,
                           uint8_t* route_status, int32_t* route_head, int32_t* order,
                           int64_t* n_order) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_actors) return fail(c, GRIDDY_ERR_BAD_ARG, "download before upload_actors");
  CUDA_TRY(c, cudaSetDevice(c->device));
  const Params& p = c->p;
  const int64_t n = p.n;
  const int tick = (int)c->tick;
  if (n == 0) {
    if (n_order) *n_order = 0;
    return GRIDDY_OK;
  }
  // device-side formatting into one staging block: [pos f64][4 x i32][4 x u8], n entries each
  const int64_t n8 = (n + 7) / 8 * 8;
  uint8_t* base = c->d_pack;
  PackOut o{};
  double* d_pos = (double*)base;
  int32_t* d_i32 = (int32_t*)(base + 8 * n);
  uint8_t* d_u8 = base + 8 * n + 16 * n;
  (void)n8;
  o.pos = pos ? d_pos : nullptr;
  o.container = container ? d_i32 : nullptr;
  o.agenda_cur = agenda_cur ? d_i32 + n : nullptr;
  o.sleep_left = sleep_left ? d_i32 + 2 * n : nullptr;
  o.route_head = route_head ? d_i32 + 3 * n : nullptr;
  o.alive = (alive || order || n_order) ? d_u8 : nullptr;
  o.loc_kind = loc_kind ? d_u8 + n : nullptr;
  o.side = side ? d_u8 + 2 * n : nullptr;
  o.route_status = route_status ? d_u8 + 3 * n : nullptr;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  CUDA_TRY(c, cudaMemsetAsync(c->d_ghost, 0, n, c->stream));
  k_flag_ghosts<<<blocks, 256, 0, c->stream>>>(p, tick & 1, c->d_ghost);
  k_pack_state<<<blocks, 256, 0, c->stream>>>(p, tick & 1, c->d_ghost, o);
  c->launches += 2;
  auto get = [&](void* dst, const void* src, size_t bytes) {
    return dst ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream) : cudaSuccess;
  };
  CUDA_TRY(c, get(pos, d_pos, n * sizeof(double)));
  CUDA_TRY(c, get(container, d_i32, n * sizeof(int32_t)));
  CUDA_TRY(c, get(agenda_cur, d_i32 + n, n * sizeof(int32_t)));
  CUDA_TRY(c, get(sleep_left, d_i32 + 2 * n, n * sizeof(int32_t)));
  CUDA_TRY(c, get(route_head, d_i32 + 3 * n, n * sizeof(int32_t)));
  CUDA_TRY(c, get(alive, d_u8, n));
  CUDA_TRY(c, get(loc_kind, d_u8 + n, n));
  CUDA_TRY(c, get(side, d_u8 + 2 * n, n));
  CUDA_TRY(c, get(route_status, d_u8 + 3 * n, n));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  CUDA_TRY(c, cudaGetLastError());
  if (order || n_order) {
    // get-actors order (world.scm:24-30): containers by descending registration index, a lane's
    // actors by descending pos-param, a segment's in list order (landing stamps, see k_relink).
    std::vector<uint32_t> st(n);
    std::vector<uint8_t> live(n);
    std::vector<int32_t> cont(n), land_tick(n), land_lane(n);
    std::vector<double> ps(n), land_pos(n);
    auto pull = [&](void* dst, const void* src, size_t bytes) { return cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost); };
    CUDA_TRY(c, pull(st.data(), p.st, n * sizeof(uint32_t)));
    CUDA_TRY(c, pull(live.data(), d_u8, n));
    CUDA_TRY(c, pull(cont.data(), p.container, n * sizeof(int32_t)));
    CUDA_TRY(c, pull(ps.data(), p.pos[tick & 1], n * sizeof(double)));
    CUDA_TRY(c, pull(land_tick.data(), p.land_tick, n * sizeof(int32_t)));
    CUDA_TRY(c, pull(land_lane.data(), p.land_lane, n * sizeof(int32_t)));
    CUDA_TRY(c, pull(land_pos.data(), p.land_pos, n * sizeof(double)));
    std::vector<int32_t> ord;
    ord.reserve(n);
    for (int64_t a = 0; a < n; ++a)
      if (live[a]) ord.push_back((int32_t)a);
    auto landed_earlier = [&](int32_t a, int32_t b, int32_t t_land) {
      if (t_land == 0) return land_lane[a] < land_lane[b];
      if (land_lane[a] != land_lane[b]) return land_lane[a] > land_lane[b];
      return land_pos[a] > land_pos[b];
    };
    auto stamp_less = [&](int32_t a, int32_t b) {
      if (land_tick[a] != land_tick[b]) return land_tick[a] < land_tick[b];
      return (st[a] & ST_CLASS_P) ? landed_earlier(b, a, land_tick[a]) : landed_earlier(a, b, land_tick[a]);
    };
    auto offroad_before = [&](int32_t a, int32_t b) {
      const bool nega = (((tick - land_tick[a]) & 1) != 0) == ((st[a] & ST_CLASS_P) != 0);
      const bool negb = (((tick - land_tick[b]) & 1) != 0) == ((st[b] & ST_CLASS_P) != 0);
      if (nega != negb) return nega;
      return nega ? stamp_less(b, a) : stamp_less(a, b);
    };
    std::sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) {
      if (cont[a] != cont[b]) return cont[a] > cont[b];
      if (st[a] & ST_ONROAD) return ps[a] > ps[b];
      return offroad_before(a, b);
    });
    if (order) std::copy(ord.begin(), ord.end(), order);
    if (n_order) *n_order = (int64_t)ord.size();
  }
  return GRIDDY_OK;
}

int griddy_download_occupancy(void* ctx, int32_t* lane_count, int32_t* junction_count) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_actors) return fail(c, GRIDDY_ERR_BAD_ARG, "download before upload_actors");
  CUDA_TRY(c, cudaSetDevice(c->device));
  const Params& p = c->p;
  std::vector<uint32_t> st(p.n), st_raw;
  std::vector<int32_t> cont(p.n);
  std::vector<double> ps(p.n);
  if (p.n) {
    CUDA_TRY(c, cudaMemcpy(st.data(), p.st, p.n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    CUDA_TRY(c, cudaMemcpy(cont.data(), p.container, p.n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CUDA_TRY(c, cudaMemcpy(ps.data(), p.pos[c->tick & 1], p.n * sizeof(double), cudaMemcpyDeviceToHost));
  }
  st_raw = st;
  TRY(drop_ghosts(c, st, ps));
  if (junction_count) {   // the device's own incrementally maintained counters, less the ghosts
    CUDA_TRY(c, cudaMemcpy(junction_count, p.occ, (size_t)p.n_items * sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (int64_t a = 0; a < p.n; ++a)
      if ((st_raw[a] & ST_ALIVE) && !(st[a] & ST_ALIVE) && c->h_kind[cont[a]] == GRIDDY_JUNCTION_LANE)
        --junction_count[c->h_lane_junction[cont[a]]];
  }
  if (lane_count) {
    std::fill(lane_count, lane_count + p.n_items, 0);
    for (int64_t a = 0; a < p.n; ++a)
      if (st[a] & ST_ALIVE) ++lane_count[cont[a]];
  }
  return GRIDDY_OK;
}

double griddy_last_step_ms(void* ctx) { return ctx ? ((Ctx*)ctx)->last_ms : 0.0; }
int64_t griddy_tick(void* ctx) { return ctx ? ((Ctx*)ctx)->tick : 0; }
int64_t griddy_kernel_launches(void* ctx) { return ctx ? ((Ctx*)ctx)->launches : 0; }

}  // extern "C"
