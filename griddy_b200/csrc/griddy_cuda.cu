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
// Data layout (HBM, structure of arrays).  Actors live in SLOTS; a slot's number is not the actor's
// id.  Every REORDER_PERIOD ticks the live actors are radix-sorted by (locality key of their lane or
// segment, coarse pos-param) and all per-slot arrays are permuted (radix_sort.cuh, k_make_keys,
// k_permute): the actors of a lane then sit next to each other in ascending pos-param, neighbouring
// lanes sit in neighbouring cache lines, and dead actors are dropped.  Between reorders slots are
// stable; order within a lane is carried exactly by the `ahead` pointers, so the sort only has to be
// good for locality, never for correctness.
//   hot, read by every on-road actor every tick
//     st         u32  alive | on_road | side | route status | agenda head kind | moved | stamp class
//                     | route head is (arrive-at _)
//     ahead      i32  slot of the next actor ahead on the same lane (ascending pos-param), -1 at the front
//     pos[2]     f64  pos-param, ping-pong by tick parity: followers read the old buffer
//     delta,buf  f64  per-tick advance and tailgate buffer on the current lane, cached at lane entry
//                     (route.scm:62-66: fl(speed*dt)*fl(1/len) and 10/len)
//   cold, touched only when an actor sleeps, changes lane, or starts / ends a route
//     aux        i32  sleeping: remaining time; travelling: head route step (lane, or -1 arrive-at)
//     container  i32  lane (on road) or segment (off road);  arrive f64  target of (arrive-at _)
//     agenda_cur, route_cur, id (the actor's id), land_* (landing stamp, see below)
//   by actor id, immutable:  speed, speed_dt, agenda CSR, explicit routes
//   by item:  lane_tail i32 rearmost actor of a lane;  occ i32 actors on a junction's lanes
// Per-lane order changes only where an actor entered or left a lane this tick; nothing is O(lanes)
// per tick (lanes outnumber actors 1.5:1 to 4:1 on the grids).
//
// Kernels per tick
//   k_advance  one thread per slot: agenda dispatch + motion, all handlers fused.  Coalesced SoA
//              reads; the leader's pos-param is a gather that the slot order keeps inside the same
//              or a neighbouring cache line.  Actors that change lane or road-side are appended to
//              the movers list and pushed on their target lane's inbox (warp-aggregated atomics);
//              their lanes are marked touched.
//   k_unlink   one thread per touched lane.  Lanes are doubly linked lists: leavers and ghosts are
//              taken out in O(1) each; actors that are done for the tick are settled.
//   k_link     one thread per touched lane: entrants are inserted by a short walk up from the rear; an
//              entrant that lands on an occupied key is resolved exactly as bbtree-set would.
//
// Processing order without a global sort.  "a is processed before b" (world.scm:24-30) is decided
// from their OLD locations: different containers -> higher registration index first; same lane ->
// higher pos-param first; same segment -> list order, which is encoded per actor by a landing stamp
// (tick, lane, pos-param, before/after class) because a segment's list reverses every tick.
//
// Equal keys among residents.  Residents of a lane keep their relative order (a follower is clamped
// to leader.old - buf, route.scm:68-74), but only weakly: in fp64 two residents an ulp apart can
// round to the same new key, and bbtree-set then drops the one visited first — the leader, since a
// lane is walked in descending pos-param.  Such ties are always list-adjacent.  They are resolved one
// tick late at no cost: an
// actor whose list follower holds an equal key is a GHOST (it is not part of the world any more).
// Every on-road actor reads its leader's key each tick, sees the tie, marks the ghost dead and has
// the lane relinked; whatever the ghost itself did this tick is discarded by k_unlink.
// Nothing else can observe a ghost: queries reach the follower first and see the same key, and the
// junction it may stand on is occupied by the follower as well.  Downloads apply the same rule.
//
// fp64 parity: every operation of route.scm:53-135 is issued with round-to-nearest intrinsics
// (__dadd_rn / __dmul_rn / __ddiv_rn), which the compiler never contracts into FMAs.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "griddy_cuda.h"
#include "radix_sort.cuh"

namespace {

constexpr uint32_t ST_ALIVE = 1u, ST_ONROAD = 2u, ST_SIDE = 4u;
constexpr int ST_RS_SHIFT = 3;    // 2 bits: GRIDDY_ROUTE_*
constexpr int ST_HEAD_SHIFT = 5;  // 2 bits: 0 no head, 1 sleep-for, 2 travel-to
constexpr uint32_t ST_MOVED = 128u;    // changed lane / road side this tick (cleared by k_unlink / k_link)
constexpr uint32_t ST_CLASS_P = 256u;  // landing stamp: came from a lane visited before the segment
constexpr uint32_t ST_ARRIVE = 512u;   // the route head is (arrive-at _): aux == HOP_ARRIVE
constexpr int HEAD_NONE = 0, HEAD_SLEEP = 1, HEAD_TRAVEL = 2;
constexpr int HOP_ARRIVE = -1;

constexpr uint32_t ST_EMIG = 1024u;     // multi-GPU: moved onto a lane another rank owns (this slot dies in k_unlink)
constexpr uint32_t ST_IMMIG = 2048u;    // multi-GPU: arrived from another rank this tick (cleared in k_link)

constexpr int DEV_ERR_BEGIN_ON_ROAD = 1, DEV_ERR_NO_CLAUSE = 2, DEV_ERR_END_ON_JLANE = 4,
              DEV_ERR_NO_LANE = 8, DEV_ERR_NO_ROUTE = 16, DEV_ERR_INTERNAL = 32, DEV_ERR_MIG_OVERFLOW = 64,
              DEV_ERR_MIG_AGENDA = 128;

// Params::counters[parity][...]
constexpr int CNT_EMIG = 0, CNT_TOUCHED = 1, CNT_SLOW = 2, CNT_STRIDE = 4;
// device scalars (Params::scal)
constexpr int SC_ERR = 0, SC_NSLOTS = 1, SC_NALIVE = 2, SC_NITEMS = 3, SC_COUNT = 4;
constexpr int MAX_RANKS = 8;     // one box
constexpr int MIG_ITEMS = 4;     // agenda items a migrating actor can carry

// st and ahead of a slot share an 8-byte word: every kernel that walks a lane needs both.
struct __align__(8) SA {
  uint32_t st;
  int32_t ahead;
};

// Everything about a slot that only matters when its actor does something rare (changes lane, wakes,
// starts or ends a route), in one 64-byte record — one DRAM burst instead of a dozen sectors.
struct __align__(16) Cold {
  // travels with the actor
  int32_t container;                  // lane (on road) or segment (off road)
  int32_t agenda_cur, route_cur;      // agenda head (index into the item arrays), route head (explicit routes)
  int32_t gid;                        // the actor's id (global id on a partitioned world)
  int32_t behind;                     // slot of the next actor behind on the lane (-1 at the rear)
  int32_t hop;                        // head of the route: lane of (turn-onto lane), or -1 for (arrive-at _)
  int32_t dest_lane, dest_from_j, dest_to_j;   // grid routing: the route's last lane and its junctions
  // list surgery of the current tick
  int32_t mv_old;                     // where a mover came from: lane, or ~segment
  int32_t spare[2];
  int32_t inbox_next, outbox_next, ghost_next;
  int32_t dead_mark;                  // tick+1 when the actor was found to be a ghost during `tick`
};
// The rest of an actor's rarely read state, also one burst.
struct __align__(16) ColdF {
  double arrive;            // target of the (arrive-at _) that ends the current route
  double land_pos;          // landing stamp (with land_tick, land_lane), see "processing order"
  double speed, speed_dt;   // max-speed and fl(max-speed * dt)
  int32_t agenda_beg, agenda_end;     // the actor's agenda in the item arrays
  int32_t land_tick, land_lane;
  int32_t pad[4];
};
static_assert(sizeof(Cold) == 64 && sizeof(ColdF) == 64, "cold records are one DRAM burst each");

// The static network, one 32-byte record per registered item: a lane change reads its target lane's
// kind, length, successor and junctions from one sector.
struct __align__(16) Item {
  double len;             // lanes: length (host-computed, spatial.scm:28-46)
  int32_t link;           // segment lane: its segment; junction lane: the lane it leads onto; segment: outer forw-side lane
  int32_t link2;          // junction lane: its junction; segment: outer back-side lane
  int32_t from_j, to_j;   // grid routing: junction a segment lane leaves / reaches
  uint32_t loc_key;       // locality key for the reorder (default: the registration index)
  uint8_t kind, dir;
  uint16_t pad;
};
// Per-lane list heads: rearmost actor, and this tick's entrants, leavers and ghosts.
struct __align__(16) LaneDyn {
  int32_t tail;      // rearmost actor; on a partitioned world -2 - k marks a lane another rank owns
  int32_t inbox, outbox, ghosts;
  int32_t mark;      // tick+1 once the lane is on this tick's touched list
  int32_t junction;  // junction lanes: their junction (whose occupancy they count towards), else -1
  int32_t pad[2];
};
static_assert(sizeof(Item) == 32 && sizeof(LaneDyn) == 32, "one sector per record");

struct Params {
  // network, by registration index
  int64_t n_items;
  Item* item;
  LaneDyn* lane;
  int32_t* occ;       // actors on a junction's lanes (route.scm:98-104); compact, so that it lives in L2
  int32_t grid_n;
  const int32_t* turn4;
  // slots
  int32_t cap;     // allocated slots
  int32_t* scal;   // SC_*: device error bits, slots in use, live actors counted by the reorder
  SA* sa;          // {st, ahead}: one 8-byte load per actor
  double* pos[2];
  double *delta, *buf;
  int32_t* aux;   // sleeping: remaining time; travelling: head route step (lane, or -1 arrive-at)
  int32_t* tj;    // junction of the lane the route head turns onto, -1 if that is not a junction lane
  Cold* cold;
  ColdF* coldf;
  // agenda items (immutable; actors that arrive from another rank append theirs)
  int64_t n_actors;
  uint8_t* item_kind;
  int32_t *item_sleep, *item_seg;
  uint8_t* item_side;
  double* item_pos;
  int32_t icap;   // item capacity
  const int64_t* route_off;   // null: grid routing table
  const int32_t* route_lane;
  // per-tick scratch, by slot
  int32_t *touched, *slow;
  int32_t* counters;   // [2 parities][CNT_STRIDE]: emigrants, touched lanes, deferred actors
  double dt, two_size;
  // multi-GPU (owner == null on a single GPU): spatial partition, see "Multi-GPU" below
  const int32_t* owner;      // rank that owns each item
  int32_t my_rank;
  const double* halo_recv;   // rearmost positive pos-param of the remote lanes my actors can enter;
                             // lane_tail[remote lane] = -2 - index into this array
  int32_t* emig;             // slots that emigrated this tick
  int32_t mig_cap[MAX_RANKS];     // by destination rank: records per tick
  // by destination rank and tick parity: that rank's receive buffer for my emigrants, in ITS memory
  // (peer access over NVLink): [int32 count, pad][MigRec x mig_cap].  k_emig_pack writes there directly.
  uint8_t* mig_peer[MAX_RANKS][2];
  int32_t* mig_mine[MAX_RANKS][2];   // headers of my own receive buffers, to reset after unpacking
};

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

// New agenda head after a pop (actor.scm:28-31): its kind bits, and the sleep counter in *aux.
__device__ __forceinline__ uint32_t load_head(const Params& p, int32_t a, int32_t ac, int32_t* aux) {
  if (ac >= p.coldf[a].agenda_end) { *aux = 0; return HEAD_NONE; }
  if (p.item_kind[ac] == GRIDDY_SLEEP_FOR) { *aux = p.item_sleep[ac]; return HEAD_SLEEP; }
  *aux = 0;
  return HEAD_TRAVEL;
}

// Lanes whose list must be rebuilt this tick are collected per block in shared memory and flushed to
// Params::touched with one global atomic per block (a same-address atomic per lane would serialise).
struct TouchSink {
  int32_t* list;   // shared, SINK_CAP entries
  int32_t* count;  // shared
};
constexpr int SLOW_THREADS = 128;
constexpr int SINK_CAP = 3 * SLOW_THREADS;   // an actor touches at most its old lane, its new lane and a ghost's lane

__device__ __forceinline__ void mark_touched(const Params& p, int32_t lane, int tick, const TouchSink& sink) {
  if (atomicExch(&p.lane[lane].mark, tick + 1) != tick + 1) sink.list[atomicAdd(sink.count, 1)] = lane;
}

__device__ __forceinline__ void enter_lane(const Params& p, int a, int32_t lane, int tick, const TouchSink& sink) {
  p.cold[a].inbox_next = atomicExch(&p.lane[lane].inbox, a);
  mark_touched(p, lane, tick, sink);
}

__device__ __forceinline__ void leave_lane(const Params& p, int a, int32_t lane, int tick, const TouchSink& sink) {
  p.cold[a].outbox_next = atomicExch(&p.lane[lane].outbox, a);
  mark_touched(p, lane, tick, sink);
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
    const double tlead = p.halo_recv[-2 - x];
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
  p.tj[a] = junction_of_hop(p, nhop);
  p.delta[a] = __dmul_rn(f.speed_dt, tinv);
  p.buf[a] = tbuf;
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
      if (atomicExch(&p.cold[ah].dead_mark, tick + 1) != tick + 1)   // the first to notice hands it to k_relink
        p.cold[ah].ghost_next = atomicExch(&p.lane[lane].ghosts, ah);
      ah = p.sa[ah].ahead;
      if (ah >= 0) lead = pos_old[ah];
    } while (ah >= 0 && lead == cur);
    mark_touched(p, lane, tick, sink);
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
    const int32_t dest_lane = outer_lane(p, p.item_seg[item], p.item_side[item]);
    if (init_lane < 0 || dest_lane < 0) { dev_error(p, DEV_ERR_NO_LANE); return cur; }
    // off-road->on-road  location.scm:49-57
    const double init_pos = p.item[init_lane].dir ? __dsub_rn(1.0, cur) : cur;
    const double ipos = p.item_pos[item];
    const double dest_pos = p.item[dest_lane].dir ? __dsub_rn(1.0, ipos) : ipos;
    int32_t rc = p.route_off ? (int32_t)p.route_off[item] : 0;
    const Item init_it = p.item[init_lane], dest_it = p.item[dest_lane];
    Cold k;
    k.dest_lane = dest_lane;
    k.dest_from_j = dest_it.from_j;
    k.dest_to_j = dest_it.to_j;
    const int32_t hop = route_head_on(p, init_lane, init_it, k, item, &rc);
    const double len = init_it.len;
    p.cold[a].route_cur = rc;
    p.cold[a].dest_lane = dest_lane;
    p.cold[a].dest_from_j = k.dest_from_j;
    p.cold[a].dest_to_j = k.dest_to_j;
    p.coldf[a].arrive = dest_pos;
    p.cold[a].hop = hop;
    p.tj[a] = junction_of_hop(p, hop);
    p.delta[a] = __dmul_rn(p.coldf[a].speed_dt, __ddiv_rn(1.0, len));
    p.buf[a] = __ddiv_rn(p.two_size, len);
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
#ifndef ADV_UNROLL
#define ADV_UNROLL 3
#endif
constexpr int ADV_THREADS = 256;

__global__ void __launch_bounds__(ADV_THREADS) k_advance(Params p, int tick) {
  __shared__ int32_t s_list[ADV_THREADS * ADV_UNROLL];
  __shared__ int32_t s_n, s_base;
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
    dlt[k] = on_route ? p.delta[a] : 0.0;
    buf[k] = on_route ? p.buf[a] : 0.0;
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
    if (slow) s_list[atomicAdd(&s_n, 1)] = turn ? ~a : a;   // ~slot: k_slow goes straight to turn_onto
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
__global__ void __launch_bounds__(SLOW_THREADS) k_slow(Params p, int tick) {
  __shared__ int32_t s_touched[SINK_CAP];
  __shared__ int32_t s_nt, s_base;
  const int par = tick & 1;
  int32_t* counters = p.counters + CNT_STRIDE * par;
  const int32_t n_slow = counters[CNT_SLOW];
  const double* __restrict__ pos_old = p.pos[par];
  double* __restrict__ pos_new = p.pos[par ^ 1];
  const TouchSink sink{s_touched, &s_nt};
  for (int32_t chunk = blockIdx.x * SLOW_THREADS; chunk < n_slow; chunk += gridDim.x * SLOW_THREADS) {
    if (threadIdx.x == 0) s_nt = 0;
    __syncthreads();
    const int32_t i = chunk + threadIdx.x;
    if (i < n_slow) {
      const int32_t e = p.slow[i];
      const int a = e < 0 ? ~e : e;
      const SA sa = p.sa[a];
      const uint32_t st = sa.st;
      const double cur = pos_old[a];
      if (e < 0) {
        pos_new[a] = turn_onto(p, a, tick, st, cur, pos_old, sink);
      } else {
      const bool on_road = st & ST_ONROAD;
      const bool on_route = on_road && ((st >> ST_RS_SHIFT) & 3) == GRIDDY_ROUTE_SOME;
      const int32_t ah = on_road ? sa.ahead : -1;
      const double lead = ah >= 0 ? pos_old[ah] : 0.0;
      const double dlt = on_route ? p.delta[a] : 0.0;
      const double buf = on_route ? p.buf[a] : 0.0;
      pos_new[a] = advance_actor(p, a, tick, st, ah, cur, dlt, buf, lead, pos_old, sink);
      }
    }
    __syncthreads();
    const int nt = s_nt;
    if (threadIdx.x == 0 && nt) s_base = atomicAdd(&counters[CNT_TOUCHED], nt);
    __syncthreads();
    for (int k = threadIdx.x; k < nt; k += SLOW_THREADS) p.touched[s_base + k] = s_touched[k];
  }
}

// ------------------------------------------------------------------------------------------------
// Processing order (world.scm:24-30) from old locations.

// earlier(a,b) among actors that landed on the same segment in the same tick
__device__ __forceinline__ bool landed_earlier(const Params& p, int a, int b, int32_t t_land) {
  const int32_t la = p.coldf[a].land_lane, lb = p.coldf[b].land_lane;
  if (t_land == 0) return la < lb;   // initial placement: land_lane holds the link! order
  if (la != lb) return la > lb;      // higher registration index comes first in static-items
  return p.coldf[a].land_pos > p.coldf[b].land_pos;   // a lane yields its actors in descending pos-param
}

__device__ __forceinline__ bool stamp_less(const Params& p, int a, int b, uint32_t sta) {
  const int32_t ta = p.coldf[a].land_tick, tb = p.coldf[b].land_tick;
  if (ta != tb) return ta < tb;
  return (sta & ST_CLASS_P) ? landed_earlier(p, b, a, ta) : landed_earlier(p, a, b, ta);
}

// Is a before b in their segment's list at `tick`?  The list reverses every tick (core.scm:169-171
// conses while get-actors walks head to tail), so a stamp's sign alternates with tick parity.
__device__ bool offroad_before(const Params& p, int a, int b, int tick) {
  const uint32_t sta = p.sa[a].st, stb = p.sa[b].st;
  const bool nega = (((tick - p.coldf[a].land_tick) & 1) != 0) == ((sta & ST_CLASS_P) != 0);
  const bool negb = (((tick - p.coldf[b].land_tick) & 1) != 0) == ((stb & ST_CLASS_P) != 0);
  if (nega != negb) return nega;
  return nega ? stamp_less(p, b, a, stb) : stamp_less(p, a, b, sta);
}

// Of two actors landing on the same (lane, pos-param) this tick: the one visited later, which is
// the one bbtree-set keeps (core.scm:173-178).
__device__ int later_processed(const Params& p, int x, int y, int32_t lane, const double* pos_old, int tick) {
  const uint32_t sx = p.sa[x].st, sy = p.sa[y].st;
  const int32_t mx = (sx & ST_MOVED) ? p.cold[x].mv_old : lane, my = (sy & ST_MOVED) ? p.cold[y].mv_old : lane;
  const int32_t cx = mx < 0 ? ~mx : mx, cy = my < 0 ? ~my : my;
  if (cx != cy) return cx < cy ? x : y;
  if (mx < 0) return offroad_before(p, x, y, tick) ? y : x;
  return pos_old[x] < pos_old[y] ? x : y;
}

// ------------------------------------------------------------------------------------------------
// k_unlink / k_link: one thread per lane that an actor entered or left this tick.  Lanes are doubly
// linked lists in ascending pos-param, so nothing walks a whole lane.
//   k_unlink  takes the lane's leavers (its outbox) and the ghosts found on it this tick out of the
//             list in O(1) each, and settles the actors that are done for this tick: a ghost's move is
//             void, an emigrant lives on another rank from now on, an actor that left the road has no
//             lane to enter.
//   k_link    (after every lane has been unlinked, so a mover's pointers are free to overwrite)
//             inserts the lane's entrants (its inbox, sorted) by walking up from the rear — entrants
//             land near pos 0, so that is a step or two — and writes their final pointers.
// Residents keep their relative order (a follower is clamped behind its leader, route.scm:68-74).
// Equal keys: an entrant that lands on a resident's or another entrant's key is resolved in k_link,
// the later-processed one stays (core.scm:173-178); two residents that round to the same key are left
// linked and resolved next tick through the ghost rule (file header) — on every lane, touched or not.
// Junction occupancy (route.scm:98-104) changes by whole lanes: one atomic per lane and kernel.
constexpr int32_t UNLINKED = -3;

__global__ void __launch_bounds__(128) k_unlink(Params p, int tick) {
  const int par = tick & 1;
  const int32_t n_touched = p.counters[CNT_STRIDE * par + CNT_TOUCHED];
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_touched; i += gridDim.x * blockDim.x) {
    const int32_t lane = p.touched[i];
    const LaneDyn ld = p.lane[lane];
    if (ld.outbox < 0 && ld.ghosts < 0) continue;
    if (ld.outbox >= 0) p.lane[lane].outbox = -1;
    if (ld.ghosts >= 0) p.lane[lane].ghosts = -1;
    int32_t tail = ld.tail, gone = 0;
    // A ghost that also moved is on both lists; the second visit finds it UNLINKED.
    auto unlink = [&](int32_t z, int32_t b) {
      const int32_t f = p.sa[z].ahead;
      if (f == UNLINKED) return false;
      if (b >= 0) p.sa[b].ahead = f; else tail = f;
      if (f >= 0) p.cold[f].behind = b;
      p.sa[z].ahead = UNLINKED;
      return true;
    };
    for (int32_t z = ld.outbox; z >= 0;) {
      const Cold k = p.cold[z];
      unlink(z, k.behind);
      ++gone;   // every mover leaves its old lane, whatever becomes of it
      const uint32_t sz = p.sa[z].st;
      if (k.dead_mark == tick + 1 || (sz & ST_EMIG)) p.sa[z].st = sz & ~(ST_MOVED | ST_ALIVE | ST_EMIG);
      else if (!(sz & ST_ONROAD)) p.sa[z].st = sz & ~ST_MOVED;   // went off the road: nothing to link
      z = k.outbox_next;
    }
    for (int32_t z = ld.ghosts; z >= 0;) {
      const Cold k = p.cold[z];
      const uint32_t sz = p.sa[z].st;
      if (unlink(z, k.behind) && !(sz & ST_MOVED)) {   // (one that also moved was settled above)
        p.sa[z].st = sz & ~ST_ALIVE;
        ++gone;
      }
      z = k.ghost_next;
    }
    if (tail != ld.tail) p.lane[lane].tail = tail;
    if (ld.junction >= 0 && gone) atomicSub(&p.occ[ld.junction], gone);
  }
}

__global__ void __launch_bounds__(128) k_link(Params p, int tick) {
  const int par = tick & 1;
  const int32_t n_touched = p.counters[CNT_STRIDE * par + CNT_TOUCHED];
  const double* __restrict__ pos_old = p.pos[par];
  const double* __restrict__ pos_new = p.pos[par ^ 1];
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_touched; i += gridDim.x * blockDim.x) {
    const int32_t lane = p.touched[i];
    const LaneDyn ld = p.lane[lane];
    const int32_t inbox = ld.inbox;
    if (inbox < 0) continue;
    p.lane[lane].inbox = -1;
    // entrants by ascending (pos-param, slot): selection over the inbox, which is short
    double e_pos = 0.0;
    int32_t e = -1;
    auto next_entrant = [&](bool first) {
      int32_t best = -1;
      double best_pos = 0.0;
      for (int32_t y = inbox; y >= 0; y = p.cold[y].inbox_next) {
        if (!(p.sa[y].st & ST_MOVED)) continue;   // a ghost's move is void: k_unlink settled it
        const double py = pos_new[y];
        if (!first && !(py > e_pos || (py == e_pos && y > e))) continue;
        if (best < 0 || py < best_pos || (py == best_pos && y < best)) { best = y; best_pos = py; }
      }
      e = best;
      e_pos = best_pos;
    };
    auto link = [&](int32_t b, int32_t z, int32_t f) {   // b -> z -> f
      if (b < 0) p.lane[lane].tail = z; else p.sa[b].ahead = z;
      p.sa[z].ahead = f;
      p.cold[z].behind = b;
      if (f >= 0) p.cold[f].behind = z;
    };
    int32_t prev = -1, x = ld.tail, delta = 0;
    for (next_entrant(true); e >= 0; next_entrant(false)) {
      while (x >= 0 && pos_new[x] < e_pos) { prev = x; x = p.sa[x].ahead; }
      if (x >= 0 && pos_new[x] == e_pos) {   // equal key: bbtree-set keeps the one processed later
        const bool e_wins = later_processed(p, e, x, lane, pos_old, tick) == e;
        const int32_t lose = e_wins ? x : e;
        const uint32_t sl = p.sa[lose].st;
        p.sa[lose].st = sl & ~ST_ALIVE;        // (still marked moved if it is an entrant: see below)
        if (!(sl & ST_MOVED) || lose == x) --delta;   // a resident, or an entrant that had been counted in
        if (!e_wins) continue;
        link(prev, e, p.sa[x].ahead);   // e takes x's place
      } else {
        link(prev, e, x);
      }
      ++delta;
      x = e;   // the next entrant (same or higher key) meets e first
    }
    // the entrants stayed marked as movers so that ties among them could look at where they came from
    for (int32_t y = inbox; y >= 0; y = p.cold[y].inbox_next) {
      const uint32_t sy = p.sa[y].st;
      if (sy & ST_MOVED) p.sa[y].st = sy & ~(ST_MOVED | ST_IMMIG);
    }
    if (ld.junction >= 0 && delta) atomicAdd(&p.occ[ld.junction], delta);
  }
}

// ------------------------------------------------------------------------------------------------
// Set-up: the network is assembled on the device from the arrays that cross the ABI (staged through
// a reusable device buffer), so that no host copy of the 32-byte records is ever made.
__global__ void __launch_bounds__(256) k_item_build(Item* __restrict__ item, int64_t n, const uint8_t* __restrict__ kind,
                                                     const double* __restrict__ len, const uint8_t* __restrict__ dir,
                                                     const int32_t* __restrict__ seg, const int32_t* __restrict__ out,
                                                     const int32_t* __restrict__ junction, const int32_t* __restrict__ of,
                                                     const int32_t* __restrict__ ob) {
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
  it.loc_key = (uint32_t)r;   // default locality key: the registration index
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
  lane[r] = LaneDyn{-1, -1, -1, -1, 0, it.kind == GRIDDY_JUNCTION_LANE ? it.link2 : -1, {0, 0}};
}

__global__ void __launch_bounds__(256) k_lane_set_tail(LaneDyn* __restrict__ lane, const int32_t* __restrict__ which,
                                                        const int32_t* __restrict__ tail, int64_t m) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < m) lane[which[i]].tail = tail[i];
}

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
constexpr size_t MIG_HEADER = 8;

// Halo out: for every lane of mine that a neighbour's actors can enter, its smallest key in (0, ...]
// (the walk route.scm:118-121 would do), +inf if there is none; for every such junction, its occupancy.
__global__ void __launch_bounds__(128) k_halo_pack(Params p, int par, const int32_t* __restrict__ items, int n_lanes,
                                                    int n_junctions, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_lanes + n_junctions) return;
  const int32_t item = items[i];
  if (i >= n_lanes) { out[i] = (double)p.occ[item]; return; }
  const double* __restrict__ pos = p.pos[par];
  int32_t x = p.lane[item].tail;
  while (x >= 0 && !(pos[x] > 0.0)) x = p.sa[x].ahead;
  out[i] = x >= 0 ? pos[x] : __longlong_as_double(0x7ff0000000000000LL);
}

// Halo in: occupancy of the neighbours' junctions (the lane values are read in place from halo_recv).
__global__ void __launch_bounds__(128) k_halo_unpack(Params p, const int32_t* __restrict__ junctions, int n,
                                                      const double* __restrict__ in) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p.occ[junctions[i]] = (int32_t)in[i];
}

// After this tick's immigrants are unpacked: empty my receive buffers of this parity.  The sender writes
// them again two ticks from now, after it has seen the token I send next tick.
__global__ void k_mig_reset(Params p, int par) {
  if (threadIdx.x < MAX_RANKS && p.mig_mine[threadIdx.x][par]) *p.mig_mine[threadIdx.x][par] = 0;
}

// Emigrants -> one record each, written straight into the receive buffer of the rank that owns their
// new lane (peer memory over NVLink: remote atomic for the slot, remote stores for the record), so only
// the actors that actually crossed travel, not a bound-sized message.
__global__ void __launch_bounds__(128) k_emig_pack(Params p, int tick) {
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
    r.delta = p.delta[a];
    r.buf = p.buf[a];
    r.arrive = p.coldf[a].arrive;
    r.land_pos = p.coldf[a].land_pos;
    r.speed = p.coldf[a].speed;
    r.speed_dt = p.coldf[a].speed_dt;
    r.gid = p.cold[a].gid;
    r.st = (int32_t)((p.sa[a].st & ~ST_EMIG) | ST_IMMIG);
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
}

// Immigrants: a fresh slot each, then onto their lane's inbox like any local mover.
__global__ void __launch_bounds__(128) k_immig_unpack(Params p, int tick, const uint8_t* __restrict__ buf, int cap) {
  const int par = tick & 1;
  const int32_t n = min(*(const int32_t*)buf, cap);
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
    p.delta[a] = r.delta;
    p.buf[a] = r.buf;
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
    if (atomicExch(&p.lane[r.container].mark, tick + 1) != tick + 1)
      p.touched[atomicAdd(&p.counters[CNT_STRIDE * par + CNT_TOUCHED], 1)] = r.container;
  }
}

// ------------------------------------------------------------------------------------------------
// Reorder: sort the live slots by (locality key of the container, coarse pos-param), permute every
// per-slot array, remap the slot-valued pointers (ahead, lane_tail) and drop the dead.

// Arrays that travel with an actor.  ahead is remapped through the inverse permutation.
constexpr int PERM4 = 2, PERM8 = 4;
enum { P4_AUX, P4_TJ };
enum { P8_SA, P8_POS, P8_DELTA, P8_BUF };

struct Permute {
  const uint32_t* src4[PERM4];
  uint32_t* dst4[PERM4];
  const uint64_t* src8[PERM8];
  uint64_t* dst8[PERM8];
  const Cold* src_cold;
  Cold* dst_cold;
  const ColdF* src_coldf;
  ColdF* dst_coldf;
};

// grid = n_pad / 256 exactly: every lane of every warp reaches the ballot
__global__ void __launch_bounds__(256) k_make_keys(Params p, int par, int key_bits, uint64_t* __restrict__ keys,
                                                    uint32_t* __restrict__ vals, uint8_t* __restrict__ tailflag) {
  const int s = blockIdx.x * 256 + threadIdx.x;
  uint64_t key = ~0ull;
  uint32_t val = 0xffffffffu;
  bool live = false;
  if (s < p.scal[SC_NSLOTS]) {
    const uint32_t st = p.sa[s].st;
    if (st & ST_ALIVE) {
      live = true;
      const int32_t c = p.cold[s].container;
      const double x = p.pos[par][s];
      // 16 bits of pos-param over [-0.5, 1.5): enough to keep a lane's actors in list order
      const double q = fmin(fmax((x + 0.5) * 32768.0, 0.0), 65535.0);
      // off-road actors after the on-road ones: warps are then all-travelling or all-parked
      key = ((uint64_t)((st & ST_ONROAD) ? 0u : 1u) << (key_bits - 1)) | ((uint64_t)p.item[c].loc_key << 16) | (uint64_t)(uint32_t)q;
      val = (uint32_t)s;
      tailflag[s] = ((st & ST_ONROAD) && p.lane[c].tail == s) ? 1 : 0;
    }
  }
  keys[s] = key;
  vals[s] = val;
  const uint32_t m = __ballot_sync(0xffffffffu, live);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(&p.scal[SC_NALIVE], __popc(m));
}

__global__ void __launch_bounds__(256) k_inverse(const uint32_t* __restrict__ vals, int n_pad, int32_t* __restrict__ newslot) {
  const int s = blockIdx.x * 256 + threadIdx.x;
  if (s >= n_pad) return;
  const uint32_t old = vals[s];
  if (old != 0xffffffffu) newslot[old] = s;
}

__global__ void __launch_bounds__(256) k_permute(Params p, Permute q, const uint32_t* __restrict__ vals,
                                                  const int32_t* __restrict__ newslot,
                                                  const uint8_t* __restrict__ tailflag) {
  const int s = blockIdx.x * 256 + threadIdx.x;
  if (s >= p.scal[SC_NALIVE]) return;
  const uint32_t old = vals[s];
  // ahead / behind are meaningful on the road only; an off-road actor's values are stale
  SA sa = ((const SA*)q.src8[P8_SA])[old];
  const bool on_road = sa.st & ST_ONROAD;
#pragma unroll
  for (int k = 0; k < PERM4; ++k) q.dst4[k][s] = q.src4[k][old];
  Cold cold = q.src_cold[old];
  if (on_road && cold.behind >= 0) {
    cold.behind = newslot[cold.behind];
    if (cold.behind < 0) dev_error(p, DEV_ERR_INTERNAL);   // a live actor's follower must be live
  } else {
    cold.behind = -1;
  }
  q.dst_cold[s] = cold;
  q.dst_coldf[s] = q.src_coldf[old];
#pragma unroll
  for (int k = 0; k < PERM8; ++k) {
    if (k == P8_SA) {
      if (on_road && sa.ahead >= 0) {
        sa.ahead = newslot[sa.ahead];
        if (sa.ahead < 0) dev_error(p, DEV_ERR_INTERNAL);   // a live actor's leader must be live
      } else {
        sa.ahead = -1;
      }
      ((SA*)q.dst8[k])[s] = sa;
    } else {
      q.dst8[k][s] = q.src8[k][old];
    }
  }
  if (tailflag[old]) p.lane[cold.container].tail = s;
}

// ------------------------------------------------------------------------------------------------
// Read-back: format the world on the device (by actor id), then one copy per output array.

// Ghost rule for downloads: flag every leader that shares its follower's key.
__global__ void __launch_bounds__(256) k_flag_ghosts(Params p, int par, uint8_t* ghost) {
  const int a = blockIdx.x * 256 + threadIdx.x;
  if (a >= p.scal[SC_NSLOTS]) return;
  const uint32_t st = p.sa[a].st;
  if ((st & (ST_ALIVE | ST_ONROAD)) != (ST_ALIVE | ST_ONROAD)) return;
  const double* pos = p.pos[par];
  const double cur = pos[a];
  for (int32_t x = p.sa[a].ahead; x >= 0 && pos[x] == cur; x = p.sa[x].ahead) ghost[x] = 1;
}

// Living actors of the world: alive and not a ghost.
__global__ void __launch_bounds__(256) k_count_alive(Params p, const uint8_t* ghost, unsigned long long* out) {
  const int a = blockIdx.x * 256 + threadIdx.x;
  const bool live = a < p.scal[SC_NSLOTS] && (p.sa[a].st & ST_ALIVE) && !ghost[a];
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
  const uint32_t st = p.sa[a].st;
  const int32_t id = o.by_slot ? a : p.cold[a].gid;
  const int head = (st >> ST_HEAD_SHIFT) & 3, rs = (st >> ST_RS_SHIFT) & 3;
  const int32_t aux = p.aux[a];
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

// ================================================================================================
// Host side

struct Ctx;

// One neighbouring rank: what I export to it and import from it every tick.
struct Neighbor {
  int rank = -1;
  int n_exp_lanes = 0, n_exp_junc = 0, n_imp_lanes = 0, n_imp_junc = 0;
  int32_t* d_exp_items = nullptr;   // my lanes, then my junctions, that `rank` reads
  int32_t* d_imp_junc = nullptr;    // its junctions whose occupancy my actors read
  double *d_halo_send = nullptr, *d_halo_recv = nullptr;
  uint8_t* d_mig_recv[2] = {nullptr, nullptr};   // by tick parity; the neighbour's k_emig_pack writes here
  uint8_t* peer_mig[2] = {nullptr, nullptr};     // the neighbour's receive buffers for me, mapped into this process
  int64_t *d_token_out = nullptr, *d_token_in = nullptr;   // NCCL transport: "my emigrants of this tick are written"
  int mig_out = 0, mig_in = 0;      // records per tick at most: a lane lets one actor out per tick
  size_t in_bytes() const { return MIG_HEADER + (size_t)mig_in * sizeof(MigRec); }
};

struct Multi {
  bool on = false;
  int rank = 0, world = 1;
  int mig_cap = 16384;
  int64_t extra_slots = 0;
  std::vector<Neighbor> nb;
  std::vector<void*> allocs;
  std::vector<std::pair<int32_t, int32_t>> import_marks;   // (remote lane, -2 - index into halo_recv)
  double* d_halo_recv = nullptr;
  int32_t* d_owner = nullptr;
  void* nccl_comm = nullptr;          // transport 1: NCCL send/recv between processes, on their own stream
  cudaStream_t comm_stream = nullptr; //   so that k_advance hides the halo exchange and k_unlink the migrants'
  cudaEvent_t ev_packed = nullptr, ev_arrived = nullptr;
  std::vector<Ctx*> peers;            // transport 2: every rank is a context of this process
  cudaEvent_t ev_ready = nullptr;     //   "my send buffers for this phase are written"
};

struct Ctx {
  int device = 0;
  std::string err;
  std::vector<void*> net_allocs, actor_allocs;
  Params p{};
  bool have_net = false, have_actors = false;
  int64_t tick = 0, launches = 0;
  double last_ms = 0.0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // slots
  int32_t ns_host = 0;         // upper bound of the slots in use (refreshed after every step)
  int reorder_period = 384;
  int64_t last_reorder_tick = -1;
  int loc_key_bits = 1;
  uint32_t *alt4[PERM4] = {}, *cur4[PERM4] = {};   // double buffers of the permuted arrays
  uint64_t *alt8[PERM8] = {}, *cur8[PERM8] = {};
  Cold *cold_cur = nullptr, *cold_alt = nullptr;
  ColdF *coldf_cur = nullptr, *coldf_alt = nullptr;
  uint64_t* sort_keys = nullptr;
  uint32_t* sort_vals = nullptr;
  rsort::Scratch sort_scratch;
  int32_t* newslot = nullptr;
  uint8_t* tailflag = nullptr;
  // read-back staging
  uint8_t* d_ghost = nullptr;
  uint8_t* d_pack = nullptr;
  // host copies needed to format downloads
  std::vector<uint8_t> h_kind;
  std::vector<int32_t> h_lane_junction, h_lane_out;
  Multi mg;
  std::vector<int32_t> mg_owner_host;
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

int bits_for(uint32_t max_value) {   // bits that hold 0..max_value with all-ones left unused
  int b = 1;
  while (b < 32 && (((uint64_t)1 << b) - 1) <= max_value) ++b;
  return b;
}

// Point Params at the current buffers of the permuted arrays.
void bind_slot_arrays(Ctx* c, int par) {
  Params& p = c->p;
  p.sa = (SA*)c->cur8[P8_SA];
  p.aux = (int32_t*)c->cur4[P4_AUX];
  p.tj = (int32_t*)c->cur4[P4_TJ];
  p.pos[par] = (double*)c->cur8[P8_POS];
  p.delta = (double*)c->cur8[P8_DELTA];
  p.buf = (double*)c->cur8[P8_BUF];
  p.cold = c->cold_cur;
  p.coldf = c->coldf_cur;
}

// Enqueue a reorder of the slots on the context's stream.  Runs between ticks: no actor is marked
// moved, every inbox is empty.  `tick` is the next tick to run (its parity names the valid pos buffer).
int reorder_slots(Ctx* c, int64_t tick) {
  Params& p = c->p;
  if (c->ns_host == 0) return GRIDDY_OK;
  const int par = (int)(tick & 1);
  const int n_tiles = (c->ns_host + rsort::TILE - 1) / rsort::TILE;
  const int n_pad = n_tiles * rsort::TILE;
  cudaStream_t s = c->stream;
  CUDA_TRY(c, cudaMemsetAsync(&p.scal[SC_NALIVE], 0, sizeof(int32_t), s));
  CUDA_TRY(c, cudaMemsetAsync(c->newslot, 0xFF, (size_t)c->ns_host * sizeof(int32_t), s));
  const int key_bits = c->loc_key_bits + 17;   // [off-road][locality key][pos16]
  k_make_keys<<<n_pad / 256, 256, 0, s>>>(p, par, key_bits, c->sort_keys, c->sort_vals, c->tailflag);
  c->launches += 1 + rsort::sort_pairs(&c->sort_keys, &c->sort_vals, n_tiles, key_bits, &c->sort_scratch, s);
  k_inverse<<<n_pad / 256, 256, 0, s>>>(c->sort_vals, n_pad, c->newslot);
  // the valid pos buffer is the one of parity `par`
  c->cur8[P8_POS] = (uint64_t*)p.pos[par];
  Permute q;
  for (int k = 0; k < PERM4; ++k) { q.src4[k] = c->cur4[k]; q.dst4[k] = c->alt4[k]; }
  for (int k = 0; k < PERM8; ++k) { q.src8[k] = c->cur8[k]; q.dst8[k] = c->alt8[k]; }
  q.src_cold = c->cold_cur; q.dst_cold = c->cold_alt;
  q.src_coldf = c->coldf_cur; q.dst_coldf = c->coldf_alt;
  k_permute<<<n_pad / 256, 256, 0, s>>>(p, q, c->sort_vals, c->newslot, c->tailflag);
  c->launches += 2;
  CUDA_TRY(c, cudaMemcpyAsync(&p.scal[SC_NSLOTS], &p.scal[SC_NALIVE], sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  for (int k = 0; k < PERM4; ++k) std::swap(c->cur4[k], c->alt4[k]);
  for (int k = 0; k < PERM8; ++k) std::swap(c->cur8[k], c->alt8[k]);
  std::swap(c->cold_cur, c->cold_alt);
  std::swap(c->coldf_cur, c->coldf_alt);
  bind_slot_arrays(c, par);
  c->last_reorder_tick = tick;
  CUDA_TRY(c, cudaGetLastError());
  return GRIDDY_OK;
}

void mg_reset(Ctx* c) {
  Multi& m = c->mg;
  for (void* ptr : m.allocs) cudaFree(ptr);
  if (m.ev_ready) cudaEventDestroy(m.ev_ready);
  if (m.ev_packed) cudaEventDestroy(m.ev_packed);
  if (m.ev_arrived) cudaEventDestroy(m.ev_arrived);
  if (m.comm_stream) cudaStreamDestroy(m.comm_stream);
  // the NCCL communicator, if any, is left to process exit: destroying it is collective
  m = Multi();
  c->mg_owner_host.clear();
  c->p.owner = nullptr;
  c->p.halo_recv = nullptr;
  c->p.my_rank = 0;
  for (auto& b : c->p.mig_peer) b[0] = b[1] = nullptr;
  for (auto& b : c->p.mig_mine) b[0] = b[1] = nullptr;
  for (auto& n : c->p.mig_cap) n = 0;
}

// Items of `owner_rank` that actors living on `reader_rank` read before they cross: the lanes they
// can enter (a junction lane fed by one of the reader's segment lanes; a segment lane that one of
// the reader's junction lanes leads onto) and the junctions whose occupancy gates the entry.  Both
// sides derive the same ascending lists from the same global arrays, so no list is ever exchanged.
void mg_plan(int64_t n_items, const uint8_t* kind, const int32_t* lane_in, const int32_t* lane_out,
             const int32_t* lane_junction, const int32_t* owner, int owner_rank, int reader_rank,
             std::vector<int32_t>& lanes, std::vector<int32_t>& junctions) {
  std::vector<uint8_t> mark((size_t)n_items, 0);
  for (int64_t r = 0; r < n_items; ++r) {
    if (kind[r] != GRIDDY_JUNCTION_LANE) continue;
    const int32_t in = lane_in[r], out = lane_out[r];
    if (owner[r] == owner_rank && in >= 0 && owner[in] == reader_rank) { mark[r] = 1; mark[lane_junction[r]] = 2; }
    if (owner[r] == reader_rank && out >= 0 && owner[out] == owner_rank) mark[out] = 1;
  }
  lanes.clear();
  junctions.clear();
  for (int64_t r = 0; r < n_items; ++r) {
    if (mark[r] == 1) lanes.push_back((int32_t)r);
    else if (mark[r] == 2) junctions.push_back((int32_t)r);
  }
}

// Lanes of `from` with a successor lane on `to`.  A lane lets at most one actor out per tick (the
// follower of an actor that reaches the end is a tail-gate buffer behind, route.scm:62-74), so this
// bounds the actors that cross from `from` to `to` in a tick — and sizes the exchange exactly.
int64_t mg_count_feeders(int64_t n_items, const uint8_t* kind, const int32_t* lane_in, const int32_t* lane_out,
                         const int32_t* owner, int from, int to) {
  std::vector<uint8_t> mark((size_t)n_items, 0);
  int64_t n = 0;
  for (int64_t r = 0; r < n_items; ++r) {
    if (kind[r] != GRIDDY_JUNCTION_LANE) continue;
    const int32_t in = lane_in[r], out = lane_out[r];
    if (owner[r] == to && in >= 0 && owner[in] == from && !mark[in]) { mark[in] = 1; ++n; }
    if (owner[r] == from && out >= 0 && owner[out] == to && !mark[r]) { mark[r] = 1; ++n; }
  }
  return n;
}

template <typename T>
int mg_alloc(Ctx* c, T** out, size_t count) {
  void* ptr = nullptr;
  CUDA_TRY(c, cudaMalloc(&ptr, std::max<size_t>(count, 1) * sizeof(T)));
  c->mg.allocs.push_back(ptr);
  *out = (T*)ptr;
  return GRIDDY_OK;
}

// ---- NCCL, bound lazily so that the library loads (and single-GPU worlds run) without it ---------
struct NcclId { char internal[128]; };   // layout of ncclUniqueId (nccl.h)
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId /* ncclUniqueId by value */, int) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

NcclApi* nccl_api() {
  static NcclApi api;
  if (api.lib) return &api;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) return nullptr;
  auto sym = [&](const char* n) { return dlsym(api.lib, n); };
  api.GetUniqueId = (int (*)(void*))sym("ncclGetUniqueId");
  api.CommInitRank = (int (*)(void**, int, NcclId, int))sym("ncclCommInitRank");
  api.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))sym("ncclSend");
  api.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))sym("ncclRecv");
  api.GroupStart = (int (*)())sym("ncclGroupStart");
  api.GroupEnd = (int (*)())sym("ncclGroupEnd");
  api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
  if (!api.GetUniqueId || !api.CommInitRank || !api.Send || !api.Recv || !api.GroupStart || !api.GroupEnd) {
    api.lib = nullptr;
    return nullptr;
  }
  return &api;
}
constexpr int NCCL_CHAR = 0;   // ncclInt8 / ncclChar in nccl.h

#define NCCL_TRY(c, expr)                                                                              \
  do {                                                                                                  \
    int r_ = (expr);                                                                                    \
    if (r_ != 0) {                                                                                      \
      NcclApi* a_ = nccl_api();                                                                         \
      return fail(c, GRIDDY_ERR_NCCL, std::string(#expr) + ": " + (a_ && a_->GetErrorString ? a_->GetErrorString(r_) : "nccl error")); \
    }                                                                                                   \
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
      cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess ||
      rsort::prepare() != cudaSuccess) {
    delete c;
    return GRIDDY_ERR_CUDA;
  }
  c->p.dt = 1.0 / 15.0;
  c->p.two_size = 10.0;
  if (const char* e = getenv("GRIDDY_REORDER_PERIOD")) c->reorder_period = atoi(e);
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

int griddy_set_reorder_period(void* ctx, int32_t ticks) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (ticks < 0) return fail(c, GRIDDY_ERR_BAD_ARG, "reorder period must be >= 0");
  c->reorder_period = ticks;
  return GRIDDY_OK;
}

int griddy_upload_network(void* ctx, int64_t n_items, const uint8_t* kind, const double* lane_len,
                          const uint8_t* lane_dir, const int32_t* lane_segment,
                          const int32_t* lane_out, const int32_t* lane_junction,
                          const int32_t* seg_outer_forw, const int32_t* seg_outer_back) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (n_items < 0 || n_items > INT32_MAX - 1 || !kind || !lane_len || !lane_dir || !lane_segment || !lane_out ||
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
  p.turn4 = nullptr;
  TRY(dev_alloc(c, c->net_allocs, &p.item, n_items));
  TRY(dev_alloc(c, c->net_allocs, &p.lane, n_items));
  TRY(dev_alloc(c, c->net_allocs, &p.occ, n_items));
  if (n_items > 0) {   // stage the ABI arrays on the device, assemble the records there
    std::vector<void*> tmp;
    const uint8_t *d_kind, *d_dir;
    const double* d_len;
    const int32_t *d_seg, *d_out, *d_jun, *d_of, *d_ob;
    int rc = dev_upload(c, tmp, &d_kind, kind, n_items);
    if (!rc) rc = dev_upload(c, tmp, &d_len, lane_len, n_items);
    if (!rc) rc = dev_upload(c, tmp, &d_dir, lane_dir, n_items);
    if (!rc) rc = dev_upload(c, tmp, &d_seg, lane_segment, n_items);
    if (!rc) rc = dev_upload(c, tmp, &d_out, lane_out, n_items);
    if (!rc) rc = dev_upload(c, tmp, &d_jun, lane_junction, n_items);
    if (!rc) rc = dev_upload(c, tmp, &d_of, seg_outer_forw, n_items);
    if (!rc) rc = dev_upload(c, tmp, &d_ob, seg_outer_back, n_items);
    if (!rc) {
      k_item_build<<<(unsigned)((n_items + 255) / 256), 256>>>(p.item, n_items, d_kind, d_len, d_dir, d_seg, d_out, d_jun, d_of, d_ob);
      if (cudaDeviceSynchronize() != cudaSuccess) rc = fail(c, GRIDDY_ERR_CUDA, "k_item_build failed");
    }
    free_pool(tmp);
    if (rc) return rc;
  }
  c->loc_key_bits = bits_for((uint32_t)std::max<int64_t>(n_items, 1));
  c->h_kind.assign(kind, kind + n_items);
  c->h_lane_junction.assign(lane_junction, lane_junction + n_items);
  c->h_lane_out.assign(lane_out, lane_out + n_items);
  mg_reset(c);
  c->have_net = true;
  return GRIDDY_OK;
}

int griddy_set_locality_keys(void* ctx, const uint32_t* item_key) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_net || !item_key) return fail(c, GRIDDY_ERR_BAD_ARG, "set_locality_keys: no network or null keys");
  CUDA_TRY(c, cudaSetDevice(c->device));
  uint32_t mx = 0;
  for (int64_t r = 0; r < c->p.n_items; ++r) mx = std::max(mx, item_key[r]);
  if (mx == UINT32_MAX) return fail(c, GRIDDY_ERR_BAD_ARG, "locality keys must be below 2^32-1");
  {
    std::vector<void*> tmp;
    const uint32_t* d_key;
    int rc = dev_upload(c, tmp, &d_key, item_key, c->p.n_items);
    if (!rc && c->p.n_items > 0) {
      k_item_set_key<<<(unsigned)((c->p.n_items + 255) / 256), 256>>>(c->p.item, c->p.n_items, d_key);
      if (cudaDeviceSynchronize() != cudaSuccess) rc = fail(c, GRIDDY_ERR_CUDA, "k_item_set_key failed");
    }
    free_pool(tmp);
    if (rc) return rc;
  }
  c->loc_key_bits = bits_for(mx);
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
  {
    std::vector<void*> tmp;
    const int32_t *d_from, *d_to;
    int rc = dev_upload(c, tmp, &d_from, lane_from_j, p.n_items);
    if (!rc) rc = dev_upload(c, tmp, &d_to, lane_to_j, p.n_items);
    if (!rc && p.n_items > 0) {
      k_item_set_grid<<<(unsigned)((p.n_items + 255) / 256), 256>>>(p.item, p.n_items, d_from, d_to);
      if (cudaDeviceSynchronize() != cudaSuccess) rc = fail(c, GRIDDY_ERR_CUDA, "k_item_set_grid failed");
    }
    free_pool(tmp);
    if (rc) return rc;
  }
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
  if (n < 0 || n > INT32_MAX - 2 * rsort::TILE || !loc_kind || !container || !pos || !side || !speed || !speed_dt || !agenda_off)
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
  p.n_actors = n;
  const int64_t extra = c->mg.on ? c->mg.extra_slots : 0;   // room for actors arriving from other ranks
  if (c->mg.on && route_off) return fail(c, GRIDDY_ERR_BAD_ARG, "multi-GPU worlds need the grid routing table, not uploaded routes");
  if (n + extra > INT32_MAX - 2 * rsort::TILE || n_agenda + extra * MIG_ITEMS > INT32_MAX - 1)
    return fail(c, GRIDDY_ERR_BAD_ARG, "too many slots");
  const int64_t cap = (n + extra + rsort::TILE - 1) / rsort::TILE * rsort::TILE + rsort::TILE;   // tile-padded
  p.cap = (int32_t)cap;
  p.icap = (int32_t)(n_agenda + extra * MIG_ITEMS);
  auto& pool = c->actor_allocs;

  // Initial world: link! every actor in id order (core.scm:169-178).  Slot = id until the first reorder.
  std::vector<SA> sa(n);
  std::vector<int32_t> aux(n, 0);
  std::vector<Cold> cold(n);
  std::vector<ColdF> coldf(n);
  std::vector<int32_t> occ(p.n_items, 0);
  std::vector<int32_t> tail_lane, tail_slot;   // the few lanes that start with actors on them, and remote-lane marks
  std::vector<int32_t> on_road;
  for (int64_t a = 0; a < n; ++a) {
    uint32_t s = ST_ALIVE;
    Cold& k = cold[a];
    memset(&k, 0, sizeof(k));
    k.container = container[a];
    k.agenda_cur = (int32_t)agenda_off[a];
    k.gid = (int32_t)a;         // global id = upload index until griddy_mg_set_global_ids
    k.behind = -1;
    k.hop = HOP_ARRIVE;
    k.dest_lane = k.dest_from_j = k.dest_to_j = -1;
    ColdF& f = coldf[a];
    memset(&f, 0, sizeof(f));
    f.speed = speed[a];
    f.speed_dt = speed_dt[a];
    f.agenda_beg = (int32_t)agenda_off[a];
    f.agenda_end = (int32_t)agenda_off[a + 1];
    f.land_lane = (int32_t)a;   // landing stamp of the initial placement: tick 0, link! order
    if (agenda_off[a] < agenda_off[a + 1]) {
      if (item_kind[agenda_off[a]] == GRIDDY_SLEEP_FOR) { s |= HEAD_SLEEP << ST_HEAD_SHIFT; aux[a] = item_sleep[agenda_off[a]]; }
      else s |= HEAD_TRAVEL << ST_HEAD_SHIFT;
    }
    if (loc_kind[a]) { s |= ST_ONROAD; on_road.push_back((int32_t)a); }
    else if (side[a]) s |= ST_SIDE;
    sa[a] = SA{s, -1};
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
    if (replaced) { sa[a].st &= ~ST_ALIVE; continue; }
    const int32_t lane = container[a];
    if (prev >= 0 && container[prev] == lane) { sa[prev].ahead = a; cold[a].behind = prev; } else { tail_lane.push_back(lane); tail_slot.push_back(a); }
    if (c->h_kind[lane] == GRIDDY_JUNCTION_LANE) ++occ[c->h_lane_junction[lane]];
    prev = a;
  }

  // permuted arrays: two buffers each
  for (int k = 0; k < PERM4; ++k) {
    TRY(dev_alloc(c, pool, &c->cur4[k], cap));
    TRY(dev_alloc(c, pool, &c->alt4[k], cap));
    CUDA_TRY(c, cudaMemset(c->cur4[k], 0, (size_t)cap * sizeof(uint32_t)));
  }
  for (int k = 0; k < PERM8; ++k) {
    TRY(dev_alloc(c, pool, &c->cur8[k], cap));
    TRY(dev_alloc(c, pool, &c->alt8[k], cap));
    CUDA_TRY(c, cudaMemset(c->cur8[k], 0, (size_t)cap * sizeof(uint64_t)));
  }
  TRY(dev_alloc(c, pool, &c->cold_cur, cap));
  TRY(dev_alloc(c, pool, &c->cold_alt, cap));
  TRY(dev_alloc(c, pool, &c->coldf_cur, cap));
  TRY(dev_alloc(c, pool, &c->coldf_alt, cap));
  CUDA_TRY(c, cudaMemset(c->cold_cur, 0, (size_t)cap * sizeof(Cold)));
  CUDA_TRY(c, cudaMemset(c->coldf_cur, 0, (size_t)cap * sizeof(ColdF)));
  if (n) {
    CUDA_TRY(c, cudaMemcpy(c->cur8[P8_SA], sa.data(), (size_t)n * sizeof(SA), cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->cur8[P8_POS], pos, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->cur4[P4_AUX], aux.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->cold_cur, cold.data(), (size_t)n * sizeof(Cold), cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->coldf_cur, coldf.data(), (size_t)n * sizeof(ColdF), cudaMemcpyHostToDevice));
  }
  CUDA_TRY(c, cudaMemset(c->cur4[P4_TJ], 0xFF, (size_t)cap * sizeof(int32_t)));
  double* pos1 = nullptr;
  TRY(dev_alloc(c, pool, &pos1, cap));
  if (n) CUDA_TRY(c, cudaMemcpy(pos1, pos, (size_t)n * 8, cudaMemcpyHostToDevice));
  p.pos[1] = pos1;
  bind_slot_arrays(c, 0);

  TRY(dev_alloc(c, pool, &p.touched, 2 * cap));
  TRY(dev_alloc(c, pool, &p.slow, cap));
  TRY(dev_alloc(c, pool, &p.emig, cap));
  TRY(dev_alloc(c, pool, &p.counters, 2 * CNT_STRIDE));
  TRY(dev_alloc(c, pool, &p.scal, SC_COUNT));
  CUDA_TRY(c, cudaMemset(p.counters, 0, 2 * CNT_STRIDE * sizeof(int32_t)));
  {
    int32_t scal[SC_COUNT] = {0, (int32_t)n, 0, (int32_t)n_agenda};
    CUDA_TRY(c, cudaMemcpy(p.scal, scal, sizeof(scal), cudaMemcpyHostToDevice));
  }
  // reorder scratch
  const int n_tiles = (int)(cap / rsort::TILE);
  TRY(dev_alloc(c, pool, &c->sort_keys, cap));
  TRY(dev_alloc(c, pool, &c->sort_vals, cap));
  TRY(dev_alloc(c, pool, &c->sort_scratch.keys_alt, cap));
  TRY(dev_alloc(c, pool, &c->sort_scratch.vals_alt, cap));
  TRY(dev_alloc(c, pool, &c->sort_scratch.hist, (int64_t)rsort::hist_words(n_tiles)));
  TRY(dev_alloc(c, pool, &c->sort_scratch.sums, (int64_t)rsort::scan_blocks(rsort::hist_words(n_tiles))));
  TRY(dev_alloc(c, pool, &c->newslot, cap));
  TRY(dev_alloc(c, pool, &c->tailflag, cap));
  CUDA_TRY(c, cudaMemset(c->tailflag, 0, (size_t)cap));
  // read-back staging, by actor id: [pos f64][4 x i32][4 x u8]
  TRY(dev_alloc(c, pool, &c->d_ghost, cap));
  TRY(dev_alloc(c, pool, &c->d_pack, 32 * cap));   // by handle, or by slot for griddy_download_local
  CUDA_TRY(c, cudaMemset(c->d_pack, 0, (size_t)(32 * cap)));

  {   // agenda items; room at the end for the items of actors that arrive from other ranks
    auto up = [&](auto** dst, const auto* src, int64_t count, int64_t capacity) -> int {
      TRY(dev_alloc(c, pool, dst, capacity));
      if (count > 0) CUDA_TRY(c, cudaMemcpy(*dst, src, (size_t)count * sizeof(**dst), cudaMemcpyHostToDevice));
      return GRIDDY_OK;
    };
    TRY(up(&p.item_kind, item_kind, n_agenda, p.icap));
    TRY(up(&p.item_sleep, item_sleep, n_agenda, p.icap));
    TRY(up(&p.item_seg, item_seg, n_agenda, p.icap));
    TRY(up(&p.item_side, item_side, n_agenda, p.icap));
    TRY(up(&p.item_pos, item_pos, n_agenda, p.icap));
  }
  if (route_off) {
    TRY(dev_upload(c, pool, &p.route_off, route_off, n_agenda + 1));
    TRY(dev_upload(c, pool, &p.route_lane, route_lane, route_off[n_agenda]));
  } else {
    p.route_off = nullptr;
    p.route_lane = nullptr;
  }
  // per-lane state of the network belongs to this world
  for (const auto& m : c->mg.import_marks) { tail_lane.push_back(m.first); tail_slot.push_back(m.second); }   // lanes other ranks own
  if (c->mg.on)
    for (int64_t a = 0; a < n; ++a)
      if (c->mg_owner_host[container[a]] != c->mg.rank)
        return fail(c, GRIDDY_ERR_BAD_ARG, "actor " + std::to_string(a) + " is not on this rank's part of the world");
  if (p.n_items > 0) k_lane_init<<<(unsigned)((p.n_items + 255) / 256), 256>>>(p.lane, p.item, p.n_items);
  if (!tail_lane.empty()) {
    std::vector<void*> tmp;
    const int32_t *d_which, *d_tail;
    int rc = dev_upload(c, tmp, &d_which, tail_lane.data(), (int64_t)tail_lane.size());
    if (!rc) rc = dev_upload(c, tmp, &d_tail, tail_slot.data(), (int64_t)tail_slot.size());
    if (!rc) k_lane_set_tail<<<(unsigned)((tail_lane.size() + 255) / 256), 256>>>(p.lane, d_which, d_tail, (int64_t)tail_lane.size());
    if (cudaDeviceSynchronize() != cudaSuccess && !rc) rc = fail(c, GRIDDY_ERR_CUDA, "k_lane_set_tail failed");
    free_pool(tmp);
    if (rc) return rc;
  }
  CUDA_TRY(c, cudaMemcpy(p.occ, occ.data(), (size_t)p.n_items * sizeof(int32_t), cudaMemcpyHostToDevice));
  c->tick = 0;
  c->ns_host = (int32_t)n;
  c->last_reorder_tick = -1;
  c->have_actors = true;
  if (c->reorder_period > 0 && n > 0) {   // spatial order from the first tick on
    TRY(reorder_slots(c, 0));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  }
  return GRIDDY_OK;
}

namespace {

// End of a step: device error bits and the slot count, in one copy.
int finish_step(Ctx* c) {
  int32_t scal[2] = {0, 0};
  CUDA_TRY(c, cudaMemcpy(scal, c->p.scal, sizeof(scal), cudaMemcpyDeviceToHost));
  c->ns_host = scal[SC_NSLOTS];
  const int32_t derr = scal[SC_ERR];
  if (!derr) return GRIDDY_OK;
  std::string m = "advance$:";
  if (derr & DEV_ERR_BEGIN_ON_ROAD) m += " begin-route$ on an on-road actor (no applicable method);";
  if (derr & DEV_ERR_NO_CLAUSE) m += " no matching clause for (agenda, route);";
  if (derr & DEV_ERR_END_ON_JLANE) m += " end-route$ on a junction lane;";
  if (derr & DEV_ERR_NO_LANE) m += " road-has-no-lanes;";
  if (derr & DEV_ERR_NO_ROUTE) m += " routing table has no next hop;";
  if (derr & DEV_ERR_INTERNAL) m += " internal: a live actor pointed at a dropped slot;";
  if (derr & DEV_ERR_MIG_OVERFLOW) m += " multi-GPU: more boundary crossings in one tick than mig_cap, or no slots left for arrivals;";
  if (derr & DEV_ERR_MIG_AGENDA) m += " multi-GPU: a migrating actor's agenda has more than 4 items;";
  return fail(c, GRIDDY_ERR_BAD_STATE, m);
}

// One tick = four launches (plus a reorder every reorder_period ticks; plus, on a partitioned world,
// the halo and migrant kernels and two exchanges).  ev (optional) receives 6 events bracketing
// reorder (+ halo pack), k_advance, k_slow (+ emigrant pack), k_unlink (+ immigrant unpack), k_link.
int tick_begin(Ctx* c, int64_t tick, cudaEvent_t* ev) {
  if (ev) cudaEventRecord(ev[0], c->stream);
  if (c->reorder_period > 0 && tick % c->reorder_period == 0 && tick != c->last_reorder_tick)
    TRY(reorder_slots(c, tick));
  if (c->mg.on) {
    for (Neighbor& nb : c->mg.nb) {
      const int n = nb.n_exp_lanes + nb.n_exp_junc;
      if (n) k_halo_pack<<<(n + 127) / 128, 128, 0, c->stream>>>(c->p, (int)(tick & 1), nb.d_exp_items, nb.n_exp_lanes, nb.n_exp_junc, nb.d_halo_send);
    }
    c->launches += (int64_t)c->mg.nb.size();
  }
  if (ev) cudaEventRecord(ev[1], c->stream);
  return GRIDDY_OK;
}

// k_advance does not read the halo (actors gated by a remote junction are deferred to k_slow), so it
// can run while the halo is in flight.
int tick_advance(Ctx* c, int64_t tick, cudaEvent_t* ev) {
  const unsigned per_block = ADV_THREADS * ADV_UNROLL;
  const unsigned adv_blocks = (unsigned)((c->ns_host + per_block - 1) / per_block);
  if (adv_blocks) k_advance<<<adv_blocks, ADV_THREADS, 0, c->stream>>>(c->p, (int)tick);
  if (ev) cudaEventRecord(ev[2], c->stream);
  ++c->launches;
  return GRIDDY_OK;
}

int tick_slow(Ctx* c, int64_t tick, cudaEvent_t* ev) {   // needs the halo
  const Params& p = c->p;
  const unsigned aux_blocks = std::min<unsigned>(std::max((unsigned)((c->ns_host + 255) / 256), 1u), 148u * 8u);
  if (c->mg.on)
    for (Neighbor& nb : c->mg.nb)
      if (nb.n_imp_junc) {
        k_halo_unpack<<<(nb.n_imp_junc + 127) / 128, 128, 0, c->stream>>>(p, nb.d_imp_junc, nb.n_imp_junc, nb.d_halo_recv + nb.n_imp_lanes);
        ++c->launches;
      }
  k_slow<<<aux_blocks * 2, SLOW_THREADS, 0, c->stream>>>(p, (int)tick);
  if (c->mg.on) {
    k_emig_pack<<<64, 128, 0, c->stream>>>(p, (int)tick);
    ++c->launches;
  }
  if (ev) cudaEventRecord(ev[3], c->stream);
  ++c->launches;
  return GRIDDY_OK;
}

// one thread per touched lane where possible: grid-striding doubles the longest chain
unsigned relink_blocks(const Ctx* c) {
  return std::min<unsigned>(std::max((unsigned)((c->ns_host / 8 + 127) / 128), 1u), 148u * 64u);
}

int tick_unlink(Ctx* c, int64_t tick, cudaEvent_t* ev) {   // local lists only: does not need the migrants
  k_unlink<<<relink_blocks(c), 128, 0, c->stream>>>(c->p, (int)tick);
  if (ev) cudaEventRecord(ev[4], c->stream);
  ++c->launches;
  return GRIDDY_OK;
}

int tick_link(Ctx* c, int64_t tick, cudaEvent_t* ev) {   // needs the migrants
  const Params& p = c->p;
  if (c->mg.on)
    for (Neighbor& nb : c->mg.nb) {
      k_immig_unpack<<<std::max(1, nb.mig_in / 512), 128, 0, c->stream>>>(p, (int)tick, nb.d_mig_recv[tick & 1], nb.mig_in);
      ++c->launches;
      // upper bound of the slots in use: the neighbour may have filled its buffer
      c->ns_host = (int32_t)std::min<int64_t>(p.cap, (int64_t)c->ns_host + nb.mig_in);
    }
  if (c->mg.on) {
    k_mig_reset<<<1, 32, 0, c->stream>>>(p, (int)(tick & 1));
    ++c->launches;
  }
  k_link<<<relink_blocks(c), 128, 0, c->stream>>>(p, (int)tick);
  if (ev) cudaEventRecord(ev[5], c->stream);
  ++c->launches;
  return GRIDDY_OK;
}

// NCCL transport: one grouped send/recv with every neighbour, on the context's stream.
// What was packed on the compute stream so far is sent; whoever needs what arrives waits for
// exchange_arrived().  In between the compute stream is free to run kernels that need neither.
int exchange_nccl(Ctx* c, bool halo) {
  NcclApi* api = nccl_api();
  if (!api || !c->mg.nccl_comm) return fail(c, GRIDDY_ERR_NCCL, "multi-GPU context is not connected (griddy_mg_connect_nccl)");
  cudaStream_t compute = c->stream;
  CUDA_TRY(c, cudaEventRecord(c->mg.ev_packed, compute));
  CUDA_TRY(c, cudaStreamWaitEvent(c->mg.comm_stream, c->mg.ev_packed, 0));
  struct OnComm {   // the sends and receives below go to the communication stream
    Ctx* c; cudaStream_t saved;
    OnComm(Ctx* c_) : c(c_), saved(c_->stream) { c->stream = c->mg.comm_stream; }
    ~OnComm() { c->stream = saved; }
  } on_comm(c);
  NCCL_TRY(c, api->GroupStart());
  for (Neighbor& nb : c->mg.nb) {
    if (halo) {
      const size_t out = (size_t)(nb.n_exp_lanes + nb.n_exp_junc) * sizeof(double);
      const size_t in = (size_t)(nb.n_imp_lanes + nb.n_imp_junc) * sizeof(double);
      if (out) NCCL_TRY(c, api->Send(nb.d_halo_send, out, NCCL_CHAR, nb.rank, c->mg.nccl_comm, c->stream));
      if (in) NCCL_TRY(c, api->Recv(nb.d_halo_recv, in, NCCL_CHAR, nb.rank, c->mg.nccl_comm, c->stream));
    } else {
      // the records themselves were written into the neighbour's memory by k_emig_pack; this token
      // orders them: it leaves after my kernel has completed, and what arrives tells me the same
      if (!nb.peer_mig[0]) return fail(c, GRIDDY_ERR_BAD_STATE, "migrant buffers of rank " + std::to_string(nb.rank) + " are not mapped (griddy_mg_ipc_import)");
      NCCL_TRY(c, api->Send(nb.d_token_out, sizeof(int64_t), NCCL_CHAR, nb.rank, c->mg.nccl_comm, c->stream));
      NCCL_TRY(c, api->Recv(nb.d_token_in, sizeof(int64_t), NCCL_CHAR, nb.rank, c->mg.nccl_comm, c->stream));
    }
  }
  NCCL_TRY(c, api->GroupEnd());
  CUDA_TRY(c, cudaEventRecord(c->mg.ev_arrived, c->mg.comm_stream));
  return GRIDDY_OK;
}

int exchange_arrived(Ctx* c) {
  CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->mg.ev_arrived, 0));
  return GRIDDY_OK;
}

int launch_tick(Ctx* c, int64_t tick, cudaEvent_t* ev) {
  TRY(tick_begin(c, tick, ev));
  if (c->mg.on) TRY(exchange_nccl(c, true));      // halo in flight ...
  TRY(tick_advance(c, tick, ev));                  // ... behind k_advance
  if (c->mg.on) TRY(exchange_arrived(c));
  TRY(tick_slow(c, tick, ev));
  if (c->mg.on) TRY(exchange_nccl(c, false));     // migrants in flight ...
  TRY(tick_unlink(c, tick, ev));                   // ... behind k_unlink
  if (c->mg.on) TRY(exchange_arrived(c));
  return tick_link(c, tick, ev);
}

// Local transport: every rank is a context of this process; `src` has recorded ev_ready after writing
// its send buffers, each neighbour copies what is addressed to it on its own stream.
int exchange_local(Ctx* src, bool halo) {
  for (Neighbor& nb : src->mg.nb) {
    Ctx* dst = src->mg.peers[nb.rank];
    Neighbor* back = nullptr;
    for (Neighbor& b : dst->mg.nb)
      if (b.rank == src->mg.rank) back = &b;
    if (!back) return fail(src, GRIDDY_ERR_BAD_STATE, "neighbour lists are not symmetric");
    if (nb.n_exp_lanes + nb.n_exp_junc != back->n_imp_lanes + back->n_imp_junc || nb.mig_out != back->mig_in)
      return fail(src, GRIDDY_ERR_BAD_STATE, "exchange plans of two neighbours disagree");
    CUDA_TRY(dst, cudaSetDevice(dst->device));
    CUDA_TRY(dst, cudaStreamWaitEvent(dst->stream, src->mg.ev_ready, 0));   // migrants: src's kernel wrote them in place
    const size_t bytes = (size_t)(nb.n_exp_lanes + nb.n_exp_junc) * sizeof(double);
    if (halo && bytes) CUDA_TRY(dst, cudaMemcpyPeerAsync(back->d_halo_recv, dst->device, nb.d_halo_send, src->device, bytes, dst->stream));
  }
  return GRIDDY_OK;
}

}  // namespace

int griddy_step(void* ctx, int64_t n_ticks) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_actors) return fail(c, GRIDDY_ERR_BAD_ARG, "step before upload_actors");
  if (n_ticks < 0 || c->tick + n_ticks > INT32_MAX - 2) return fail(c, GRIDDY_ERR_BAD_ARG, "bad n_ticks");
  if (c->mg.on && !c->mg.nccl_comm)
    return fail(c, GRIDDY_ERR_BAD_ARG, "a partitioned world steps with griddy_mg_step_local, or after griddy_mg_connect_nccl");
  CUDA_TRY(c, cudaSetDevice(c->device));
  CUDA_TRY(c, cudaEventRecord(c->ev0, c->stream));
  if (c->ns_host > 0 || c->mg.on)
    for (int64_t t = 0; t < n_ticks; ++t) TRY(launch_tick(c, c->tick + t, nullptr));
  CUDA_TRY(c, cudaEventRecord(c->ev1, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  CUDA_TRY(c, cudaGetLastError());
  float ms = 0.f;
  CUDA_TRY(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->last_ms = ms;
  c->tick += n_ticks;
  return finish_step(c);
}

int griddy_mg_step_local(void** ctxs, int32_t n, int64_t n_ticks) {
  if (!ctxs || n < 1 || n > MAX_RANKS || n_ticks < 0) return GRIDDY_ERR_BAD_ARG;
  std::vector<Ctx*> cs((Ctx**)ctxs, (Ctx**)ctxs + n);
  for (Ctx* c : cs)
    if (!c || !c->have_actors || !c->mg.on || (int)c->mg.peers.size() != n)
      return c ? fail(c, GRIDDY_ERR_BAD_ARG, "mg_step_local: context is not part of a connected local group") : GRIDDY_ERR_BAD_ARG;
  for (Ctx* c : cs) {
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaEventRecord(c->ev0, c->stream));
  }
  auto each = [&](int (*phase)(Ctx*, int64_t, cudaEvent_t*), int64_t t, bool ready) -> int {
    for (Ctx* c : cs) {
      CUDA_TRY(c, cudaSetDevice(c->device));
      TRY(phase(c, c->tick + t, nullptr));
      if (ready) CUDA_TRY(c, cudaEventRecord(c->mg.ev_ready, c->stream));
    }
    return GRIDDY_OK;
  };
  for (int64_t t = 0; t < n_ticks; ++t) {
    TRY(each(tick_begin, t, true));
    for (Ctx* c : cs) TRY(exchange_local(c, true));
    TRY(each(tick_advance, t, false));
    TRY(each(tick_slow, t, true));
    for (Ctx* c : cs) TRY(exchange_local(c, false));
    TRY(each(tick_unlink, t, false));
    TRY(each(tick_link, t, false));
  }
  int rc = GRIDDY_OK;
  for (Ctx* c : cs) {
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaGetLastError());
    float ms = 0.f;
    CUDA_TRY(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->last_ms = ms;
    c->tick += n_ticks;
    const int r = finish_step(c);
    if (r && !rc) rc = r;
  }
  return rc;
}

int griddy_profile_step(void* ctx, int64_t n_ticks, double* kernel_ms) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_actors || !kernel_ms) return fail(c, GRIDDY_ERR_BAD_ARG, "profile_step: no actors or null output");
  if (n_ticks < 0 || c->tick + n_ticks > INT32_MAX - 2) return fail(c, GRIDDY_ERR_BAD_ARG, "bad n_ticks");
  CUDA_TRY(c, cudaSetDevice(c->device));
  cudaEvent_t ev[6];
  for (auto& e : ev) CUDA_TRY(c, cudaEventCreate(&e));
  double total = 0.0;
  for (int64_t t = 0; t < n_ticks && c->ns_host > 0; ++t) {
    TRY(launch_tick(c, c->tick + t, ev));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    // kernel_ms = {k_advance, k_unlink, k_link, reorder, k_slow}
    static const int slot_of[5] = {3, 0, 4, 1, 2};
    for (int k = 0; k < 5; ++k) {
      float ms = 0.f;
      CUDA_TRY(c, cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
      kernel_ms[slot_of[k]] += ms;
      total += ms;
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  CUDA_TRY(c, cudaGetLastError());
  c->last_ms = total;
  c->tick += n_ticks;
  return finish_step(c);
}

namespace {

// Read-back shared by griddy_download_actors (outputs indexed by actor handle = upload id, n entries)
// and griddy_download_local (outputs indexed by slot, plus each slot's global id).
int download_state(Ctx* c, bool by_slot, int32_t* gid, uint8_t* alive, uint8_t* loc_kind, int32_t* container,
                   uint8_t* side, double* pos, int32_t* agenda_cur, int32_t* sleep_left, uint8_t* route_status,
                   int32_t* route_head, int32_t* order, int64_t* n_order) {
  CUDA_TRY(c, cudaSetDevice(c->device));
  const Params& p = c->p;
  const int32_t ns = c->ns_host;
  const int64_t n = by_slot ? ns : p.n_actors;   // entries of every output array
  const int tick = (int)c->tick;
  if (n == 0) {
    if (n_order) *n_order = 0;
    return GRIDDY_OK;
  }
  // device-side formatting into one staging block: [pos f64][5 x i32][4 x u8], n entries each
  uint8_t* base = c->d_pack;
  PackOut o{};
  o.by_slot = by_slot ? 1 : 0;
  double* d_pos = (double*)base;
  int32_t* d_i32 = (int32_t*)(base + 8 * n);
  uint8_t* d_u8 = base + 8 * n + 20 * n;
  o.pos = pos ? d_pos : nullptr;
  o.container = container ? d_i32 : nullptr;
  o.agenda_cur = agenda_cur ? d_i32 + n : nullptr;
  o.sleep_left = sleep_left ? d_i32 + 2 * n : nullptr;
  o.route_head = route_head ? d_i32 + 3 * n : nullptr;
  o.gid = gid ? d_i32 + 4 * n : nullptr;
  o.alive = alive ? d_u8 : nullptr;
  o.loc_kind = loc_kind ? d_u8 + n : nullptr;
  o.side = side ? d_u8 + 2 * n : nullptr;
  o.route_status = route_status ? d_u8 + 3 * n : nullptr;
  const unsigned blocks = (unsigned)((ns + 255) / 256);
  if (alive) CUDA_TRY(c, cudaMemsetAsync(d_u8, 0, (size_t)n, c->stream));   // actors without a slot are gone
  if (ns > 0) {
    CUDA_TRY(c, cudaMemsetAsync(c->d_ghost, 0, (size_t)ns, c->stream));
    k_flag_ghosts<<<blocks, 256, 0, c->stream>>>(p, tick & 1, c->d_ghost);
    k_pack_state<<<blocks, 256, 0, c->stream>>>(p, tick & 1, c->d_ghost, o);
    c->launches += 2;
  }
  auto get = [&](void* dst, const void* src, size_t bytes) {
    return dst ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream) : cudaSuccess;
  };
  CUDA_TRY(c, get(pos, d_pos, n * sizeof(double)));
  CUDA_TRY(c, get(container, d_i32, n * sizeof(int32_t)));
  CUDA_TRY(c, get(agenda_cur, d_i32 + n, n * sizeof(int32_t)));
  CUDA_TRY(c, get(sleep_left, d_i32 + 2 * n, n * sizeof(int32_t)));
  CUDA_TRY(c, get(route_head, d_i32 + 3 * n, n * sizeof(int32_t)));
  CUDA_TRY(c, get(gid, d_i32 + 4 * n, n * sizeof(int32_t)));
  CUDA_TRY(c, get(alive, d_u8, n));
  CUDA_TRY(c, get(loc_kind, d_u8 + n, n));
  CUDA_TRY(c, get(side, d_u8 + 2 * n, n));
  CUDA_TRY(c, get(route_status, d_u8 + 3 * n, n));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  CUDA_TRY(c, cudaGetLastError());
  if (order || n_order) {
    // get-actors order (world.scm:24-30): containers by descending registration index, a lane's
    // actors by descending pos-param, a segment's in list order (landing stamps, see k_relink).
    std::vector<SA> sa(ns);
    std::vector<uint32_t> st(ns);
    std::vector<uint8_t> ghost(ns);
    std::vector<int32_t> cont(ns), land_tick(ns), land_lane(ns), ident(ns);
    std::vector<double> ps(ns), land_pos(ns);
    auto pull = [&](void* dst, const void* src, size_t bytes) { return bytes ? cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) : cudaSuccess; };
    CUDA_TRY(c, pull(sa.data(), p.sa, ns * sizeof(SA)));
    for (int32_t a = 0; a < ns; ++a) st[a] = sa[a].st;
    CUDA_TRY(c, pull(ghost.data(), c->d_ghost, ns));
    std::vector<Cold> cold(ns);
    std::vector<ColdF> coldf(ns);
    CUDA_TRY(c, pull(cold.data(), p.cold, ns * sizeof(Cold)));
    CUDA_TRY(c, pull(coldf.data(), p.coldf, ns * sizeof(ColdF)));
    for (int32_t a = 0; a < ns; ++a) {
      cont[a] = cold[a].container; ident[a] = cold[a].gid;
      land_tick[a] = coldf[a].land_tick; land_lane[a] = coldf[a].land_lane; land_pos[a] = coldf[a].land_pos;
    }
    CUDA_TRY(c, pull(ps.data(), p.pos[tick & 1], ns * sizeof(double)));
    std::vector<int32_t> ord;
    ord.reserve(ns);
    for (int32_t a = 0; a < ns; ++a)
      if ((st[a] & ST_ALIVE) && !ghost[a]) ord.push_back(a);
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
    if (order)
      for (size_t k = 0; k < ord.size(); ++k) order[k] = by_slot ? ord[k] : ident[ord[k]];
    if (n_order) *n_order = (int64_t)ord.size();
  }
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
  if (c->mg.on) return fail(c, GRIDDY_ERR_BAD_ARG, "a multi-GPU context is read back with griddy_download_local");
  return download_state(c, false, nullptr, alive, loc_kind, container, side, pos, agenda_cur, sleep_left,
                        route_status, route_head, order, n_order);
}

int griddy_download_local(void* ctx, int64_t capacity, int64_t* n_local, int32_t* gid, uint8_t* alive,
                          uint8_t* loc_kind, int32_t* container, uint8_t* side, double* pos,
                          int32_t* agenda_cur, int32_t* sleep_left, uint8_t* route_status,
                          int32_t* route_head, int32_t* order, int64_t* n_order) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_actors || !n_local) return fail(c, GRIDDY_ERR_BAD_ARG, "download_local: no actors or null n_local");
  *n_local = c->ns_host;
  if (capacity < c->ns_host) return fail(c, GRIDDY_ERR_BAD_ARG, "download_local: arrays too small, need " + std::to_string(c->ns_host));
  return download_state(c, true, gid, alive, loc_kind, container, side, pos, agenda_cur, sleep_left,
                        route_status, route_head, order, n_order);
}

int griddy_download_occupancy(void* ctx, int32_t* lane_count, int32_t* junction_count) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_actors) return fail(c, GRIDDY_ERR_BAD_ARG, "download before upload_actors");
  CUDA_TRY(c, cudaSetDevice(c->device));
  const Params& p = c->p;
  const int32_t ns = c->ns_host;
  std::vector<SA> sa(ns);
  std::vector<uint32_t> st(ns);
  std::vector<int32_t> cont(ns), ahead(ns);
  std::vector<double> ps(ns);
  if (ns) {
    CUDA_TRY(c, cudaMemcpy(sa.data(), p.sa, ns * sizeof(SA), cudaMemcpyDeviceToHost));
    for (int32_t a = 0; a < ns; ++a) { st[a] = sa[a].st; ahead[a] = sa[a].ahead; }
    std::vector<Cold> cold(ns);
    CUDA_TRY(c, cudaMemcpy(cold.data(), p.cold, ns * sizeof(Cold), cudaMemcpyDeviceToHost));
    for (int32_t a = 0; a < ns; ++a) cont[a] = cold[a].container;
    CUDA_TRY(c, cudaMemcpy(ps.data(), p.pos[c->tick & 1], ns * sizeof(double), cudaMemcpyDeviceToHost));
  }
  // ghost rule: an on-road actor whose list follower holds an equal key is not part of the world
  std::vector<uint8_t> ghost(ns, 0);
  for (int32_t a = 0; a < ns; ++a) {
    if ((st[a] & (ST_ALIVE | ST_ONROAD)) != (ST_ALIVE | ST_ONROAD)) continue;
    for (int32_t x = ahead[a]; x >= 0 && ps[x] == ps[a]; x = ahead[x]) ghost[x] = 1;
  }
  if (junction_count) {   // the device's own incrementally maintained counters, less the ghosts
    CUDA_TRY(c, cudaMemcpy(junction_count, p.occ, (size_t)p.n_items * sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (int32_t a = 0; a < ns; ++a)
      if ((st[a] & ST_ALIVE) && ghost[a] && c->h_kind[cont[a]] == GRIDDY_JUNCTION_LANE)
        --junction_count[c->h_lane_junction[cont[a]]];
  }
  if (lane_count) {
    std::fill(lane_count, lane_count + p.n_items, 0);
    for (int32_t a = 0; a < ns; ++a)
      if ((st[a] & ST_ALIVE) && !ghost[a]) ++lane_count[cont[a]];
  }
  return GRIDDY_OK;
}

// ---- multi-GPU set-up --------------------------------------------------------------------------
int griddy_mg_plan(int64_t n_items, const uint8_t* kind, const int32_t* lane_in, const int32_t* lane_out,
                   const int32_t* lane_junction, const int32_t* owner, int32_t owner_rank, int32_t reader_rank,
                   int32_t* lanes, int64_t* n_lanes, int32_t* junctions, int64_t* n_junctions) {
  if (n_items < 0 || !kind || !lane_in || !lane_out || !lane_junction || !owner || !n_lanes || !n_junctions)
    return GRIDDY_ERR_BAD_ARG;
  std::vector<int32_t> l, j;
  mg_plan(n_items, kind, lane_in, lane_out, lane_junction, owner, owner_rank, reader_rank, l, j);
  if (lanes) std::copy(l.begin(), l.end(), lanes);
  if (junctions) std::copy(j.begin(), j.end(), junctions);
  *n_lanes = (int64_t)l.size();
  *n_junctions = (int64_t)j.size();
  return GRIDDY_OK;
}

int griddy_mg_configure(void* ctx, int32_t rank, int32_t world, const int32_t* owner, const int32_t* lane_in,
                        int32_t mig_cap, int64_t extra_slots) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_net) return fail(c, GRIDDY_ERR_BAD_ARG, "mg_configure before upload_network");
  if (world < 1 || world > MAX_RANKS || rank < 0 || rank >= world || !owner || !lane_in || mig_cap < 1 || extra_slots < 0)
    return fail(c, GRIDDY_ERR_BAD_ARG, "mg_configure: bad arguments");
  if (c->p.grid_n == 0) return fail(c, GRIDDY_ERR_BAD_ARG, "mg_configure needs the grid routing table (griddy_set_routing_grid)");
  CUDA_TRY(c, cudaSetDevice(c->device));
  const Params& q = c->p;
  const int64_t n_items = q.n_items;
  for (int64_t r = 0; r < n_items; ++r) {
    if (owner[r] < 0 || owner[r] >= world) return fail(c, GRIDDY_ERR_BAD_ARG, "owner out of range");
    const uint8_t k = c->h_kind[r];
    // a segment and its lanes share an owner, a junction and its lanes too: actors go on and off the
    // road, and junction occupancy is counted, without leaving the rank
    if (k == GRIDDY_JUNCTION_LANE && owner[r] != owner[c->h_lane_junction[r]])
      return fail(c, GRIDDY_ERR_BAD_ARG, "junction lane " + std::to_string(r) + " and its junction have different owners");
  }
  free_pool(c->actor_allocs);
  c->have_actors = false;
  mg_reset(c);
  Multi& m = c->mg;
  m.on = true;
  m.rank = rank;
  m.world = world;
  m.mig_cap = mig_cap;
  m.extra_slots = extra_slots;
  c->mg_owner_host.assign(owner, owner + n_items);
  TRY(mg_alloc(c, &m.d_owner, (size_t)n_items));
  CUDA_TRY(c, cudaMemcpy(m.d_owner, owner, (size_t)n_items * sizeof(int32_t), cudaMemcpyHostToDevice));
  CUDA_TRY(c, cudaEventCreateWithFlags(&m.ev_ready, cudaEventDisableTiming));
  CUDA_TRY(c, cudaEventCreateWithFlags(&m.ev_packed, cudaEventDisableTiming));
  CUDA_TRY(c, cudaEventCreateWithFlags(&m.ev_arrived, cudaEventDisableTiming));
  CUDA_TRY(c, cudaStreamCreateWithFlags(&m.comm_stream, cudaStreamNonBlocking));
  // what I export to / import from every other rank
  struct Lists { std::vector<int32_t> exp_l, exp_j, imp_l, imp_j; };
  std::vector<Lists> lists(world);
  size_t halo_in = 0;
  for (int r = 0; r < world; ++r) {
    if (r == rank) continue;
    Lists& L = lists[r];
    mg_plan(n_items, c->h_kind.data(), lane_in, c->h_lane_out.data(), c->h_lane_junction.data(), owner, rank, r, L.exp_l, L.exp_j);
    mg_plan(n_items, c->h_kind.data(), lane_in, c->h_lane_out.data(), c->h_lane_junction.data(), owner, r, rank, L.imp_l, L.imp_j);
    if (L.exp_l.empty() && L.exp_j.empty() && L.imp_l.empty() && L.imp_j.empty()) continue;
    halo_in += L.imp_l.size() + L.imp_j.size();
  }
  TRY(mg_alloc(c, &m.d_halo_recv, halo_in));
  size_t off = 0;
  for (int r = 0; r < world; ++r) {
    Lists& L = lists[r];
    if (r == rank || (L.exp_l.empty() && L.exp_j.empty() && L.imp_l.empty() && L.imp_j.empty())) continue;
    Neighbor nb;
    nb.rank = r;
    nb.n_exp_lanes = (int)L.exp_l.size();
    nb.n_exp_junc = (int)L.exp_j.size();
    nb.n_imp_lanes = (int)L.imp_l.size();
    nb.n_imp_junc = (int)L.imp_j.size();
    std::vector<int32_t> exp(L.exp_l);
    exp.insert(exp.end(), L.exp_j.begin(), L.exp_j.end());
    TRY(mg_alloc(c, &nb.d_exp_items, exp.size()));
    if (!exp.empty()) CUDA_TRY(c, cudaMemcpy(nb.d_exp_items, exp.data(), exp.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    TRY(mg_alloc(c, &nb.d_imp_junc, L.imp_j.size()));
    if (!L.imp_j.empty()) CUDA_TRY(c, cudaMemcpy(nb.d_imp_junc, L.imp_j.data(), L.imp_j.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    TRY(mg_alloc(c, &nb.d_halo_send, exp.size()));
    nb.d_halo_recv = m.d_halo_recv + off;
    for (size_t k = 0; k < L.imp_l.size(); ++k) m.import_marks.emplace_back(L.imp_l[k], (int32_t)(-2 - (int64_t)(off + k)));
    off += L.imp_l.size() + L.imp_j.size();
    nb.mig_out = (int)std::min<int64_t>(mig_cap, mg_count_feeders(n_items, c->h_kind.data(), lane_in, c->h_lane_out.data(), owner, rank, r));
    nb.mig_in = (int)std::min<int64_t>(mig_cap, mg_count_feeders(n_items, c->h_kind.data(), lane_in, c->h_lane_out.data(), owner, r, rank));
    for (int par = 0; par < 2; ++par) {   // plain cudaMalloc: these are shared with the neighbour through cudaIpc
      TRY(mg_alloc(c, &nb.d_mig_recv[par], nb.in_bytes()));
      CUDA_TRY(c, cudaMemset(nb.d_mig_recv[par], 0, nb.in_bytes()));
      c->p.mig_mine[r][par] = (int32_t*)nb.d_mig_recv[par];
    }
    TRY(mg_alloc(c, &nb.d_token_out, 1));
    TRY(mg_alloc(c, &nb.d_token_in, 1));
    CUDA_TRY(c, cudaMemset(nb.d_token_out, 0, sizeof(int64_t)));
    c->p.mig_cap[r] = nb.mig_out;
    m.nb.push_back(nb);
  }
  c->p.owner = m.d_owner;
  c->p.my_rank = rank;
  c->p.halo_recv = m.d_halo_recv;
  return GRIDDY_OK;
}

int griddy_mg_set_global_ids(void* ctx, const int32_t* gid) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_actors || !gid) return fail(c, GRIDDY_ERR_BAD_ARG, "mg_set_global_ids: no actors or null ids");
  CUDA_TRY(c, cudaSetDevice(c->device));
  for (int64_t a = 1; a < c->p.n_actors; ++a)
    if (gid[a] <= gid[a - 1]) return fail(c, GRIDDY_ERR_BAD_ARG, "global ids must ascend: they carry the order of the link! calls");
  // right after upload_actors the initial reorder has already permuted the slots: map through the
  // upload indices that Cold::gid still holds
  const int32_t ns = c->ns_host;
  std::vector<Cold> cur(ns);
  if (ns) CUDA_TRY(c, cudaMemcpy(cur.data(), c->p.cold, (size_t)ns * sizeof(Cold), cudaMemcpyDeviceToHost));
  for (int32_t s = 0; s < ns; ++s) {
    if (cur[s].gid < 0 || cur[s].gid >= c->p.n_actors) return fail(c, GRIDDY_ERR_BAD_STATE, "mg_set_global_ids must directly follow upload_actors");
    cur[s].gid = gid[cur[s].gid];
  }
  if (ns) CUDA_TRY(c, cudaMemcpy(c->p.cold, cur.data(), (size_t)ns * sizeof(Cold), cudaMemcpyHostToDevice));
  return GRIDDY_OK;
}

int griddy_mg_nccl_unique_id(uint8_t* out128) {
  NcclApi* api = nccl_api();
  if (!api || !out128) return GRIDDY_ERR_NCCL;
  NcclId id;
  if (api->GetUniqueId(&id) != 0) return GRIDDY_ERR_NCCL;
  memcpy(out128, id.internal, 128);
  return GRIDDY_OK;
}

int griddy_mg_connect_nccl(void* ctx, const uint8_t* id128) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->mg.on || !id128) return fail(c, GRIDDY_ERR_BAD_ARG, "mg_connect_nccl before mg_configure");
  NcclApi* api = nccl_api();
  if (!api) return fail(c, GRIDDY_ERR_NCCL, "libnccl.so.2 could not be loaded");
  CUDA_TRY(c, cudaSetDevice(c->device));
  NcclId id;
  memcpy(id.internal, id128, 128);
  NCCL_TRY(c, api->CommInitRank(&c->mg.nccl_comm, c->mg.world, id, c->mg.rank));
  return GRIDDY_OK;
}

int griddy_mg_neighbors(void* ctx, int32_t* ranks, int32_t* n) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->mg.on || !n) return fail(c, GRIDDY_ERR_BAD_ARG, "mg_neighbors before mg_configure");
  *n = (int32_t)c->mg.nb.size();
  if (ranks)
    for (size_t k = 0; k < c->mg.nb.size(); ++k) ranks[k] = c->mg.nb[k].rank;
  return GRIDDY_OK;
}

// One process per GPU: the receive buffers for `from_rank`'s emigrants are handed to that rank as
// cudaIpc handles (2 x 64 bytes, one per tick parity), which it maps with griddy_mg_ipc_import.
int griddy_mg_ipc_export(void* ctx, int32_t from_rank, uint8_t* handles128) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->mg.on || !handles128) return fail(c, GRIDDY_ERR_BAD_ARG, "mg_ipc_export before mg_configure");
  CUDA_TRY(c, cudaSetDevice(c->device));
  for (Neighbor& nb : c->mg.nb)
    if (nb.rank == from_rank) {
      for (int par = 0; par < 2; ++par) {
        cudaIpcMemHandle_t h;
        CUDA_TRY(c, cudaIpcGetMemHandle(&h, nb.d_mig_recv[par]));
        static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
        memcpy(handles128 + 64 * par, &h, 64);
      }
      return GRIDDY_OK;
    }
  return fail(c, GRIDDY_ERR_BAD_ARG, "rank " + std::to_string(from_rank) + " is not a neighbour");
}

int griddy_mg_ipc_import(void* ctx, int32_t to_rank, const uint8_t* handles128) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->mg.on || !handles128) return fail(c, GRIDDY_ERR_BAD_ARG, "mg_ipc_import before mg_configure");
  CUDA_TRY(c, cudaSetDevice(c->device));
  for (Neighbor& nb : c->mg.nb)
    if (nb.rank == to_rank) {
      for (int par = 0; par < 2; ++par) {
        cudaIpcMemHandle_t h;
        memcpy(&h, handles128 + 64 * par, 64);
        void* ptr = nullptr;
        CUDA_TRY(c, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        nb.peer_mig[par] = (uint8_t*)ptr;
        c->p.mig_peer[to_rank][par] = (uint8_t*)ptr;
      }
      return GRIDDY_OK;
    }
  return fail(c, GRIDDY_ERR_BAD_ARG, "rank " + std::to_string(to_rank) + " is not a neighbour");
}

int griddy_mg_connect_local(void** ctxs, int32_t n) {
  if (!ctxs || n < 1 || n > MAX_RANKS) return GRIDDY_ERR_BAD_ARG;
  std::vector<Ctx*> cs((Ctx**)ctxs, (Ctx**)ctxs + n);
  for (int r = 0; r < n; ++r)
    if (!cs[r] || !cs[r]->mg.on || cs[r]->mg.rank != r || cs[r]->mg.world != n)
      return cs[r] ? fail(cs[r], GRIDDY_ERR_BAD_ARG, "mg_connect_local: contexts must be configured as ranks 0..n-1, in order") : GRIDDY_ERR_BAD_ARG;
  for (Ctx* c : cs) {
    c->mg.peers = cs;
    for (Ctx* d : cs)
      if (d->device != c->device) {
        cudaSetDevice(c->device);
        const cudaError_t e = cudaDeviceEnablePeerAccess(d->device, 0);
        cudaGetLastError();
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return fail(c, GRIDDY_ERR_CUDA, "no peer access between devices " + std::to_string(c->device) + " and " + std::to_string(d->device));
      }
    // same process: the neighbours' receive buffers are directly addressable
    for (Neighbor& nb : c->mg.nb)
      for (Neighbor& back : cs[nb.rank]->mg.nb)
        if (back.rank == c->mg.rank)
          for (int par = 0; par < 2; ++par) c->p.mig_peer[nb.rank][par] = nb.peer_mig[par] = back.d_mig_recv[par];
  }
  return GRIDDY_OK;
}

int griddy_count_alive(void* ctx, int64_t* n_alive) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_actors || !n_alive) return fail(c, GRIDDY_ERR_BAD_ARG, "count_alive: no actors or null output");
  CUDA_TRY(c, cudaSetDevice(c->device));
  *n_alive = 0;
  const int32_t ns = c->ns_host;
  if (ns == 0) return GRIDDY_OK;
  const unsigned blocks = (unsigned)((ns + 255) / 256);
  unsigned long long* d_out = (unsigned long long*)c->sort_keys;   // reorder scratch, idle between steps
  CUDA_TRY(c, cudaMemsetAsync(d_out, 0, sizeof(unsigned long long), c->stream));
  CUDA_TRY(c, cudaMemsetAsync(c->d_ghost, 0, (size_t)ns, c->stream));
  k_flag_ghosts<<<blocks, 256, 0, c->stream>>>(c->p, (int)(c->tick & 1), c->d_ghost);
  k_count_alive<<<blocks, 256, 0, c->stream>>>(c->p, c->d_ghost, d_out);
  c->launches += 2;
  unsigned long long h = 0;
  CUDA_TRY(c, cudaMemcpyAsync(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  *n_alive = (int64_t)h;
  return GRIDDY_OK;
}

// Diagnostics: the reorder's radix sort on caller data (host pointers), for tests/.
int griddy_debug_sort_pairs(uint64_t* keys, uint32_t* vals, int64_t n, int32_t key_bits) {
  if (!keys || !vals || n < 0 || key_bits < 1 || key_bits > 64) return GRIDDY_ERR_BAD_ARG;
  if (rsort::prepare() != cudaSuccess) return GRIDDY_ERR_CUDA;
  const int n_tiles = (int)((n + rsort::TILE - 1) / rsort::TILE);
  if (n_tiles == 0) return GRIDDY_OK;
  const size_t n_pad = (size_t)n_tiles * rsort::TILE;
  void* buf[6] = {};
  const size_t bytes[6] = {n_pad * 8, n_pad * 8, n_pad * 4, n_pad * 4, rsort::hist_words(n_tiles) * 4,
                           rsort::scan_blocks(rsort::hist_words(n_tiles)) * 4};
  bool ok = true;
  for (int k = 0; k < 6 && ok; ++k) ok = cudaMalloc(&buf[k], bytes[k]) == cudaSuccess;
  if (ok) {
    uint64_t* kk = (uint64_t*)buf[0];
    uint32_t* vv = (uint32_t*)buf[2];
    rsort::Scratch sc;
    sc.keys_alt = (uint64_t*)buf[1];
    sc.vals_alt = (uint32_t*)buf[3];
    sc.hist = (uint32_t*)buf[4];
    sc.sums = (uint32_t*)buf[5];
    cudaMemset(kk, 0xFF, n_pad * 8);   // padding sorts last
    cudaMemset(vv, 0xFF, n_pad * 4);
    cudaMemcpy(kk, keys, (size_t)n * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(vv, vals, (size_t)n * 4, cudaMemcpyHostToDevice);
    rsort::sort_pairs(&kk, &vv, n_tiles, key_bits, &sc, 0);
    ok = cudaDeviceSynchronize() == cudaSuccess;
    cudaMemcpy(keys, kk, (size_t)n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(vals, vv, (size_t)n * 4, cudaMemcpyDeviceToHost);
  }
  for (void* ptr : buf) cudaFree(ptr);
  return ok ? GRIDDY_OK : GRIDDY_ERR_CUDA;
}

double griddy_last_step_ms(void* ctx) { return ctx ? ((Ctx*)ctx)->last_ms : 0.0; }
int64_t griddy_tick(void* ctx) { return ctx ? ((Ctx*)ctx)->tick : 0; }
int64_t griddy_kernel_launches(void* ctx) { return ctx ? ((Ctx*)ctx)->launches : 0; }
int64_t griddy_slots_in_use(void* ctx) { return ctx ? ((Ctx*)ctx)->ns_host : 0; }

}  // extern "C"
