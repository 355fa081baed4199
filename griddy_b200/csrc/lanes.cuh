// lanes.cuh — part of griddy_cuda.cu (one translation unit).  Per-lane order maintenance: processing order, k_unlink, k_link.
#pragma once

namespace griddy {

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
  const uint32_t sta = p.hot[a].st, stb = p.hot[b].st;
  const bool nega = (((tick - p.coldf[a].land_tick) & 1) != 0) == ((sta & ST_CLASS_P) != 0);
  const bool negb = (((tick - p.coldf[b].land_tick) & 1) != 0) == ((stb & ST_CLASS_P) != 0);
  if (nega != negb) return nega;
  return nega ? stamp_less(p, b, a, stb) : stamp_less(p, a, b, sta);
}

// Of two actors landing on the same (lane, pos-param) this tick: the one visited later, which is
// the one bbtree-set keeps (core.scm:173-178).
__device__ int later_processed(const Params& p, int x, int y, int32_t lane, const double* pos_old, int tick) {
  const uint32_t sx = p.hot[x].st, sy = p.hot[y].st;
  const int32_t mx = (sx & ST_MOVED) ? p.cold[x].mv_old : lane, my = (sy & ST_MOVED) ? p.cold[y].mv_old : lane;
  const int32_t cx = mx < 0 ? ~mx : mx, cy = my < 0 ? ~my : my;
  if (cx != cy) return cx < cy ? x : y;
  if (mx < 0) return offroad_before(p, x, y, tick) ? y : x;
  return pos_old[x] < pos_old[y] ? x : y;
}

// ------------------------------------------------------------------------------------------------
// k_unlink: one thread per actor that leaves its lane's list this tick; k_link: one per lane that an actor entered.
// Lanes are doubly linked lists in ascending pos-param, so nothing walks a whole lane.
//   k_unlink  takes every leaver (k_slow's records: actors that changed lane, went off the road or jumped back, and
//             ghosts) out of its list and settles the actors that are done for this tick: a ghost's move is void, an
//             emigrant lives on another rank from now on, an actor that left the road has no lane to enter.  No
//             per-lane state is involved: the leavers of a lane form RUNS of list neighbours (nearly always of one),
//             a leaver's own pointers stay as they were, and the foremost leaver of a run — the one whose leader
//             stays — walks back over the run and joins the two residents around it.  Two runs never write the
//             same word: they are separated by a resident, of which one run writes `ahead` and the other `behind`.
//   k_link    (after every leaver is out, so a mover's pointers are free to overwrite) inserts the lane's entrants
//             (its inbox, sorted) by walking up from the rear — entrants land near pos 0, so that is a step or
//             two — and writes their final pointers.
// The usual lane change — the front actor leaves, lands at the rear of the next lane — touches the hot records of the
// mover and of its two new / old neighbours and the new lane's record; LaneRec::meta says how many entrants were
// pushed, so the chain through the cold records is followed only when there are several.
// Residents keep their relative order (a follower is clamped behind its leader, route.scm:68-74).
// Equal keys: an entrant that lands on a resident's or another entrant's key is resolved in k_link,
// the later-processed one stays (core.scm:173-178); two residents that round to the same key are left
// linked and resolved next tick through the ghost rule (file header) — on every lane, touched or not.
// Junction occupancy (route.scm:98-104): one atomic per leaver of a junction lane, one per lane that gained actors.

// Is the actor with these flags leaving its lane's list this tick (or gone already)?  Every linked actor is alive and on
// the road; k_slow marks movers and ghosts, and what k_unlink itself clears (below) leaves one of the other marks.
__device__ __forceinline__ bool is_leaver(uint32_t st) {
  return (st & (ST_MOVED | ST_GHOST)) || (st & (ST_ALIVE | ST_ONROAD)) != (ST_ALIVE | ST_ONROAD);
}

// immig_blocks > 0 (a partitioned world, one process per GPU): the first immig_blocks x immig_ranks blocks unpack the
// actors that arrived from the other ranks instead (multi_gpu.cuh, immig_unpack_block).
__global__ void __launch_bounds__(128) k_unlink(Params p, int tick, int immig_blocks, int immig_ranks) {
  grid_dependency_wait();
  TRACE(p, tick, TR_UNLINK);
  const int par = tick & 1;
  // The FIRST immig_blocks x immig_ranks blocks: block 0 raises my migrant flags (k_slow is complete), then each waits
  // for its neighbour's — whose block 0 has done the same before it waits for anyone — and unpacks beside the unlinking.
  const int n_immig = immig_blocks * immig_ranks;
  if (p.publish_late && blockIdx.x == 0) emig_publish_late(p, tick);
  if ((int)blockIdx.x < n_immig) {
    immig_unpack_block(p, tick, (int)blockIdx.x / immig_blocks, (int)blockIdx.x % immig_blocks, immig_blocks);
    return;
  }
  const int n_blocks = (int)gridDim.x - n_immig;   // blocks that unlink
  const int32_t n_leavers = p.counters[CNT_STRIDE * par + CNT_LEFT];
  for (int32_t i = ((int)blockIdx.x - n_immig) * blockDim.x + threadIdx.x; i < n_leavers; i += n_blocks * blockDim.x) {
    const int4 rec = p.leavers[i];            // slot, lane, junction
    const int32_t z = rec.x;
    const int4 h = *(const int4*)&p.hot[z];   // st, ahead, behind, tj: the pointers are the old world's
    const uint32_t sz = (uint32_t)h.x;
    const int32_t f = h.y;
    if (!(f >= 0 && is_leaver(p.hot[f].st))) {   // the foremost of its run
      int32_t b = h.z;
      while (b >= 0) {
        const int4 hb = *(const int4*)&p.hot[b];
        if (!is_leaver((uint32_t)hb.x)) break;
        b = hb.z;
      }
      if (b >= 0) p.hot[b].ahead = f; else p.lane[rec.y].tail = f;
      if (f >= 0) p.hot[f].behind = b;
    }
    // every mover leaves its old lane, whatever becomes of it
    bool died = false;
    if (sz & (ST_GHOST | ST_EMIG)) { p.hot[z].st = sz & ~(ST_MOVED | ST_ALIVE | ST_EMIG); died = true; }
    else if (!(sz & ST_ONROAD)) p.hot[z].st = sz & ~ST_MOVED;   // went off the road: nothing to link
    if (rec.z >= 0) atomicSub(&p.occ[rec.z], 1);
    if (died) atomicAdd(&p.scal[SC_NDEAD], 1);   // dead slots pile up until the next reorder drops them
  }
}

// After this tick's immigrants are unpacked (k_link runs behind k_immig_unpack): nothing to reset — the
// sender overwrites header and records two ticks from now, after it has received the halo I send next tick.

__global__ void __launch_bounds__(128) k_link(Params p, int tick) {
  grid_dependency_wait();
  TRACE(p, tick, TR_LINK);
  const int par = tick & 1;
  const int32_t n_touched = p.counters[CNT_STRIDE * par + CNT_ENTERED];
  const double* __restrict__ pos_old = p.pos[par];
  const double* __restrict__ pos_new = p.pos[par ^ 1];
  const int32_t* __restrict__ entered = p.touched;
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_touched; i += gridDim.x * blockDim.x) {
    const int32_t lane = entered[i];
    LaneRec* lr = &p.lane[lane];
    const int4 links = *(const int4*)&lr->len;    // len, link, link2
    const int4 first = *(const int4*)&lr->to_j;   // to_j, meta, e_pos
    const uint32_t meta = (uint32_t)first.y;
    const double lone_pos = __hiloint2double(first.w, first.z);
    const int4 heads = *(const int4*)&lr->tail;   // tail, inbox, outbox, ghosts
    const int32_t inbox = heads.y;
    const bool single = !(meta & META_MULTI);
    if (inbox < 0) continue;
    lr->inbox = -1;
    if (single) {
      // One entrant, no tie — nearly every lane: its key came with the lane's record, its hot record is written once.
      const int4 he = *(const int4*)&p.hot[inbox];   // st, ahead, behind, tj
      if (!((uint32_t)he.x & ST_MOVED)) continue;    // a ghost's move is void: k_unlink settled it
      int32_t prev = -1, x = heads.x;
      double px = 0.0;
      while (x >= 0 && (px = pos_new[x]) < lone_pos) { prev = x; x = p.hot[x].ahead; }
      if (!(x >= 0 && px == lone_pos)) {
        if (prev < 0) lr->tail = inbox; else p.hot[prev].ahead = inbox;
        *(int4*)&p.hot[inbox] = make_int4((int)((uint32_t)he.x & ~(ST_MOVED | ST_IMMIG)), x, prev, he.w);
        if (x >= 0) p.hot[x].behind = inbox;
        if (meta_kind(meta) == GRIDDY_JUNCTION_LANE) atomicAdd(&p.occ[links.w], 1);
        continue;
      }
    } else {
      atomicAnd(&lr->meta, ~META_MULTI);
    }
    auto next_in = [&](int32_t y) { return single ? -1 : p.cold[y].inbox_next; };
    // entrants by ascending (pos-param, slot): selection over the inbox, which is short
    double e_pos = 0.0;
    int32_t e = -1;
    auto next_entrant = [&](bool first) {
      int32_t best = -1;
      double best_pos = 0.0;
      for (int32_t y = inbox; y >= 0; y = next_in(y)) {
        if (!(p.hot[y].st & ST_MOVED)) continue;   // a ghost's move is void: k_unlink settled it
        const double py = single ? lone_pos : pos_new[y];   // (one entrant: its key came with the lane's record)
        if (!first && !(py > e_pos || (py == e_pos && y > e))) continue;
        if (best < 0 || py < best_pos || (py == best_pos && y < best)) { best = y; best_pos = py; }
      }
      e = best;
      e_pos = best_pos;
    };
    int32_t tail = heads.x;
    auto link = [&](int32_t b, int32_t z, int32_t f) {   // b -> z -> f
      if (b < 0) tail = z; else p.hot[b].ahead = z;
      p.hot[z].ahead = f;
      p.hot[z].behind = b;
      if (f >= 0) p.hot[f].behind = z;
    };
    int32_t prev = -1, x = heads.x, delta = 0, died = 0;
    for (next_entrant(true); e >= 0; next_entrant(false)) {
      while (x >= 0 && pos_new[x] < e_pos) { prev = x; x = p.hot[x].ahead; }
      if (x >= 0 && pos_new[x] == e_pos) {   // equal key: bbtree-set keeps the one processed later
        const bool e_wins = later_processed(p, e, x, lane, pos_old, tick) == e;
        const int32_t lose = e_wins ? x : e;
        const uint32_t sl = p.hot[lose].st;
        p.hot[lose].st = sl & ~ST_ALIVE;        // (still marked moved if it is an entrant: see below)
        ++died;
        if (!(sl & ST_MOVED) || lose == x) --delta;   // a resident, or an entrant that had been counted in
        if (!e_wins) { if (single) break; continue; }
        link(prev, e, p.hot[x].ahead);   // e takes x's place
      } else {
        link(prev, e, x);
      }
      ++delta;
      x = e;   // the next entrant (same or higher key) meets e first
      if (single) break;
    }
    // the entrants stayed marked as movers so that ties among them could look at where they came from
    for (int32_t y = inbox; y >= 0; y = next_in(y)) {
      const uint32_t sy = p.hot[y].st;
      if (sy & ST_MOVED) p.hot[y].st = sy & ~(ST_MOVED | ST_IMMIG);
    }
    if (tail != heads.x) lr->tail = tail;
    if (delta && meta_kind(meta) == GRIDDY_JUNCTION_LANE) atomicAdd(&p.occ[links.w], delta);
    if (died) atomicAdd(&p.scal[SC_NDEAD], died);
  }
}

}  // namespace griddy
