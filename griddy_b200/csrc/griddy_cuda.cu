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
// Data layout (HBM; records.cuh).  Actors live in SLOTS; a slot's number is not the actor's id.  Every
// reorder_period ticks — and whenever a step begins with more than 1/16 of the slots dead — the live actors are
// radix-sorted by (locality key of their lane or segment, coarse pos-param) and all per-slot arrays are permuted
// (radix_sort.cuh, reorder.cuh): the actors of a lane then sit next to each other in ascending pos-param,
// neighbouring lanes sit in neighbouring cache lines, and dead actors are dropped.  Between reorders slots are
// stable; order within a lane is carried exactly by the ahead / behind pointers, so the sort only has to be
// good for locality, never for correctness.
//   Hot      32 B per slot, read by every actor every tick: st (alive | on_road | side | route status | agenda head kind
//            | moved | stamp class | route head is (arrive-at _) | ghost), ahead / behind (neighbours on the lane), tj
//            (junction gating the route head; asleep: remaining sleep time), delta / buf (per-tick advance and tailgate
//            buffer on the current lane, route.scm:62-66, cached at lane entry)
//   pos[2]   f64 pos-param, ping-pong by tick parity: followers read the old buffer
//   Cold     64 B per slot, touched only when an actor changes lane, wakes, or starts / ends a route: container, route
//            head, cached destination, list-surgery links, speed, cursors, actor id;  ColdF: arrive target, landing stamp
//   LaneRec  64 B per registered item: length, links, marks, this tick's entrant (inbox head, its key), rearmost actor,
//            routing-table row
//   occ      i32 per junction: actors on its lanes
// Per-lane order changes only where an actor entered or left a lane this tick; nothing is O(lanes)
// per tick (lanes outnumber actors 1.5:1 to 4:1 on the grids).
//
// Kernels per tick (advance.cuh, lanes.cuh; multi_gpu.cuh for a partitioned world)
//   k_advance  one thread per two slots: agenda dispatch + motion for the common cases (on the lane, waiting for a
//              full junction, asleep, idle).  Coalesced reads of Hot and pos; the leader's pos-param comes from the
//              block's shared-memory tile or a gather that the slot order keeps close.  Everything else is deferred.
//   k_slow     one thread per deferred actor: lane changes with carry-over, arrive-at, begin / end of a route, waking,
//              ghost detection.  Movers are pushed on their target lane's inbox (one atomic exchange); every actor that
//              leaves a lane's list is put on the leaver list.  The lane an actor leaves is not touched.
//   k_unlink   one thread per leaver.  Lanes are doubly linked lists: the foremost leaver of a run of list neighbours
//              joins the residents around the run; actors that are done for the tick are settled.
//   k_link     one thread per lane an actor entered: a lone entrant's key comes with the lane's record and it is linked
//              behind the lane's rearmost actor with one write of its hot record; several entrants are inserted by a
//              short walk up from the rear; an entrant that lands on an occupied key is resolved exactly as bbtree-set would.
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
// Every on-road actor reads its leader's key each tick, sees the tie, flags the ghost (ST_GHOST) and lists it as
// a leaver; whatever the ghost itself did this tick is discarded by k_unlink.
// Nothing else can observe a ghost: queries reach the follower first and see the same key, and the
// junction it may stand on is occupied by the follower as well.  Downloads apply the same rule.
//
// fp64 parity: every operation of route.scm:53-135 is issued with round-to-nearest intrinsics
// (__dadd_rn / __dmul_rn / __ddiv_rn), which the compiler never contracts into FMAs.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "griddy_cuda.h"
#include "radix_sort.cuh"

// The device code lives in the .cuh files next to this one, included in dependency order:
#include "records.cuh"
#include "advance.cuh"
#include "multi_gpu.cuh"
#include "lanes.cuh"
#include "setup.cuh"
#include "reorder.cuh"
#include "readback.cuh"
#include "routes.cuh"

using namespace griddy;

namespace {

// ================================================================================================
// Host side

struct Ctx;

// One neighbouring rank: what I export to it and import from it every tick, and where in the two windows it goes.
struct Neighbor {
  int rank = -1;
  int n_exp_lanes = 0, n_exp_junc = 0, n_imp_lanes = 0, n_imp_junc = 0;
  int exp_first = 0;                // my export list: entries [exp_first, exp_first + n_exp_lanes + n_exp_junc)
  int mig_out = 0, mig_in = 0;      // records per tick at most: a lane lets one actor out per tick
  int64_t item_out = 0, item_in = 0;   // agenda items per tick at most
  // what this neighbour writes into MY window, by tick parity (byte offsets)
  size_t halo_off[2] = {0, 0}, hflag_off[2] = {0, 0}, mig_off[2] = {0, 0};
  // its window as mapped here, and where I write in it
  uint8_t* peer_base = nullptr;
  size_t peer_halo_off[2] = {0, 0}, peer_hflag_off[2] = {0, 0}, peer_mig_off[2] = {0, 0};
  bool mapped = false;
  size_t mig_bytes_in() const { return sizeof(MigHeader) + (size_t)mig_in * sizeof(MigRec) + (size_t)item_in * sizeof(AgendaItem); }
};

struct Multi {
  bool on = false;
  int rank = 0, world = 1;
  int mig_cap = 16384;
  int max_agenda_items = 4;
  int64_t extra_slots = 0;
  std::vector<Neighbor> nb;
  std::vector<void*> allocs;
  std::vector<std::pair<int32_t, int32_t>> import_marks;   // (remote lane, -2 - index into halo_recv)
  uint8_t* window = nullptr;          // what my neighbours write into (one allocation, shared through cudaIpc)
  size_t window_bytes = 0;
  size_t halo_base[2] = {0, 0};       // byte offset of the halo values of either parity (all neighbours back to back)
  int32_t* d_exp_items = nullptr;     // my export list: my lane, or ~(my junction), by neighbour
  int n_exp = 0;
  int32_t *d_imp_junc = nullptr, *d_imp_where = nullptr;   // the neighbours' junctions my actors read, and where in halo_recv
  int n_imp_junc = 0;
  int32_t* d_owner = nullptr;
  bool connected = false;             // every neighbour's window is mapped
  std::vector<Ctx*> peers;            // local transport: every rank is a context of this process
  cudaEvent_t ev_ready = nullptr;     //   "what I write into my neighbours' windows in this phase is enqueued"
};

// Which items this context holds (griddy_set_item_ranges).  Ids are global registration indices everywhere;
// the per-item device arrays are sparse (virtual address space for every id, memory behind the held ranges
// only) and the host-side copies are compact, in range order.
struct Tiling {
  int64_t n_global = 0, n_local = 0;
  std::vector<int64_t> beg, end, off;   // ascending, disjoint; off = compact index of each range's first item
  bool tiled() const { return !(beg.size() == 1 && beg[0] == 0 && end[0] == n_global) && n_global > 0; }
  void whole(int64_t n) { n_global = n_local = n; beg.assign(1, 0); end.assign(1, n); off.assign(1, 0); }
  int64_t local(int64_t g) const {   // compact index of a held item, -1 if it is not held
    size_t lo = 0, hi = beg.size();
    while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      if (g < beg[mid]) hi = mid; else if (g >= end[mid]) lo = mid + 1; else return off[mid] + (g - beg[mid]);
    }
    return -1;
  }
};

// A per-item device array over the GLOBAL id space.  Untiled: one cudaMalloc.  Tiled: a virtual address
// range for all ids with physical memory mapped behind the held ranges only (CUDA virtual memory management,
// bound through cudaGetDriverEntryPoint so that the library does not link libcuda).
struct SparseArray {
  void* base = nullptr;
  size_t va_bytes = 0;
  bool vmm = false;
  std::vector<std::pair<size_t, size_t>> spans;   // mapped (offset, bytes)
  std::vector<unsigned long long> handles;
};

struct Ctx {
  int device = 0;
  std::string err;
  std::vector<void*> net_allocs, actor_allocs;
  std::vector<SparseArray> net_sparse;
  Tiling tl, tl_next;      // tl_next: set by griddy_set_item_ranges for the next griddy_upload_network
  bool have_ranges = false;
  Params p{};
  bool have_net = false, have_actors = false;
  int64_t tick = 0, launches = 0, reorders = 0;
  double last_ms = 0.0;
  int32_t* h_scal = nullptr;        // pinned mirror of Params::scal for griddy_step_async
  cudaEvent_t ev_scal = nullptr;
  bool scal_pending = false;
  int64_t ticks_after_scal = 0;     // ticks enqueued since that copy was: arrivals they may bring are not in it
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // slots
  int32_t ns_host = 0;         // upper bound of the slots in use (refreshed after every step)
  int32_t dead_host = 0;       // slots whose actor is gone, as of the last step's end
  int reorder_period = 384;
  int64_t last_reorder_tick = -1;
  int loc_key_bits = 1;
  Hot *hot_cur = nullptr, *hot_alt = nullptr;       // double buffers of the permuted arrays
  double* pos_alt = nullptr;                        // (the two pos-param buffers of the tick and this one rotate)
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
  // drawing read-back (griddy_download_positions, griddy_positions_begin / _end)
  const double2* d_geom = nullptr;          // 4 x double2 per item (griddy_upload_geometry)
  cudaStream_t draw_stream = nullptr;        // carries the copies of finished frames while the next ticks run
  float2* d_xy32[2] = {nullptr, nullptr};    // two frames in flight at most
  uint8_t* d_xy_sync = nullptr;              // staging of griddy_download_positions: [xy64 16][xy32 8][gid 4] per slot
  cudaEvent_t ev_drawn[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  int64_t frame_n[2] = {0, 0};
  uint64_t frames_begun = 0, frames_ended = 0;
  // delta frames (griddy_delta_begin / _end): what was last sent per slot, two staging sets, the counters the frame
  // kernel hands room out with, and the pinned words its last block writes the totals into
  unsigned long long* d_drawn = nullptr;
  float2* d_dvals[2] = {nullptr, nullptr};
  uint32_t *d_dmask[2] = {nullptr, nullptr}, *d_doff[2] = {nullptr, nullptr};
  unsigned int* d_dcnt = nullptr;            // [2 frames][2]
  unsigned int* h_dcnt = nullptr;            // [2 frames], page-locked and mapped
  cudaEvent_t ev_delta[2] = {nullptr, nullptr}, ev_dcopied[2] = {nullptr, nullptr};
  uint64_t deltas_flushed = 0;               // frames whose copies have been enqueued
  cudaEvent_t ev_fmt0[2] = {nullptr, nullptr};   // with ev_delta: device time of the frame kernel (griddy_last_frame_ms)
  cudaEvent_t ev_copy0[2] = {nullptr, nullptr};  // with ev_dcopied: time of the frame's copies on their stream
  double last_frame_ms = 0.0, last_copy_ms = 0.0;
  int64_t delta_n[2] = {0, 0};
  uint8_t* delta_dst[2] = {nullptr, nullptr};
  int64_t delta_ns[2] = {0, 0}, delta_cap[2] = {0, 0}, delta_tick[2] = {0, 0};
  int delta_key[2] = {0, 0};
  uint64_t deltas_begun = 0, deltas_ended = 0;
  int64_t delta_reorders_seen = -1;          // reorders when the last delta frame was formatted (-1: none yet)
  int32_t delta_cover = 0;                   // slots below this may hold something other than DRAWN_NONE in d_drawn
  // host copies of the held items (compact, range order) that uploads validate against
  std::vector<uint8_t> h_kind, h_dir;
  std::vector<int32_t> h_lane_junction, h_lane_out, h_of, h_ob, h_from, h_to;
  // travel-to destinations resolved by the caller (griddy_set_agenda_destinations), for the next upload_actors
  std::vector<int32_t> dst_lane, dst_from, dst_to;
  std::vector<uint8_t> dst_dir;
  bool have_dst = false;
  // find-route on the device (griddy_set_route_graph): the lane graph, kept on the device with the network
  const int64_t* g_nbr_off = nullptr;
  const int32_t* g_nbr = nullptr;
  const double* g_cost = nullptr;
  const double2* g_mid = nullptr;
  int64_t g_edges = 0;
  std::vector<int64_t> h_route_off;     // the routes the last upload_actors used (given or computed), for griddy_download_routes
  std::vector<int32_t> h_route_lane;
  // explicit routes on a partitioned world (griddy_mg_set_routes): the whole world's table, and the route of every
  // agenda item of the next upload
  std::vector<int64_t> mg_route_off;
  std::vector<int32_t> mg_route_lane, mg_item_route;
  bool have_mg_routes = false;
  AgendaItem* agenda_alt = nullptr;   // second item pool: a reorder of a partitioned world compacts into it
  int32_t* d_nitems_new = nullptr;
  Multi mg;
  std::vector<int32_t> mg_owner_host;   // compact
  uint8_t kind_of(int64_t g) const { const int64_t l = tl.local(g); return l < 0 ? (uint8_t)255 : h_kind[(size_t)l]; }
  int32_t owner_of(int64_t g) const { const int64_t l = tl.local(g); return l < 0 ? -1 : mg_owner_host[(size_t)l]; }
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
  CUDA_TRY(c, cudaMemsetAsync(&p.scal[SC_NDEAD], 0, sizeof(int32_t), s));
  c->dead_host = 0;
  CUDA_TRY(c, cudaMemsetAsync(c->newslot, 0xFF, (size_t)c->ns_host * sizeof(int32_t), s));
  const int key_bits = c->loc_key_bits + 17;   // [off-road][locality key][pos16]
  k_make_keys<<<n_pad / 256, 256, 0, s>>>(p, par, key_bits, c->sort_keys, c->sort_vals, c->tailflag);
  c->launches += 1 + rsort::sort_pairs(&c->sort_keys, &c->sort_vals, n_tiles, key_bits, &c->sort_scratch, s);
  k_inverse<<<n_pad / 256, 256, 0, s>>>(c->sort_vals, n_pad, c->newslot);
  Permute q;
  q.src_hot = c->hot_cur; q.dst_hot = c->hot_alt;
  q.src_pos = p.pos[par]; q.dst_pos = c->pos_alt;   // the valid pos buffer is the one of parity `par`
  q.src_cold = c->cold_cur; q.dst_cold = c->cold_alt;
  q.src_coldf = c->coldf_cur; q.dst_coldf = c->coldf_alt;
  // a partitioned world: the surviving actors' remaining agenda items move to the second pool, so that the
  // room taken by the items of actors that have left (or popped them) is reclaimed
  q.agenda_dst = c->agenda_alt;
  q.n_items_new = c->d_nitems_new;
  if (c->agenda_alt) CUDA_TRY(c, cudaMemsetAsync(c->d_nitems_new, 0, sizeof(int32_t), s));
  k_permute<<<n_pad / 256, 256, 0, s>>>(p, q, c->sort_vals, c->newslot, c->tailflag);
  c->launches += 2;
  CUDA_TRY(c, cudaMemcpyAsync(&p.scal[SC_NSLOTS], &p.scal[SC_NALIVE], sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  if (c->agenda_alt) {
    CUDA_TRY(c, cudaMemcpyAsync(&p.scal[SC_NITEMS], c->d_nitems_new, sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    std::swap(p.agenda, c->agenda_alt);
  }
  std::swap(c->hot_cur, c->hot_alt);
  std::swap(c->cold_cur, c->cold_alt);
  std::swap(c->coldf_cur, c->coldf_alt);
  std::swap(p.pos[par], c->pos_alt);
  p.hot = c->hot_cur;
  p.cold = c->cold_cur;
  p.coldf = c->coldf_cur;
  c->last_reorder_tick = tick;
  ++c->reorders;
  CUDA_TRY(c, cudaGetLastError());
  return GRIDDY_OK;
}

void mg_reset(Ctx* c) {
  Multi& m = c->mg;
  for (Neighbor& nb : m.nb)
    if (nb.mapped && nb.peer_base && m.peers.empty()) cudaIpcCloseMemHandle(nb.peer_base);
  for (void* ptr : m.allocs) cudaFree(ptr);
  if (m.ev_ready) cudaEventDestroy(m.ev_ready);
  m = Multi();
  c->mg_owner_host.clear();
  c->have_mg_routes = false;
  c->mg_route_off.clear();
  c->mg_route_lane.clear();
  c->mg_item_route.clear();
  Params& p = c->p;
  p.owner = nullptr;
  p.halo_recv[0] = p.halo_recv[1] = nullptr;
  p.my_rank = 0;
  for (int r = 0; r < MAX_RANKS; ++r) {
    for (int q = 0; q < 2; ++q) {
      p.halo_peer[r][q] = nullptr;
      p.hflag_peer[r][q] = nullptr;
      p.hflag_mine[r][q] = nullptr;
      p.mig_peer[r][q] = nullptr;
      p.mig_mine[r][q] = nullptr;
    }
    p.halo_first[r] = p.halo_n[r] = 0;
    p.mig_cap[r] = p.mig_icap[r] = p.mig_in[r] = 0;
  }
  p.halo_done = p.mig_done = p.mig_count = nullptr;
  p.halo_items = nullptr;
  p.halo_total = 0;
}

// What a rank (or the host-only helper griddy_mg_plan) sees of the network: compact arrays over the held
// items, ids inside them global.  An item that is not held has no known owner (-1), which matches no rank.
struct NetView {
  const Tiling* tl;
  const uint8_t* kind;
  const int32_t *lane_in, *lane_out, *lane_junction, *owner;
  int32_t own(int64_t g) const { const int64_t l = g < 0 ? -1 : tl->local(g); return l < 0 ? -1 : owner[l]; }
  template <typename F>
  void each(F f) const {   // f(compact index, global id), ascending
    for (size_t k = 0; k < tl->beg.size(); ++k)
      for (int64_t g = tl->beg[k]; g < tl->end[k]; ++g) f(tl->off[k] + (g - tl->beg[k]), g);
  }
};

// Items of `owner_rank` that actors living on `reader_rank` read before they cross: the lanes they
// can enter (a junction lane fed by one of the reader's segment lanes; a segment lane that one of
// the reader's junction lanes leads onto) and the junctions whose occupancy gates the entry.  Both
// sides derive the same ascending lists — each from the items it holds, which include everything next to
// its own part — so no list is ever exchanged.
void mg_plan(const NetView& v, int owner_rank, int reader_rank, std::vector<int32_t>& lanes, std::vector<int32_t>& junctions) {
  std::vector<uint8_t> mark((size_t)v.tl->n_local, 0);
  v.each([&](int64_t l, int64_t) {
    if (v.kind[l] != GRIDDY_JUNCTION_LANE) return;
    const int32_t in = v.lane_in[l], out = v.lane_out[l];
    if (v.owner[l] == owner_rank && v.own(in) == reader_rank) {
      mark[(size_t)l] = 1;
      const int64_t j = v.tl->local(v.lane_junction[l]);
      if (j >= 0) mark[(size_t)j] = 2;
    }
    if (v.owner[l] == reader_rank && v.own(out) == owner_rank) mark[(size_t)v.tl->local(out)] = 1;
  });
  lanes.clear();
  junctions.clear();
  v.each([&](int64_t l, int64_t g) {
    if (mark[(size_t)l] == 1) lanes.push_back((int32_t)g);
    else if (mark[(size_t)l] == 2) junctions.push_back((int32_t)g);
  });
}

// Lanes of `from` with a successor lane on `to`.  A lane lets at most one actor out per tick (the
// follower of an actor that reaches the end is a tail-gate buffer behind, route.scm:62-74), so this
// bounds the actors that cross from `from` to `to` in a tick — and sizes the exchange exactly.
int64_t mg_count_feeders(const NetView& v, int from, int to) {
  std::vector<uint8_t> mark((size_t)v.tl->n_local, 0);
  int64_t n = 0;
  v.each([&](int64_t l, int64_t) {
    if (v.kind[l] != GRIDDY_JUNCTION_LANE) return;
    const int32_t in = v.lane_in[l], out = v.lane_out[l];
    if (v.owner[l] == to && v.own(in) == from) {
      const int64_t li = v.tl->local(in);
      if (!mark[(size_t)li]) { mark[(size_t)li] = 1; ++n; }
    }
    if (v.owner[l] == from && v.own(out) == to && !mark[(size_t)l]) { mark[(size_t)l] = 1; ++n; }
  });
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

// ---- sparse per-item arrays (tiled networks) --------------------------------------------------------
// CUDA virtual memory management, bound at run time: the library keeps no link-time dependency on libcuda.
struct VmmApi {
  CUresult (*AddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*Create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*GetGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  bool ok = false;
};

VmmApi* vmm_api() {
  static VmmApi api;
  static bool tried = false;
  if (tried) return api.ok ? &api : nullptr;
  tried = true;
  auto get = [](const char* name, void** fn) {
    cudaDriverEntryPointQueryResult q;
    return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
  };
  api.ok = get("cuMemAddressReserve", (void**)&api.AddressReserve) && get("cuMemAddressFree", (void**)&api.AddressFree) &&
           get("cuMemCreate", (void**)&api.Create) && get("cuMemRelease", (void**)&api.Release) &&
           get("cuMemMap", (void**)&api.Map) && get("cuMemUnmap", (void**)&api.Unmap) &&
           get("cuMemSetAccess", (void**)&api.SetAccess) && get("cuMemGetAllocationGranularity", (void**)&api.GetGranularity);
  cudaGetLastError();
  return api.ok ? &api : nullptr;
}

void sparse_free(SparseArray& a) {
  if (!a.base) return;
  if (!a.vmm) {
    cudaFree(a.base);
  } else if (VmmApi* v = vmm_api()) {
    for (size_t k = 0; k < a.spans.size(); ++k) {
      v->Unmap((CUdeviceptr)a.base + a.spans[k].first, a.spans[k].second);
      v->Release((CUmemGenericAllocationHandle)a.handles[k]);
    }
    v->AddressFree((CUdeviceptr)a.base, a.va_bytes);
  }
  a = SparseArray();
}

// Allocate a zero-filled array of `elem`-byte entries indexed by global item id, backed where the context holds items.
int sparse_alloc(Ctx* c, size_t elem, void** out) {
  const Tiling& tl = c->tl;
  SparseArray a;
  if (!tl.tiled()) {
    const size_t bytes = std::max<size_t>((size_t)tl.n_global * elem, 16);
    CUDA_TRY(c, cudaMalloc(&a.base, bytes));
    CUDA_TRY(c, cudaMemset(a.base, 0, bytes));
    c->net_sparse.push_back(a);
    *out = a.base;
    return GRIDDY_OK;
  }
  VmmApi* v = vmm_api();
  if (!v) return fail(c, GRIDDY_ERR_CUDA, "a tiled network needs CUDA virtual memory management (cuMemCreate / cuMemMap), which this driver does not offer");
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = c->device;
  size_t gran = 0;
  if (v->GetGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0)
    return fail(c, GRIDDY_ERR_CUDA, "cuMemGetAllocationGranularity failed");
  auto up = [&](size_t x) { return (x + gran - 1) / gran * gran; };
  a.vmm = true;
  a.va_bytes = up(std::max<size_t>((size_t)tl.n_global * elem, 1));
  CUdeviceptr base = 0;
  if (v->AddressReserve(&base, a.va_bytes, gran, 0, 0) != CUDA_SUCCESS)
    return fail(c, GRIDDY_ERR_OOM, "cuMemAddressReserve failed for " + std::to_string(a.va_bytes) + " bytes of address space");
  a.base = (void*)base;
  for (size_t k = 0; k < tl.beg.size(); ++k) {   // granule-aligned spans behind the ranges, merged where they touch
    if (tl.end[k] == tl.beg[k]) continue;
    const size_t lo = (size_t)tl.beg[k] * elem / gran * gran, hi = up((size_t)tl.end[k] * elem);
    if (!a.spans.empty() && lo <= a.spans.back().first + a.spans.back().second)
      a.spans.back().second = std::max(a.spans.back().second, hi - a.spans.back().first);
    else
      a.spans.emplace_back(lo, hi - lo);
  }
  CUmemAccessDesc acc = {};
  acc.location = prop.location;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  for (size_t k = 0; k < a.spans.size(); ++k) {
    CUmemGenericAllocationHandle h = 0;
    const CUresult r1 = v->Create(&h, a.spans[k].second, &prop, 0);
    if (r1 != CUDA_SUCCESS) {
      a.spans.resize(k);
      sparse_free(a);
      return fail(c, GRIDDY_ERR_OOM, "cuMemCreate failed for " + std::to_string(a.spans[k].second) + " bytes");
    }
    a.handles.push_back((unsigned long long)h);
    if (v->Map(base + a.spans[k].first, a.spans[k].second, 0, h, 0) != CUDA_SUCCESS ||
        v->SetAccess(base + a.spans[k].first, a.spans[k].second, &acc, 1) != CUDA_SUCCESS) {
      v->Release(h);
      a.handles.pop_back();
      a.spans.resize(k);
      sparse_free(a);
      return fail(c, GRIDDY_ERR_CUDA, "cuMemMap / cuMemSetAccess failed");
    }
    if (cudaMemset((uint8_t*)a.base + a.spans[k].first, 0, a.spans[k].second) != cudaSuccess) {
      sparse_free(a);
      return fail(c, GRIDDY_ERR_CUDA, "clearing a mapped span failed");
    }
  }
  c->net_sparse.push_back(a);
  *out = a.base;
  return GRIDDY_OK;
}

void free_sparse(Ctx* c) {
  for (SparseArray& a : c->net_sparse) sparse_free(a);
  c->net_sparse.clear();
}

// Compact host array (held items in range order) -> sparse device array, range by range.
template <typename T>
int sparse_upload(Ctx* c, T* dst, const T* host, int per_item = 1) {
  for (size_t k = 0; k < c->tl.beg.size(); ++k) {
    const int64_t n = c->tl.end[k] - c->tl.beg[k];
    if (n > 0)
      CUDA_TRY(c, cudaMemcpy(dst + c->tl.beg[k] * per_item, host + c->tl.off[k] * per_item, (size_t)n * per_item * sizeof(T), cudaMemcpyHostToDevice));
  }
  return GRIDDY_OK;
}

// Sparse device array -> compact host array.
template <typename T>
int sparse_download(Ctx* c, T* host, const T* src, cudaStream_t s) {
  for (size_t k = 0; k < c->tl.beg.size(); ++k) {
    const int64_t n = c->tl.end[k] - c->tl.beg[k];
    if (n > 0) CUDA_TRY(c, cudaMemcpyAsync(host + c->tl.off[k], src + c->tl.beg[k], (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, s));
  }
  return GRIDDY_OK;
}

// Launch one of the tick's kernels behind its predecessor on the stream with programmatic dependent
// launch (records.cuh, grid_dependency_wait).
template <typename... Args>
cudaError_t launch_dependent(void (*kernel)(Args...), unsigned blocks, unsigned threads, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(threads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = GRIDDY_PDL;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

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
      cudaEventCreateWithFlags(&c->ev_scal, cudaEventDisableTiming) != cudaSuccess ||
      cudaHostAlloc((void**)&c->h_scal, SC_COUNT * sizeof(int32_t), cudaHostAllocMapped) != cudaSuccess ||
      rsort::prepare() != cudaSuccess) {
    delete c;
    return GRIDDY_ERR_CUDA;
  }
  c->p.dt = 1.0 / 15.0;
  c->p.two_size = 10.0;
  if (const char* e = getenv("GRIDDY_REORDER_PERIOD")) c->reorder_period = atoi(e);
  if (const char* e = getenv("GRIDDY_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e));   // experiments: 32 / 64 / 128
  *ctx = c;
  return GRIDDY_OK;
}

void griddy_destroy(void* ctx) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return;
  cudaSetDevice(c->device);
  mg_reset(c);
  free_pool(c->net_allocs);
  free_pool(c->actor_allocs);
  free_sparse(c);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->ev_scal) cudaEventDestroy(c->ev_scal);
  if (c->h_scal) cudaFreeHost(c->h_scal);
  for (int f = 0; f < 2; ++f) {
    if (c->ev_drawn[f]) cudaEventDestroy(c->ev_drawn[f]);
    if (c->ev_copied[f]) cudaEventDestroy(c->ev_copied[f]);
    if (c->ev_delta[f]) cudaEventDestroy(c->ev_delta[f]);
    if (c->ev_fmt0[f]) cudaEventDestroy(c->ev_fmt0[f]);
    if (c->ev_copy0[f]) cudaEventDestroy(c->ev_copy0[f]);
    if (c->ev_dcopied[f]) cudaEventDestroy(c->ev_dcopied[f]);
  }
  if (c->h_dcnt) cudaFreeHost(c->h_dcnt);
  if (c->draw_stream) cudaStreamDestroy(c->draw_stream);
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

int griddy_set_item_ranges(void* ctx, int64_t n_items_global, int32_t n_ranges, const int64_t* range_beg, const int64_t* range_end) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (n_ranges == 0) {   // back to an untiled network
    c->have_ranges = false;
    return GRIDDY_OK;
  }
  if (n_items_global < 0 || n_items_global > INT32_MAX - 1 || n_ranges < 0 || n_ranges > MAX_RANGES || !range_beg || !range_end)
    return fail(c, GRIDDY_ERR_BAD_ARG, "set_item_ranges: bad arguments (at most " + std::to_string(MAX_RANGES) + " ranges)");
  Tiling t;
  t.n_global = n_items_global;
  int64_t prev = 0;
  for (int k = 0; k < n_ranges; ++k) {
    if (range_beg[k] < prev || range_end[k] < range_beg[k] || range_end[k] > n_items_global)
      return fail(c, GRIDDY_ERR_BAD_ARG, "set_item_ranges: ranges must ascend, be disjoint and lie inside [0, n_items_global)");
    t.beg.push_back(range_beg[k]);
    t.end.push_back(range_end[k]);
    t.off.push_back(t.n_local);
    t.n_local += range_end[k] - range_beg[k];
    prev = range_end[k];
  }
  c->tl_next = t;
  c->have_ranges = true;
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
  Tiling tl;
  if (c->have_ranges) {
    tl = c->tl_next;
    if (tl.n_local != n_items)
      return fail(c, GRIDDY_ERR_BAD_ARG, "upload_network: n_items must equal the " + std::to_string(tl.n_local) + " items of the ranges set with griddy_set_item_ranges");
  } else {
    tl.whole(n_items);
  }
  // ids inside the arrays are global; a referenced item that this context does not hold cannot be checked
  auto ok_ref = [&](int32_t v, int want_kind) {
    if (v < 0 || v >= tl.n_global) return false;
    const int64_t l = tl.local(v);
    return l < 0 || kind[l] == want_kind;
  };
  for (int64_t r = 0; r < n_items; ++r) {
    const bool lane = kind[r] == GRIDDY_SEGMENT_LANE || kind[r] == GRIDDY_JUNCTION_LANE;
    if (lane && !(lane_len[r] > 0.0)) return fail(c, GRIDDY_ERR_BAD_ARG, "lane (entry " + std::to_string(r) + ") has no positive length");
    if (kind[r] == GRIDDY_SEGMENT_LANE && !ok_ref(lane_segment[r], GRIDDY_SEGMENT))
      return fail(c, GRIDDY_ERR_BAD_ARG, "segment lane (entry " + std::to_string(r) + ") has no segment");
    if (kind[r] == GRIDDY_JUNCTION_LANE &&
        (!ok_ref(lane_out[r], GRIDDY_SEGMENT_LANE) || !ok_ref(lane_junction[r], GRIDDY_JUNCTION)))
      return fail(c, GRIDDY_ERR_BAD_ARG, "junction lane (entry " + std::to_string(r) + ") is not linked");
    if (kind[r] == GRIDDY_SEGMENT)
      for (int32_t o : {seg_outer_forw[r], seg_outer_back[r]})
        if (o != -1 && !ok_ref(o, GRIDDY_SEGMENT_LANE))
          return fail(c, GRIDDY_ERR_BAD_ARG, "segment (entry " + std::to_string(r) + ") has a bad outer lane");
  }
  CUDA_TRY(c, cudaSetDevice(c->device));
  free_pool(c->net_allocs);
  free_pool(c->actor_allocs);
  free_sparse(c);
  c->have_net = c->have_actors = false;
  c->tl = tl;
  Params& p = c->p;
  p.n_items = tl.n_global;
  p.grid_n = 0;
  c->d_geom = nullptr;
  TRY(sparse_alloc(c, sizeof(LaneRec), (void**)&p.lane));
  TRY(sparse_alloc(c, sizeof(int32_t), (void**)&p.occ));
  {   // locality keys: the registration index until griddy_set_locality_keys
    uint32_t* d_key = nullptr;
    TRY(sparse_alloc(c, sizeof(uint32_t), (void**)&d_key));
    std::vector<uint32_t> ident((size_t)n_items);
    for (size_t k = 0; k < tl.beg.size(); ++k)
      for (int64_t g = tl.beg[k]; g < tl.end[k]; ++g) ident[(size_t)(tl.off[k] + (g - tl.beg[k]))] = (uint32_t)g;
    TRY(sparse_upload(c, d_key, ident.data()));
    p.loc_key = d_key;
  }
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
      for (size_t k = 0; k < tl.beg.size(); ++k) {
        const int64_t m = tl.end[k] - tl.beg[k], o = tl.off[k];
        if (m > 0)
          k_item_build<<<(unsigned)((m + 255) / 256), 256>>>(p.lane + tl.beg[k], m, d_kind + o, d_len + o, d_dir + o, d_seg + o,
                                                              d_out + o, d_jun + o, d_of + o, d_ob + o);
      }
      if (cudaDeviceSynchronize() != cudaSuccess) rc = fail(c, GRIDDY_ERR_CUDA, "k_item_build failed");
    }
    free_pool(tmp);
    if (rc) return rc;
  }
  c->loc_key_bits = bits_for((uint32_t)std::max<int64_t>(tl.n_global, 1));
  c->h_kind.assign(kind, kind + n_items);
  c->h_dir.assign(lane_dir, lane_dir + n_items);
  c->h_lane_junction.assign(lane_junction, lane_junction + n_items);
  c->h_lane_out.assign(lane_out, lane_out + n_items);
  c->h_of.assign(seg_outer_forw, seg_outer_forw + n_items);
  c->h_ob.assign(seg_outer_back, seg_outer_back + n_items);
  c->h_from.clear();
  c->h_to.clear();
  c->have_dst = false;
  c->g_nbr_off = nullptr;
  c->g_edges = 0;
  mg_reset(c);
  c->have_net = true;
  return GRIDDY_OK;
}

int griddy_set_locality_keys(void* ctx, const uint32_t* item_key) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_net || !item_key) return fail(c, GRIDDY_ERR_BAD_ARG, "set_locality_keys: no network or null keys");
  CUDA_TRY(c, cudaSetDevice(c->device));
  const Tiling& tl = c->tl;
  uint32_t mx = 0;
  for (int64_t r = 0; r < tl.n_local; ++r) mx = std::max(mx, item_key[r]);
  if (mx == UINT32_MAX) return fail(c, GRIDDY_ERR_BAD_ARG, "locality keys must be below 2^32-1");
  TRY(sparse_upload(c, (uint32_t*)c->p.loc_key, item_key));
  c->loc_key_bits = bits_for(mx);
  return GRIDDY_OK;
}

int griddy_set_routing_grid(void* ctx, int32_t n, const int32_t* lane_from_j, const int32_t* lane_to_j,
                            const int32_t* turn4) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_net) return fail(c, GRIDDY_ERR_BAD_ARG, "set_routing_grid before upload_network");
  if (c->mg.on) return fail(c, GRIDDY_ERR_BAD_ARG, "set_routing_grid after griddy_mg_configure (the partition marks the lanes that reach another rank's junction)");
  const Tiling& tl = c->tl;
  if (n < 1 || (int64_t)n * n > tl.n_global || !lane_from_j || !lane_to_j || !turn4)
    return fail(c, GRIDDY_ERR_BAD_ARG, "set_routing_grid: bad arguments");
  // Every turn of a segment lane must be a junction lane of the junction the lane reaches: the device
  // takes a route head's gating junction from lane_to_j instead of reading the junction lane's record.
  for (int64_t r = 0; r < tl.n_local; ++r) {
    if (c->h_kind[r] != GRIDDY_SEGMENT_LANE) continue;
    for (int h = 0; h < 4; ++h) {
      const int32_t t = turn4[4 * r + h];
      if (t < 0) continue;
      const int64_t lt = t < tl.n_global ? tl.local(t) : -1;
      if (t >= tl.n_global || (lt >= 0 && (c->h_kind[lt] != GRIDDY_JUNCTION_LANE || c->h_lane_junction[lt] != lane_to_j[r])))
        return fail(c, GRIDDY_ERR_BAD_ARG, "set_routing_grid: turn " + std::to_string(h) + " of the lane at entry " + std::to_string(r) +
                                               " is not a junction lane of the junction the lane reaches");
    }
  }
  CUDA_TRY(c, cudaSetDevice(c->device));
  Params& p = c->p;
  {
    std::vector<void*> tmp;
    const int32_t *d_to, *d_turn;
    int rc = dev_upload(c, tmp, &d_to, lane_to_j, tl.n_local);
    if (!rc) rc = dev_upload(c, tmp, &d_turn, turn4, 4 * tl.n_local);
    if (!rc) {
      for (size_t k = 0; k < tl.beg.size(); ++k) {
        const int64_t m = tl.end[k] - tl.beg[k];
        if (m > 0) k_item_set_grid<<<(unsigned)((m + 255) / 256), 256>>>(p.lane + tl.beg[k], m, d_to + tl.off[k], d_turn + 4 * tl.off[k]);
      }
      if (cudaDeviceSynchronize() != cudaSuccess) rc = fail(c, GRIDDY_ERR_CUDA, "k_item_set_grid failed");
    }
    free_pool(tmp);
    if (rc) return rc;
  }
  p.grid_n = n;
  c->h_from.assign(lane_from_j, lane_from_j + tl.n_local);
  c->h_to.assign(lane_to_j, lane_to_j + tl.n_local);
  return GRIDDY_OK;
}

int griddy_set_agenda_destinations(void* ctx, int64_t n_agenda_items, const int32_t* dest_lane, const uint8_t* dest_dir,
                                   const int32_t* dest_from_j, const int32_t* dest_to_j) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (n_agenda_items < 0 || (n_agenda_items > 0 && (!dest_lane || !dest_dir)))
    return fail(c, GRIDDY_ERR_BAD_ARG, "set_agenda_destinations: null array or bad count");
  c->dst_lane.assign(dest_lane, dest_lane + n_agenda_items);
  c->dst_dir.assign(dest_dir, dest_dir + n_agenda_items);
  if (dest_from_j && dest_to_j) {
    c->dst_from.assign(dest_from_j, dest_from_j + n_agenda_items);
    c->dst_to.assign(dest_to_j, dest_to_j + n_agenda_items);
  } else {
    c->dst_from.assign((size_t)n_agenda_items, -1);
    c->dst_to.assign((size_t)n_agenda_items, -1);
  }
  c->have_dst = true;
  return GRIDDY_OK;
}

int griddy_set_route_graph(void* ctx, const int64_t* nbr_off, const int32_t* nbr, const double* cost_out, const double* midpoint_xy) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_net) return fail(c, GRIDDY_ERR_BAD_ARG, "set_route_graph before upload_network");
  if (c->tl.tiled()) return fail(c, GRIDDY_ERR_BAD_ARG, "set_route_graph: a route search needs the whole network, this context holds a tile");
  if (!nbr_off || !cost_out || !midpoint_xy) return fail(c, GRIDDY_ERR_BAD_ARG, "set_route_graph: null array");
  const int64_t n = c->p.n_items;
  if (nbr_off[0] != 0) return fail(c, GRIDDY_ERR_BAD_ARG, "set_route_graph: nbr_off must start at 0");
  for (int64_t r = 0; r < n; ++r) {
    if (nbr_off[r + 1] < nbr_off[r]) return fail(c, GRIDDY_ERR_BAD_ARG, "set_route_graph: nbr_off not monotone");
    const bool lane = c->h_kind[r] == GRIDDY_SEGMENT_LANE || c->h_kind[r] == GRIDDY_JUNCTION_LANE;
    if (!lane && nbr_off[r + 1] != nbr_off[r]) return fail(c, GRIDDY_ERR_BAD_ARG, "set_route_graph: only lanes have neighbours");
    if (lane && !(cost_out[r] >= 0.0)) return fail(c, GRIDDY_ERR_BAD_ARG, "set_route_graph: negative cost");
  }
  const int64_t e = nbr_off[n];
  if (e > 0 && !nbr) return fail(c, GRIDDY_ERR_BAD_ARG, "set_route_graph: null neighbour list");
  for (int64_t k = 0; k < e; ++k) {
    const int32_t v = nbr[k];
    if (v < 0 || v >= n || (c->h_kind[v] != GRIDDY_SEGMENT_LANE && c->h_kind[v] != GRIDDY_JUNCTION_LANE))
      return fail(c, GRIDDY_ERR_BAD_ARG, "set_route_graph: neighbour " + std::to_string(k) + " is not a lane");
  }
  CUDA_TRY(c, cudaSetDevice(c->device));
  TRY(dev_upload(c, c->net_allocs, &c->g_nbr_off, nbr_off, n + 1));
  TRY(dev_upload(c, c->net_allocs, &c->g_nbr, nbr, e));
  TRY(dev_upload(c, c->net_allocs, &c->g_cost, cost_out, n));
  TRY(dev_upload(c, c->net_allocs, &c->g_mid, (const double2*)midpoint_xy, n));
  c->g_edges = e;
  return GRIDDY_OK;
}

namespace {

// find-route for every travel-to item, on the device (routes.cuh): item k's route starts on the lane the actor is
// beside when it gets there — its first location, then the destination of the travel-to before — and ends on the
// item's destination lane.  Fills the CSR upload_actors would have been given.
int find_routes_on_device(Ctx* c, int64_t n, const uint8_t* loc_kind, const int32_t* container, const uint8_t* side,
                          const int64_t* agenda_off, const std::vector<AgendaItem>& agenda, std::vector<int64_t>& route_off,
                          std::vector<int32_t>& route_lane) {
  const int64_t n_agenda = (int64_t)agenda.size(), V = c->p.n_items;
  std::vector<int32_t> start, goal;
  std::vector<int64_t> item_of;
  for (int64_t a = 0; a < n; ++a) {
    int32_t cur = -1;
    if (!loc_kind[a]) cur = side[a] ? c->h_ob[container[a]] : c->h_of[container[a]];   // off-road->on-road, location.scm:49-57
    for (int64_t k = agenda_off[a]; k < agenda_off[a + 1]; ++k) {
      if (agenda[(size_t)k].kind != GRIDDY_TRAVEL_TO) continue;
      start.push_back(loc_kind[a] ? -1 : cur);   // (begin-route$ on an on-road actor has no applicable method)
      goal.push_back(agenda[(size_t)k].dlane);
      item_of.push_back(k);
      cur = agenda[(size_t)k].dlane;
    }
  }
  route_off.assign((size_t)n_agenda + 1, 0);
  route_lane.clear();
  const int64_t n_search = (int64_t)start.size();
  std::vector<int32_t> len((size_t)n_search, 0);
  std::vector<std::vector<int32_t>> paths((size_t)n_search);
  if (n_search > 0) {
    const int64_t heap_cap = 4 * c->g_edges + 16;
    const size_t per = (size_t)V * (sizeof(double) + 2 * sizeof(int32_t)) + (size_t)heap_cap * sizeof(HeapEnt);
    const int64_t batch = std::max<int64_t>(1, std::min<int64_t>(n_search, (int64_t)((size_t)1 << 30) / (int64_t)std::max<size_t>(per, 1)));
    std::vector<void*> tmp;
    int32_t *d_start, *d_goal, *d_came, *d_path, *d_len;
    double* d_best;
    HeapEnt* d_heap;
    int rc = dev_alloc(c, tmp, &d_start, batch);
    if (!rc) rc = dev_alloc(c, tmp, &d_goal, batch);
    if (!rc) rc = dev_alloc(c, tmp, &d_len, batch);
    if (!rc) rc = dev_alloc(c, tmp, &d_best, batch * V);
    if (!rc) rc = dev_alloc(c, tmp, &d_came, batch * V);
    if (!rc) rc = dev_alloc(c, tmp, &d_path, batch * V);
    if (!rc) rc = dev_alloc(c, tmp, &d_heap, batch * heap_cap);
    std::vector<int32_t> h_path;
    for (int64_t b0 = 0; !rc && b0 < n_search; b0 += batch) {
      const int64_t m = std::min(batch, n_search - b0);
      cudaMemcpyAsync(d_start, start.data() + b0, (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream);
      cudaMemcpyAsync(d_goal, goal.data() + b0, (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream);
      k_find_routes<<<(unsigned)((m + 31) / 32), 32, 0, c->stream>>>((int)m, d_start, d_goal, V, c->g_nbr_off, c->g_nbr, c->g_cost, c->g_mid,
                                                                      d_best, d_came, d_heap, heap_cap, d_path, d_len);
      ++c->launches;
      cudaMemcpyAsync(len.data() + b0, d_len, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
      if (cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = fail(c, GRIDDY_ERR_CUDA, std::string("k_find_routes: ") + cudaGetErrorString(cudaGetLastError())); break; }
      for (int64_t s = 0; s < m && !rc; ++s) {
        const int32_t l = len[(size_t)(b0 + s)];
        if (l == -2) rc = fail(c, GRIDDY_ERR_BAD_STATE, "find-route: the search's queue overflowed");
        if (l > 0) {
          paths[(size_t)(b0 + s)].resize((size_t)l);
          if (cudaMemcpy(paths[(size_t)(b0 + s)].data(), d_path + s * V, (size_t)l * sizeof(int32_t), cudaMemcpyDeviceToHost) != cudaSuccess)
            rc = fail(c, GRIDDY_ERR_CUDA, "find-route: copying a route failed");
        }
      }
    }
    free_pool(tmp);
    if (rc) return rc;
  }
  // CSR over ALL agenda items; a travel-to without a route ('no-path) gets the single step -2, which begin-route$ reports
  std::vector<int64_t> cnt((size_t)n_agenda, 0);
  for (int64_t s = 0; s < n_search; ++s) cnt[(size_t)item_of[(size_t)s]] = len[(size_t)s] < 0 ? 1 : len[(size_t)s];
  for (int64_t k = 0; k < n_agenda; ++k) route_off[(size_t)k + 1] = route_off[(size_t)k] + cnt[(size_t)k];
  route_lane.resize((size_t)route_off[(size_t)n_agenda]);
  for (int64_t s = 0; s < n_search; ++s) {
    const int64_t o = route_off[(size_t)item_of[(size_t)s]];
    if (len[(size_t)s] < 0) route_lane[(size_t)o] = -2;
    else std::copy(paths[(size_t)s].begin(), paths[(size_t)s].end(), route_lane.begin() + o);
  }
  return GRIDDY_OK;
}

}  // namespace

int griddy_download_routes(void* ctx, int64_t* route_off, int32_t* route_lane, int64_t capacity, int64_t* n_steps) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_actors || !n_steps) return fail(c, GRIDDY_ERR_BAD_ARG, "download_routes: no actors or null n_steps");
  if (c->h_route_off.empty()) return fail(c, GRIDDY_ERR_BAD_ARG, "download_routes: this world uses the grid routing table, it has no route lists");
  *n_steps = (int64_t)c->h_route_lane.size();
  if (route_off) std::copy(c->h_route_off.begin(), c->h_route_off.end(), route_off);
  if (route_lane) {
    if (capacity < *n_steps) return fail(c, GRIDDY_ERR_BAD_ARG, "download_routes: route_lane too small, need " + std::to_string(*n_steps));
    std::copy(c->h_route_lane.begin(), c->h_route_lane.end(), route_lane);
  }
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
  const bool table_routes = c->mg.on && c->have_mg_routes && !route_off;   // routes by index into the replicated table
  if (table_routes && (int64_t)c->mg_item_route.size() != n_agenda)
    return fail(c, GRIDDY_ERR_BAD_ARG, "griddy_mg_set_routes was given " + std::to_string(c->mg_item_route.size()) +
                                           " items, the agendas hold " + std::to_string(n_agenda));
  if (!route_off && !table_routes && c->p.grid_n == 0 && !c->g_nbr_off)
    return fail(c, GRIDDY_ERR_BAD_ARG, "no routes uploaded, no route graph (griddy_set_route_graph) and no grid routing table set");
  const Params& q = c->p;
  const Tiling& tl = c->tl;
  for (int64_t a = 0; a < n; ++a) {
    const int32_t ct = container[a];
    if (ct < 0 || ct >= q.n_items || tl.local(ct) < 0)
      return fail(c, GRIDDY_ERR_BAD_ARG, "actor " + std::to_string(a) + ": container out of range (or not held by this context)");
    const uint8_t k = c->kind_of(ct);
    if (loc_kind[a] ? (k != GRIDDY_SEGMENT_LANE && k != GRIDDY_JUNCTION_LANE) : (k != GRIDDY_SEGMENT))
      return fail(c, GRIDDY_ERR_BAD_ARG, "actor " + std::to_string(a) + ": location kind does not match its container");
    if (!(speed[a] > 0.0)) return fail(c, GRIDDY_ERR_BAD_ARG, "actor " + std::to_string(a) + ": max-speed must be positive");
    if (agenda_off[a + 1] < agenda_off[a]) return fail(c, GRIDDY_ERR_BAD_ARG, "agenda_off not monotone");
  }
  if (c->have_dst && (int64_t)c->dst_lane.size() != n_agenda)
    return fail(c, GRIDDY_ERR_BAD_ARG, "griddy_set_agenda_destinations was given " + std::to_string(c->dst_lane.size()) +
                                           " items, the agendas hold " + std::to_string(n_agenda));
  // Agenda items as the device keeps them: a travel-to carries its destination resolved now —
  // off-road->on-road (location.scm:49-57): the outer lane of (segment, side) (location.scm:16-24), the
  // pos-param flipped on a 'back lane, the lane's junctions for the grid routing table.
  std::vector<AgendaItem> agenda((size_t)n_agenda);
  for (int64_t k = 0; k < n_agenda; ++k) {
    AgendaItem& ai = agenda[(size_t)k];
    memset(&ai, 0, sizeof(ai));
    ai.dlane = ai.dfrom = ai.dto = -1;
    ai.kind = item_kind[k];
    ai.rid = (int32_t)k;   // one GPU: the routes are uploaded per item
    if (item_kind[k] == GRIDDY_SLEEP_FOR) {
      ai.sleep = item_sleep[k];
    } else if (item_kind[k] == GRIDDY_TRAVEL_TO) {
      const int32_t sg = item_seg[k];
      if (sg < 0 || sg >= q.n_items) return fail(c, GRIDDY_ERR_BAD_ARG, "agenda item " + std::to_string(k) + ": destination out of range");
      const int64_t ls = tl.local(sg);
      if (ls >= 0 && c->h_kind[ls] != GRIDDY_SEGMENT)
        return fail(c, GRIDDY_ERR_BAD_ARG, "agenda item " + std::to_string(k) + ": destination is not a segment");
      uint8_t dir = 0;
      if (c->have_dst) {
        ai.dlane = c->dst_lane[(size_t)k];
        dir = c->dst_dir[(size_t)k];
        ai.dfrom = c->dst_from[(size_t)k];
        ai.dto = c->dst_to[(size_t)k];
        if (ai.dlane >= q.n_items) return fail(c, GRIDDY_ERR_BAD_ARG, "agenda item " + std::to_string(k) + ": destination lane out of range");
      } else {
        if (ls < 0)
          return fail(c, GRIDDY_ERR_BAD_ARG, "agenda item " + std::to_string(k) + ": its destination segment is not held by this context; "
                                                 "a tiled network needs griddy_set_agenda_destinations");
        ai.dlane = item_side[k] ? c->h_ob[ls] : c->h_of[ls];
        const int64_t ll = ai.dlane >= 0 ? tl.local(ai.dlane) : -1;
        if (ai.dlane >= 0 && ll < 0)
          return fail(c, GRIDDY_ERR_BAD_ARG, "agenda item " + std::to_string(k) + ": its destination lane is not held by this context; "
                                                 "a tiled network needs griddy_set_agenda_destinations");
        if (ll >= 0) {
          dir = c->h_dir[ll];
          if (!c->h_from.empty()) { ai.dfrom = c->h_from[ll]; ai.dto = c->h_to[ll]; }
        }
      }
      ai.dpos = dir ? 1.0 - item_pos[k] : item_pos[k];
    } else {
      return fail(c, GRIDDY_ERR_BAD_ARG, "agenda item " + std::to_string(k) + ": unknown kind");
    }
  }
  std::vector<int64_t> found_off;
  std::vector<int32_t> found_lane;
  if (!route_off && c->g_nbr_off && c->p.grid_n == 0) {   // find-route on the device
    CUDA_TRY(c, cudaSetDevice(c->device));
    TRY(find_routes_on_device(c, n, loc_kind, container, side, agenda_off, agenda, found_off, found_lane));
    route_off = found_off.data();
    route_lane = found_lane.data();
  }
  c->h_route_off.clear();
  c->h_route_lane.clear();
  if (route_off) {
    c->h_route_off.assign(route_off, route_off + n_agenda + 1);
    c->h_route_lane.assign(route_lane, route_lane + route_off[n_agenda]);
  }
  int64_t n_routes = n_agenda;   // entries of the route table the device gets
  if (table_routes) {
    route_off = c->mg_route_off.data();
    route_lane = c->mg_route_lane.data();
    n_routes = (int64_t)c->mg_route_off.size() - 1;
    for (int64_t k = 0; k < n_agenda; ++k) {
      if (item_kind[k] != GRIDDY_TRAVEL_TO) continue;
      const int32_t rid = c->mg_item_route[(size_t)k];
      if (rid < 0 || rid >= n_routes) return fail(c, GRIDDY_ERR_BAD_ARG, "agenda item " + std::to_string(k) + ": route index out of range");
      agenda[(size_t)k].rid = rid;
    }
  }
  if (route_off) {
    if (route_off[n_routes] > 0 && !route_lane) return fail(c, GRIDDY_ERR_BAD_ARG, "route_lane is null");
    if (route_off[n_routes] > INT32_MAX - 1) return fail(c, GRIDDY_ERR_BAD_ARG, "too many route steps");
    for (int64_t k = 0; k < route_off[n_routes]; ++k) {
      const int32_t l = route_lane[k];
      if (l == -2) continue;   // 'no-path, reported when the actor gets there
      const uint8_t kd = (l >= 0 && l < q.n_items) ? c->kind_of(l) : (uint8_t)254;
      if (kd != 255 && kd != GRIDDY_SEGMENT_LANE && kd != GRIDDY_JUNCTION_LANE)
        return fail(c, GRIDDY_ERR_BAD_ARG, "route step " + std::to_string(k) + " is not a lane");
    }
  }

  CUDA_TRY(c, cudaSetDevice(c->device));
  free_pool(c->actor_allocs);
  c->have_actors = false;
  c->d_xy32[0] = c->d_xy32[1] = nullptr;
  c->d_xy_sync = nullptr;
  c->frames_begun = c->frames_ended = 0;
  c->d_drawn = nullptr;
  c->deltas_begun = c->deltas_ended = c->deltas_flushed = 0;
  c->delta_reorders_seen = -1;
  c->delta_cover = 0;
  Params& p = c->p;
  p.n_actors = n;
  const int64_t extra = c->mg.on ? c->mg.extra_slots : 0;   // room for actors arriving from other ranks
  if (c->mg.on && route_off && !table_routes)
    return fail(c, GRIDDY_ERR_BAD_ARG, "a partitioned world takes explicit routes through griddy_mg_set_routes (every rank holds the whole table), not per item");
  if (table_routes && tl.tiled()) return fail(c, GRIDDY_ERR_BAD_ARG, "explicit routes on a partitioned world need the whole network on every rank (no griddy_set_item_ranges)");
  const int64_t extra_items = extra * std::max(c->mg.max_agenda_items, 1);
  if (n + extra > INT32_MAX - 2 * rsort::TILE || n_agenda + extra_items > INT32_MAX - 1)
    return fail(c, GRIDDY_ERR_BAD_ARG, "too many slots");
  const int64_t cap = (n + extra + rsort::TILE - 1) / rsort::TILE * rsort::TILE + rsort::TILE;   // tile-padded
  p.cap = (int32_t)cap;
  p.icap = (int32_t)(n_agenda + extra_items);
  auto& pool = c->actor_allocs;

  // Initial world: link! every actor in id order (core.scm:169-178).  Slot = id until the first reorder.
  std::vector<Hot> hot(n);
  std::vector<Cold> cold(n);
  std::vector<ColdF> coldf(n);
  std::vector<int32_t> occ((size_t)tl.n_local, 0);   // compact, like every host copy of a per-item array
  std::vector<int32_t> tail_lane, tail_slot;   // the few lanes that start with actors on them, and remote-lane marks
  std::vector<int32_t> on_road;
  for (int64_t a = 0; a < n; ++a) {
    uint32_t s = ST_ALIVE;
    Hot& h = hot[a];
    memset(&h, 0, sizeof(h));
    h.ahead = h.behind = -1;
    h.tj = -1;
    Cold& k = cold[a];
    memset(&k, 0, sizeof(k));
    k.container = container[a];
    k.agenda_cur = (int32_t)agenda_off[a];
    k.gid = (int32_t)a;         // global id = upload index until griddy_mg_set_global_ids
    k.hop = HOP_ARRIVE;
    k.dest_lane = k.dest_from_j = k.dest_to_j = -1;
    k.inbox_next = k.outbox_next = k.ghost_next = -1;
    k.speed = speed[a];
    k.speed_dt = speed_dt[a];
    ColdF& f = coldf[a];
    memset(&f, 0, sizeof(f));
    f.agenda_beg = (int32_t)agenda_off[a];
    f.agenda_end = (int32_t)agenda_off[a + 1];
    f.land_lane = (int32_t)a;   // landing stamp of the initial placement: tick 0, link! order
    if (agenda_off[a] < agenda_off[a + 1]) {
      if (item_kind[agenda_off[a]] == GRIDDY_SLEEP_FOR) { s |= HEAD_SLEEP << ST_HEAD_SHIFT; h.tj = item_sleep[agenda_off[a]]; }
      else s |= HEAD_TRAVEL << ST_HEAD_SHIFT;
    }
    if (loc_kind[a]) { s |= ST_ONROAD; on_road.push_back((int32_t)a); }
    else if (side[a]) s |= ST_SIDE;
    h.st = s;
  }
  // lanes: ascending pos-param; an equal key replaces the earlier actor (bbtree-set)
  std::sort(on_road.begin(), on_road.end(), [&](int32_t x, int32_t y) {
    if (container[x] != container[y]) return container[x] < container[y];
    if (pos[x] != pos[y]) return pos[x] < pos[y];
    return x < y;
  });
  int32_t prev = -1, dead0 = 0;
  for (size_t k = 0; k < on_road.size(); ++k) {
    const int32_t a = on_road[k];
    const bool replaced = k + 1 < on_road.size() && container[on_road[k + 1]] == container[a] && pos[on_road[k + 1]] == pos[a];
    if (replaced) { hot[a].st &= ~ST_ALIVE; ++dead0; continue; }
    const int32_t lane = container[a];
    if (prev >= 0 && container[prev] == lane) { hot[prev].ahead = a; hot[a].behind = prev; } else { tail_lane.push_back(lane); tail_slot.push_back(a); }
    const int64_t ll = tl.local(lane);
    if (c->h_kind[ll] == GRIDDY_JUNCTION_LANE) {
      const int64_t lj = tl.local(c->h_lane_junction[ll]);
      if (lj < 0) return fail(c, GRIDDY_ERR_BAD_ARG, "an actor stands on a junction lane whose junction this context does not hold");
      ++occ[(size_t)lj];
    }
    prev = a;
  }

  // permuted arrays: two buffers each (the tick's two pos-param buffers and a third one rotate)
  TRY(dev_alloc(c, pool, &c->hot_cur, cap));
  TRY(dev_alloc(c, pool, &c->hot_alt, cap));
  CUDA_TRY(c, cudaMemset(c->hot_cur, 0, (size_t)cap * sizeof(Hot)));
  TRY(dev_alloc(c, pool, &c->cold_cur, cap));
  TRY(dev_alloc(c, pool, &c->cold_alt, cap));
  TRY(dev_alloc(c, pool, &c->coldf_cur, cap));
  TRY(dev_alloc(c, pool, &c->coldf_alt, cap));
  CUDA_TRY(c, cudaMemset(c->cold_cur, 0, (size_t)cap * sizeof(Cold)));
  CUDA_TRY(c, cudaMemset(c->coldf_cur, 0, (size_t)cap * sizeof(ColdF)));
  for (int q = 0; q < 2; ++q) {
    TRY(dev_alloc(c, pool, &p.pos[q], cap));
    CUDA_TRY(c, cudaMemset(p.pos[q], 0, (size_t)cap * sizeof(double)));
  }
  TRY(dev_alloc(c, pool, &c->pos_alt, cap));
  if (n) {
    CUDA_TRY(c, cudaMemcpy(c->hot_cur, hot.data(), (size_t)n * sizeof(Hot), cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(p.pos[0], pos, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(p.pos[1], pos, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->cold_cur, cold.data(), (size_t)n * sizeof(Cold), cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->coldf_cur, coldf.data(), (size_t)n * sizeof(ColdF), cudaMemcpyHostToDevice));
  }
  p.hot = c->hot_cur;
  p.cold = c->cold_cur;
  p.coldf = c->coldf_cur;

  TRY(dev_alloc(c, pool, &p.touched, cap));
  TRY(dev_alloc(c, pool, &p.leavers, cap));
  TRY(dev_alloc(c, pool, &p.slow, cap));
  TRY(dev_alloc(c, pool, &p.counters, 2 * CNT_STRIDE));
  TRY(dev_alloc(c, pool, &p.scal, SC_COUNT));
#ifdef GRIDDY_TRACE
  TRY(dev_alloc(c, pool, &p.trace, 64 * TR_COUNT));
  CUDA_TRY(c, cudaMemset(p.trace, 0, 64 * TR_COUNT * sizeof(unsigned long long)));
#else
  p.trace = nullptr;
#endif
  CUDA_TRY(c, cudaMemset(p.counters, 0, 2 * CNT_STRIDE * sizeof(int32_t)));
  {
    int32_t scal[SC_COUNT] = {0, (int32_t)n, 0, (int32_t)n_agenda, dead0, 0, 0, 0};
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

  // agenda items; room at the end for the items of actors that arrive from other ranks
  TRY(dev_alloc(c, pool, &p.agenda, p.icap));
  if (n_agenda > 0) CUDA_TRY(c, cudaMemcpy(p.agenda, agenda.data(), (size_t)n_agenda * sizeof(AgendaItem), cudaMemcpyHostToDevice));
  c->agenda_alt = nullptr;
  c->d_nitems_new = nullptr;
  if (c->mg.on) {   // a reorder compacts the pool into a second one (arrivals keep appending their items)
    TRY(dev_alloc(c, pool, &c->agenda_alt, p.icap));
    TRY(dev_alloc(c, pool, &c->d_nitems_new, 1));
  }
  if (route_off) {
    TRY(dev_upload(c, pool, &p.route_off, route_off, n_routes + 1));
    TRY(dev_upload(c, pool, &p.route_lane, route_lane, route_off[n_routes]));
  } else {
    p.route_off = nullptr;
    p.route_lane = nullptr;
  }
  // per-lane state of the network belongs to this world
  for (const auto& m : c->mg.import_marks) { tail_lane.push_back(m.first); tail_slot.push_back(m.second); }   // lanes other ranks own
  if (c->mg.on)
    for (int64_t a = 0; a < n; ++a)
      if (c->owner_of(container[a]) != c->mg.rank)
        return fail(c, GRIDDY_ERR_BAD_ARG, "actor " + std::to_string(a) + " is not on this rank's part of the world");
  for (size_t k = 0; k < tl.beg.size(); ++k) {
    const int64_t m = tl.end[k] - tl.beg[k];
    if (m > 0) k_lane_init<<<(unsigned)((m + 255) / 256), 256>>>(p.lane + tl.beg[k], m);
  }
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
  TRY(sparse_upload(c, p.occ, occ.data()));
  // a partitioned world starts over: no records, no flags (they count ticks) in my window
  if (c->mg.on) {
    CUDA_TRY(c, cudaMemset(c->mg.window, 0, c->mg.window_bytes));
    CUDA_TRY(c, cudaMemset(c->p.mig_done, 0, sizeof(int32_t)));
    CUDA_TRY(c, cudaMemset(c->p.halo_done, 0, sizeof(int32_t)));
    CUDA_TRY(c, cudaMemset(c->p.mig_count, 0, 2 * MAX_RANKS * sizeof(int32_t)));
  }
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

// The device's scalars as last read: slot count, dead slots, error bits.
int take_scalars(Ctx* c, const int32_t* scal) {
  c->ns_host = std::min(scal[SC_NSLOTS], c->p.cap);
  c->dead_host = scal[SC_NDEAD];
  const int32_t derr = scal[SC_ERR];
  if (!derr) return GRIDDY_OK;
  std::string m = "advance$:";
  if (derr & DEV_ERR_BEGIN_ON_ROAD) m += " begin-route$ on an on-road actor (no applicable method);";
  if (derr & DEV_ERR_NO_CLAUSE) m += " no matching clause for (agenda, route);";
  if (derr & DEV_ERR_END_ON_JLANE) m += " end-route$ on a junction lane;";
  if (derr & DEV_ERR_NO_LANE) m += " road-has-no-lanes;";
  if (derr & DEV_ERR_NO_ROUTE) m += " routing table has no next hop;";
  if (derr & DEV_ERR_INTERNAL) m += " internal: a live actor pointed at a dropped slot;";
  if (derr & DEV_ERR_MIG_OVERFLOW) m += " multi-GPU: more boundary crossings (or migrating agenda items) in one tick than the window holds, or no slots / agenda room left for arrivals;";
  if (derr & DEV_ERR_MIG_TIMEOUT) m += " multi-GPU: a neighbour never signalled that its halo or its migrants of a tick were written;";
  return fail(c, GRIDDY_ERR_BAD_STATE, m);
}

// End of a step: device error bits and the slot count, in one copy.
int finish_step(Ctx* c) {
  int32_t scal[SC_COUNT] = {};
  CUDA_TRY(c, cudaMemcpy(scal, c->p.scal, sizeof(scal), cudaMemcpyDeviceToHost));
  c->scal_pending = false;
  return take_scalars(c, scal);
}

// One tick = four launches on the compute stream (plus a reorder every reorder_period ticks; on a
// partitioned world the halo and migrant kernels and the two exchanges run beside them, see launch_tick).
// ev (optional) receives 6 events bracketing reorder, k_advance, k_slow, k_unlink, k_link.
int tick_begin(Ctx* c, int64_t tick, cudaEvent_t* ev) {
  if (ev) cudaEventRecord(ev[0], c->stream);
  // Reorder every reorder_period ticks, and — checked wherever a step begins, which is when the host knows
  // the count — as soon as more than 1/16 of the slots hold actors that are gone: k_advance streams dead slots too.
  if (c->reorder_period > 0 && tick != c->last_reorder_tick &&
      (tick % c->reorder_period == 0 || ((int64_t)c->dead_host * 16 > c->ns_host && tick - c->last_reorder_tick >= 16)))
    TRY(reorder_slots(c, tick));
  if (ev) cudaEventRecord(ev[1], c->stream);
  return GRIDDY_OK;
}

// fused (a partitioned world, one process per GPU): the halo and the immigrants are taken in by the last blocks of
// k_advance and k_unlink; otherwise (all ranks on one GPU, ordered by events) by kernels of their own.
int tick_advance(Ctx* c, int64_t tick, cudaEvent_t* ev, bool fused = false) {
  const unsigned per_block = ADV_THREADS * ADV_UNROLL;
  const unsigned adv_blocks = (unsigned)((c->ns_host + per_block - 1) / per_block);
  // a partitioned world: its first blocks send the halo (at least one: it raises the flags)
  const unsigned halo_blocks = c->mg.nb.empty() ? 0u : (unsigned)std::max(1, (c->mg.n_exp + ADV_THREADS - 1) / ADV_THREADS);
  const unsigned unpack_blocks = (fused && !c->mg.nb.empty()) ? (unsigned)std::max(1, (c->mg.n_imp_junc + ADV_THREADS - 1) / ADV_THREADS) : 0u;
  if (adv_blocks + halo_blocks + unpack_blocks)
    CUDA_TRY(c, launch_dependent(k_advance, adv_blocks + halo_blocks + unpack_blocks, ADV_THREADS, c->stream, c->p, (int)tick,
                                 (int)halo_blocks, (int)unpack_blocks));
  if (ev) cudaEventRecord(ev[2], c->stream);
  ++c->launches;
  return GRIDDY_OK;
}

int tick_slow(Ctx* c, int64_t tick, cudaEvent_t* ev) {   // on a partitioned world: needs the halo
  const unsigned aux_blocks = std::min<unsigned>(std::max((unsigned)((c->ns_host + 255) / 256), 1u), 148u * 8u);
  CUDA_TRY(c, launch_dependent(k_slow, aux_blocks * 2, SLOW_THREADS, c->stream, c->p, (int)tick));
  if (ev) cudaEventRecord(ev[3], c->stream);
  ++c->launches;
  return GRIDDY_OK;
}

// one thread per touched lane where possible: grid-striding doubles the longest chain
unsigned relink_blocks(const Ctx* c) {
  return std::min<unsigned>(std::max((unsigned)((c->ns_host / 8 + 127) / 128), 1u), 148u * 64u);
}

int tick_unlink(Ctx* c, int64_t tick, cudaEvent_t* ev, bool fused = false) {   // local lists only: does not need the migrants
  unsigned immig_blocks = 0;
  if (fused && !c->mg.nb.empty()) {
    int most = 1;
    for (Neighbor& nb : c->mg.nb) {
      most = std::max(most, nb.mig_in);
      // upper bound of the slots in use: the neighbour may have filled its buffer
      c->ns_host = (int32_t)std::min<int64_t>(c->p.cap, (int64_t)c->ns_host + nb.mig_in);
    }
    immig_blocks = (unsigned)std::max(1, most / 512);
  }
  CUDA_TRY(c, launch_dependent(k_unlink, relink_blocks(c) + immig_blocks * (unsigned)c->mg.world, 128u, c->stream, c->p, (int)tick,
                               (int)immig_blocks, (int)c->mg.world));
  if (ev) cudaEventRecord(ev[4], c->stream);
  ++c->launches;
  return GRIDDY_OK;
}

int tick_link(Ctx* c, int64_t tick, cudaEvent_t* ev) {   // on a partitioned world: needs the migrants
  CUDA_TRY(c, launch_dependent(k_link, relink_blocks(c), 128u, c->stream, c->p, (int)tick));
  if (ev) cudaEventRecord(ev[5], c->stream);
  ++c->launches;
  return GRIDDY_OK;
}

// ---- the small kernels of a partitioned world ----------------------------------------------------------
// Per tick, all on the compute stream:
//   k_advance        its first blocks send the halo: they read the OLD world only, like the rest of the kernel, and
//                    store into the neighbours' windows, so the halo travels over NVLink while the actors stream
//   k_halo_unpack    waits for the neighbours' halo flags; k_slow behind it may read the halo
//   k_slow           a mover whose new lane another rank owns writes its record into that rank's window; the
//                    block that finishes last raises the migrant flags
//   k_unlink         local lists only — meanwhile the neighbours' flags arrive
//   k_immig_unpack   waits for the neighbours' migrant flags
//   k_link
int mg_halo_unpack(Ctx* c, int64_t tick, cudaStream_t s) {
  if (c->mg.nb.empty()) return GRIDDY_OK;
  const unsigned blocks = (unsigned)std::max(1, (c->mg.n_imp_junc + 127) / 128);   // at least one: it waits for the flags
  CUDA_TRY(c, launch_dependent(k_halo_unpack, blocks, 128u, s, c->p, (int)tick));
  ++c->launches;
  return GRIDDY_OK;
}

int mg_immig_unpack(Ctx* c, int64_t tick, cudaStream_t s) {
  if (c->mg.nb.empty()) return GRIDDY_OK;
  int most = 1;
  for (Neighbor& nb : c->mg.nb) {
    most = std::max(most, nb.mig_in);
    // upper bound of the slots in use: the neighbour may have filled its buffer
    c->ns_host = (int32_t)std::min<int64_t>(c->p.cap, (int64_t)c->ns_host + nb.mig_in);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)std::max(1, most / 512), (unsigned)c->mg.world);   // blockIdx.y: the rank they came from
  cfg.blockDim = dim3(128);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = GRIDDY_PDL;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CUDA_TRY(c, cudaLaunchKernelEx(&cfg, k_immig_unpack, c->p, (int)tick));
  ++c->launches;
  return GRIDDY_OK;
}

int launch_tick(Ctx* c, int64_t tick, cudaEvent_t* ev) {
  TRY(tick_begin(c, tick, ev));
  if (!c->mg.on) {
    TRY(tick_advance(c, tick, ev));
    TRY(tick_slow(c, tick, ev));
    TRY(tick_unlink(c, tick, ev));
    return tick_link(c, tick, ev);
  }
  c->p.publish_late = 1;                          // (the first block of k_unlink raises the migrant flags)
  TRY(tick_advance(c, tick, ev, true));           // its first blocks send the halo, which travels while the rest run; its last take the neighbours' in
  TRY(tick_slow(c, tick, ev));                    // movers that cross write themselves into the neighbours' windows
  TRY(tick_unlink(c, tick, ev, true));            // its last blocks wait for the neighbours' migrant flags and unpack the arrivals
  return tick_link(c, tick, ev);
}

// ---- local transport: every rank is a context of this process (tests) -------------------------------
// The same kernels and the same windows; what differs is only that the ranks' streams are ordered with
// events (a rank's waiting kernel is enqueued behind the event its neighbour recorded after the writing
// kernel), so that the flags are always up when a waiting kernel starts — several ranks may share one GPU,
// where a spinning kernel could otherwise keep the writer from running.
int local_begin(Ctx* c, int64_t tick, cudaEvent_t* ev) {
  c->p.publish_late = 0;
  TRY(tick_begin(c, tick, ev));
  return tick_advance(c, tick, ev);   // (its first blocks write the halo into the neighbours' windows)
}

int local_slow(Ctx* c, int64_t tick, cudaEvent_t* ev) {
  TRY(mg_halo_unpack(c, tick, c->stream));
  return tick_slow(c, tick, ev);
}

int local_link(Ctx* c, int64_t tick, cudaEvent_t* ev) {
  TRY(mg_immig_unpack(c, tick, c->stream));
  return tick_link(c, tick, ev);
}

int local_unlink(Ctx* c, int64_t tick, cudaEvent_t* ev) { return tick_unlink(c, tick, ev); }

int wait_for_neighbours(Ctx* src) {   // every neighbour's stream waits for what src has enqueued so far
  for (Neighbor& nb : src->mg.nb) {
    Ctx* dst = src->mg.peers[nb.rank];
    CUDA_TRY(dst, cudaSetDevice(dst->device));
    CUDA_TRY(dst, cudaStreamWaitEvent(dst->stream, src->mg.ev_ready, 0));
  }
  return GRIDDY_OK;
}

}  // namespace

namespace {
int delta_flush(Ctx* c, bool wait);   // drawing read-back, below
}

int griddy_step(void* ctx, int64_t n_ticks) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_actors) return fail(c, GRIDDY_ERR_BAD_ARG, "step before upload_actors");
  if (n_ticks < 0 || c->tick + n_ticks > INT32_MAX - 2) return fail(c, GRIDDY_ERR_BAD_ARG, "bad n_ticks");
  if (c->mg.on && (!c->mg.connected || !c->mg.peers.empty()))
    return fail(c, GRIDDY_ERR_BAD_ARG, "a partitioned world steps with griddy_mg_step_local, or — one process per GPU — with griddy_step "
                                       "once every neighbour's window is mapped (griddy_mg_ipc_import)");
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

int griddy_step_async(void* ctx, int64_t n_ticks) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_actors) return fail(c, GRIDDY_ERR_BAD_ARG, "step before upload_actors");
  if (n_ticks < 0 || c->tick + n_ticks > INT32_MAX - 2) return fail(c, GRIDDY_ERR_BAD_ARG, "bad n_ticks");
  if (c->mg.on && (!c->mg.connected || !c->mg.peers.empty()))
    return fail(c, GRIDDY_ERR_BAD_ARG, "a partitioned world steps with griddy_mg_step_local, or once every neighbour's window is mapped");
  CUDA_TRY(c, cudaSetDevice(c->device));
  TRY(delta_flush(c, false));   // a delta frame that is formatted by now starts its way to the host before the next ticks are enqueued
  // the scalars an earlier call asked for, if they have landed: slot count and error bits without a wait
  if (c->scal_pending && cudaEventQuery(c->ev_scal) == cudaSuccess) {
    c->scal_pending = false;
    TRY(take_scalars(c, c->h_scal));
    // the copy is older than the ticks enqueued after it: every one of them may have brought arrivals
    int64_t per_tick = 0;
    for (const Neighbor& nb : c->mg.nb) per_tick += nb.mig_in;
    c->ns_host = (int32_t)std::min<int64_t>(c->p.cap, (int64_t)c->ns_host + per_tick * c->ticks_after_scal);
  }
  if (c->ns_host > 0 || c->mg.on)
    for (int64_t t = 0; t < n_ticks; ++t) TRY(launch_tick(c, c->tick + t, nullptr));
  c->tick += n_ticks;
  c->ticks_after_scal += n_ticks;
  if (!c->scal_pending) {
    // (a kernel's stores into the mapped mirror, not a copy: a device-to-host copy here would queue up behind the
    // frame that the copy engine is carrying for griddy_positions_begin / griddy_delta_begin and stall this stream)
    int32_t* mirror = nullptr;
    CUDA_TRY(c, cudaHostGetDevicePointer((void**)&mirror, c->h_scal, 0));
    k_mirror_scalars<<<1, 32, 0, c->stream>>>(c->p.scal, mirror);
    ++c->launches;
    CUDA_TRY(c, cudaEventRecord(c->ev_scal, c->stream));
    c->scal_pending = true;
    c->ticks_after_scal = 0;
  }
  return GRIDDY_OK;
}

int griddy_reorder_now(void* ctx, double* ms) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_actors) return fail(c, GRIDDY_ERR_BAD_ARG, "reorder before upload_actors");
  CUDA_TRY(c, cudaSetDevice(c->device));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  TRY(finish_step(c));
  CUDA_TRY(c, cudaEventRecord(c->ev0, c->stream));
  TRY(reorder_slots(c, c->tick));
  CUDA_TRY(c, cudaEventRecord(c->ev1, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  float t = 0.f;
  CUDA_TRY(c, cudaEventElapsedTime(&t, c->ev0, c->ev1));
  if (ms) *ms = t;
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
    TRY(each(local_begin, t, true));                       // (k_advance's first blocks write the neighbours' windows)
    for (Ctx* c : cs) TRY(wait_for_neighbours(c));
    TRY(each(local_slow, t, false));
    for (Ctx* c : cs) {                                    // (k_slow wrote the emigrants into the neighbours' windows)
      CUDA_TRY(c, cudaSetDevice(c->device));
      CUDA_TRY(c, cudaEventRecord(c->mg.ev_ready, c->stream));
    }
    TRY(each(local_unlink, t, false));
    for (Ctx* c : cs) TRY(wait_for_neighbours(c));
    TRY(each(local_link, t, false));
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
    // actors by descending pos-param, a segment's in list order (landing stamps, see lanes.cuh).
    // What the comparison needs of every slot, packed on the device into 32 bytes (instead of pulling the hot and
    // both cold records, 160 bytes a slot); the living actors' slots are then sorted here.
    OrderRec* d_rec = nullptr;
    std::vector<void*> tmp;
    TRY(dev_alloc(c, tmp, &d_rec, ns));
    k_pack_order<<<blocks, 256, 0, c->stream>>>(p, tick & 1, c->d_ghost, d_rec);
    ++c->launches;
    std::vector<OrderRec> rec(ns);
    const cudaError_t e = cudaMemcpy(rec.data(), d_rec, (size_t)ns * sizeof(OrderRec), cudaMemcpyDeviceToHost);
    free_pool(tmp);
    CUDA_TRY(c, e);
    std::vector<uint32_t> st(ns);
    std::vector<uint8_t> ghost(ns);
    std::vector<int32_t> cont(ns), land_tick(ns), land_lane(ns), ident(ns);
    std::vector<double> ps(ns), land_pos(ns);
    for (int32_t a = 0; a < ns; ++a) {
      const OrderRec& r = rec[a];
      st[a] = r.st & ~ORDER_GHOST; ghost[a] = (r.st & ORDER_GHOST) ? 1 : 0;
      cont[a] = r.container; ident[a] = r.gid;
      land_tick[a] = r.land_tick; land_lane[a] = r.land_lane;
      // on the road the order is by pos-param, off the road by landing stamp: one slot serves both
      ps[a] = r.key; land_pos[a] = r.key;
    }
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
  if (c->tl.n_local == 0 || (!lane_count && !junction_count)) return GRIDDY_OK;
  // counted on the device: one atomic per living actor; the junction counters are the device's own
  // incrementally maintained ones (route.scm:98-104), less the ghosts (file header).  The scratch arrays
  // are indexed by global id like every per-item array; the outputs are compact (the held items in range order).
  const size_t before = c->net_sparse.size();
  int32_t *d_lane = nullptr, *d_junc = nullptr;
  auto done = [&](int r) {
    while (c->net_sparse.size() > before) { sparse_free(c->net_sparse.back()); c->net_sparse.pop_back(); }
    return r;
  };
  if (lane_count) { const int rc = sparse_alloc(c, sizeof(int32_t), (void**)&d_lane); if (rc) return done(rc); }
  if (junction_count) { const int rc = sparse_alloc(c, sizeof(int32_t), (void**)&d_junc); if (rc) return done(rc); }
  cudaStream_t s = c->stream;
  if (d_junc)
    for (size_t k = 0; k < c->tl.beg.size(); ++k) {
      const int64_t m = c->tl.end[k] - c->tl.beg[k];
      if (m > 0 && cudaMemcpyAsync(d_junc + c->tl.beg[k], p.occ + c->tl.beg[k], (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToDevice, s) != cudaSuccess)
        return done(fail(c, GRIDDY_ERR_CUDA, "copy of the junction counters failed"));
    }
  if (ns > 0) {
    const unsigned blocks = (unsigned)((ns + 255) / 256);
    cudaMemsetAsync(c->d_ghost, 0, (size_t)ns, s);
    k_flag_ghosts<<<blocks, 256, 0, s>>>(p, (int)(c->tick & 1), c->d_ghost);
    k_occupancy<<<blocks, 256, 0, s>>>(p, c->d_ghost, d_lane, d_junc);
    c->launches += 2;
  }
  if (d_lane) { const int rc = sparse_download(c, lane_count, (const int32_t*)d_lane, s); if (rc) return done(rc); }
  if (d_junc) { const int rc = sparse_download(c, junction_count, (const int32_t*)d_junc, s); if (rc) return done(rc); }
  const cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return done(fail(c, GRIDDY_ERR_CUDA, std::string("download_occupancy: ") + cudaGetErrorString(e)));
  return done(GRIDDY_OK);
}

// ---- multi-GPU set-up --------------------------------------------------------------------------
int griddy_mg_plan(int64_t n_items, const uint8_t* kind, const int32_t* lane_in, const int32_t* lane_out,
                   const int32_t* lane_junction, const int32_t* owner, int32_t owner_rank, int32_t reader_rank,
                   int32_t* lanes, int64_t* n_lanes, int32_t* junctions, int64_t* n_junctions) {
  if (n_items < 0 || !kind || !lane_in || !lane_out || !lane_junction || !owner || !n_lanes || !n_junctions)
    return GRIDDY_ERR_BAD_ARG;
  Tiling tl;
  tl.whole(n_items);
  const NetView v{&tl, kind, lane_in, lane_out, lane_junction, owner};
  std::vector<int32_t> l, j;
  mg_plan(v, owner_rank, reader_rank, l, j);
  if (lanes) std::copy(l.begin(), l.end(), lanes);
  if (junctions) std::copy(j.begin(), j.end(), junctions);
  *n_lanes = (int64_t)l.size();
  *n_junctions = (int64_t)j.size();
  return GRIDDY_OK;
}

int griddy_mg_plan_tile(int64_t n_items_global, int32_t n_ranges, const int64_t* range_beg, const int64_t* range_end,
                        const uint8_t* kind, const int32_t* lane_in, const int32_t* lane_out, const int32_t* lane_junction,
                        const int32_t* owner, int32_t owner_rank, int32_t reader_rank, int32_t* lanes, int64_t* n_lanes,
                        int32_t* junctions, int64_t* n_junctions, int64_t* n_feeders) {
  if (n_items_global < 0 || n_ranges < 1 || !range_beg || !range_end || !kind || !lane_in || !lane_out || !lane_junction ||
      !owner || !n_lanes || !n_junctions)
    return GRIDDY_ERR_BAD_ARG;
  Tiling tl;
  tl.n_global = n_items_global;
  for (int k = 0; k < n_ranges; ++k) {
    if (range_end[k] < range_beg[k] || (k > 0 && range_beg[k] < range_end[k - 1]) || range_end[k] > n_items_global) return GRIDDY_ERR_BAD_ARG;
    tl.beg.push_back(range_beg[k]);
    tl.end.push_back(range_end[k]);
    tl.off.push_back(tl.n_local);
    tl.n_local += range_end[k] - range_beg[k];
  }
  const NetView v{&tl, kind, lane_in, lane_out, lane_junction, owner};
  std::vector<int32_t> l, j;
  mg_plan(v, owner_rank, reader_rank, l, j);
  if (lanes) std::copy(l.begin(), l.end(), lanes);
  if (junctions) std::copy(j.begin(), j.end(), junctions);
  *n_lanes = (int64_t)l.size();
  *n_junctions = (int64_t)j.size();
  if (n_feeders) *n_feeders = mg_count_feeders(v, owner_rank, reader_rank);
  return GRIDDY_OK;
}

int griddy_mg_configure(void* ctx, int32_t rank, int32_t world, const int32_t* owner, const int32_t* lane_in,
                        int32_t mig_cap, int64_t extra_slots, int32_t max_agenda_items) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_net) return fail(c, GRIDDY_ERR_BAD_ARG, "mg_configure before upload_network");
  if (world < 1 || world > MAX_RANKS || rank < 0 || rank >= world || !owner || !lane_in || mig_cap < 1 || extra_slots < 0 || max_agenda_items < 0)
    return fail(c, GRIDDY_ERR_BAD_ARG, "mg_configure: bad arguments");
  CUDA_TRY(c, cudaSetDevice(c->device));
  const Tiling& tl = c->tl;
  const int64_t n_local = tl.n_local;
  for (int64_t r = 0; r < n_local; ++r) {
    if (owner[r] < 0 || owner[r] >= world) return fail(c, GRIDDY_ERR_BAD_ARG, "owner out of range");
    // a segment and its lanes share an owner, a junction and its lanes too: actors go on and off the
    // road, and junction occupancy is counted, without leaving the rank
    if (c->h_kind[r] == GRIDDY_JUNCTION_LANE) {
      const int64_t lj = tl.local(c->h_lane_junction[r]);
      if (lj >= 0 ? owner[r] != owner[lj] : owner[r] == rank)
        return fail(c, GRIDDY_ERR_BAD_ARG, "junction lane (entry " + std::to_string(r) + ") and its junction have different owners, or the junction is not held");
    }
  }
  free_pool(c->actor_allocs);
  c->have_actors = false;
  mg_reset(c);
  Multi& m = c->mg;
  m.on = true;
  m.rank = rank;
  m.world = world;
  m.mig_cap = mig_cap;
  m.max_agenda_items = std::max(max_agenda_items, 1);
  m.extra_slots = extra_slots;
  c->mg_owner_host.assign(owner, owner + n_local);
  {   // owner, by global id like every per-item device array (left in net_sparse: it lives as long as the network)
    int32_t* d_owner = nullptr;
    TRY(sparse_alloc(c, sizeof(int32_t), (void**)&d_owner));
    TRY(sparse_upload(c, d_owner, owner));
    m.d_owner = d_owner;
  }
  CUDA_TRY(c, cudaEventCreateWithFlags(&m.ev_ready, cudaEventDisableTiming));
  const NetView v{&tl, c->h_kind.data(), lane_in, c->h_lane_out.data(), c->h_lane_junction.data(), owner};
  // what I export to / import from every other rank
  struct Lists { std::vector<int32_t> exp_l, exp_j, imp_l, imp_j; };
  std::vector<Lists> lists(world);
  std::vector<int32_t> exp_all, imp_junc_all, imp_where_all;
  size_t off = 0;   // index into halo_recv (values)
  for (int r = 0; r < world; ++r) {
    if (r == rank) continue;
    Lists& L = lists[r];
    mg_plan(v, rank, r, L.exp_l, L.exp_j);
    mg_plan(v, r, rank, L.imp_l, L.imp_j);
    if (L.exp_l.empty() && L.exp_j.empty() && L.imp_l.empty() && L.imp_j.empty()) continue;
    Neighbor nb;
    nb.rank = r;
    nb.n_exp_lanes = (int)L.exp_l.size();
    nb.n_exp_junc = (int)L.exp_j.size();
    nb.n_imp_lanes = (int)L.imp_l.size();
    nb.n_imp_junc = (int)L.imp_j.size();
    nb.exp_first = (int)exp_all.size();
    exp_all.insert(exp_all.end(), L.exp_l.begin(), L.exp_l.end());
    for (int32_t j : L.exp_j) exp_all.push_back(~j);
    nb.halo_off[0] = off;   // in values for now; bytes below
    for (size_t k = 0; k < L.imp_l.size(); ++k) m.import_marks.emplace_back(L.imp_l[k], (int32_t)(-2 - (int64_t)(off + k)));
    for (size_t k = 0; k < L.imp_j.size(); ++k) {
      imp_junc_all.push_back(L.imp_j[k]);
      imp_where_all.push_back((int32_t)(off + L.imp_l.size() + k));
    }
    off += L.imp_l.size() + L.imp_j.size();
    nb.mig_out = (int)std::min<int64_t>(mig_cap, mg_count_feeders(v, rank, r));
    nb.mig_in = (int)std::min<int64_t>(mig_cap, mg_count_feeders(v, r, rank));
    nb.item_out = (int64_t)nb.mig_out * m.max_agenda_items;
    nb.item_in = (int64_t)nb.mig_in * m.max_agenda_items;
    m.nb.push_back(nb);
  }
  // my window: [halo values, parity 0][halo values, parity 1][per neighbour: halo flag x 2 (64 bytes each)]
  //            [per neighbour and parity: MigHeader, records, items]
  auto align = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t halo_bytes = align(off * sizeof(double));
  size_t w = 0;
  m.halo_base[0] = w; w += halo_bytes;
  m.halo_base[1] = w; w += halo_bytes;
  for (Neighbor& nb : m.nb) {
    const size_t first = nb.halo_off[0];   // value index
    for (int q = 0; q < 2; ++q) nb.halo_off[q] = m.halo_base[q] + first * sizeof(double);
    for (int q = 0; q < 2; ++q) { nb.hflag_off[q] = w; w += 64; }
  }
  for (Neighbor& nb : m.nb)
    for (int q = 0; q < 2; ++q) { nb.mig_off[q] = w; w += align(nb.mig_bytes_in()); }
  m.window_bytes = std::max<size_t>(w, 256);
  TRY(mg_alloc(c, &m.window, m.window_bytes));   // plain cudaMalloc: shared with the neighbours through cudaIpc
  CUDA_TRY(c, cudaMemset(m.window, 0, m.window_bytes));
  Params& p = c->p;
  for (int q = 0; q < 2; ++q) p.halo_recv[q] = (const double*)(m.window + m.halo_base[q]);
  for (Neighbor& nb : m.nb) {
    for (int q = 0; q < 2; ++q) {
      p.hflag_mine[nb.rank][q] = (const int32_t*)(m.window + nb.hflag_off[q]);
      p.mig_mine[nb.rank][q] = (int32_t*)(m.window + nb.mig_off[q]);
    }
    p.halo_first[nb.rank] = nb.exp_first;
    p.halo_n[nb.rank] = nb.n_exp_lanes + nb.n_exp_junc;
    p.mig_cap[nb.rank] = nb.mig_out;
    p.mig_icap[nb.rank] = (int32_t)std::min<int64_t>(nb.item_out, INT32_MAX);
    p.mig_in[nb.rank] = nb.mig_in;
  }
  TRY(mg_alloc(c, &p.mig_done, 1));
  TRY(mg_alloc(c, &p.halo_done, 1));
  TRY(mg_alloc(c, &p.mig_count, 2 * MAX_RANKS));
  CUDA_TRY(c, cudaMemset(p.mig_done, 0, sizeof(int32_t)));
  CUDA_TRY(c, cudaMemset(p.halo_done, 0, sizeof(int32_t)));
  CUDA_TRY(c, cudaMemset(p.mig_count, 0, 2 * MAX_RANKS * sizeof(int32_t)));
  m.n_exp = (int)exp_all.size();
  m.n_imp_junc = (int)imp_junc_all.size();
  TRY(mg_alloc(c, &m.d_exp_items, exp_all.size()));
  TRY(mg_alloc(c, &m.d_imp_junc, imp_junc_all.size()));
  TRY(mg_alloc(c, &m.d_imp_where, imp_where_all.size()));
  if (m.n_exp) CUDA_TRY(c, cudaMemcpy(m.d_exp_items, exp_all.data(), exp_all.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  if (m.n_imp_junc) {
    CUDA_TRY(c, cudaMemcpy(m.d_imp_junc, imp_junc_all.data(), imp_junc_all.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(m.d_imp_where, imp_where_all.data(), imp_where_all.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  {   // segment lanes that reach a junction of another rank (META_TOJ_REMOTE)
    for (size_t k = 0; k < c->tl.beg.size(); ++k) {
      const int64_t n_k = c->tl.end[k] - c->tl.beg[k];
      if (n_k > 0) k_lane_clear_remote<<<(unsigned)((n_k + 255) / 256), 256>>>(p.lane + c->tl.beg[k], n_k);
    }
    std::vector<int32_t> marks;
    for (size_t k = 0; k < c->tl.beg.size() && !c->h_to.empty(); ++k)
      for (int64_t g = c->tl.beg[k]; g < c->tl.end[k]; ++g) {
        const int64_t l = c->tl.off[k] + (g - c->tl.beg[k]);
        if (c->h_kind[(size_t)l] == GRIDDY_SEGMENT_LANE && c->h_to[(size_t)l] >= 0 && c->owner_of(c->h_to[(size_t)l]) != rank)
          marks.push_back((int32_t)g);
      }
    if (!marks.empty()) {
      int32_t* d_marks = nullptr;
      TRY(mg_alloc(c, &d_marks, marks.size()));
      CUDA_TRY(c, cudaMemcpy(d_marks, marks.data(), marks.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
      k_lane_mark_remote<<<(unsigned)((marks.size() + 255) / 256), 256>>>(p.lane, d_marks, (int64_t)marks.size());
      CUDA_TRY(c, cudaDeviceSynchronize());
    }
  }
  p.owner = m.d_owner;
  p.imp_junc = m.d_imp_junc;
  p.imp_where = m.d_imp_where;
  p.n_imp_junc = m.n_imp_junc;
  p.halo_items = m.d_exp_items;
  p.halo_total = m.n_exp;
  p.my_rank = rank;
  m.connected = m.nb.empty();
  return GRIDDY_OK;
}

int griddy_mg_set_routes(void* ctx, int64_t n_routes, const int64_t* route_off, const int32_t* route_lane, int64_t n_items,
                         const int32_t* item_route) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->mg.on) return fail(c, GRIDDY_ERR_BAD_ARG, "mg_set_routes before mg_configure");
  if (n_routes < 0 || n_items < 0 || !route_off || (n_items > 0 && !item_route) || route_off[0] != 0)
    return fail(c, GRIDDY_ERR_BAD_ARG, "mg_set_routes: bad arguments");
  for (int64_t k = 0; k < n_routes; ++k)
    if (route_off[k + 1] < route_off[k]) return fail(c, GRIDDY_ERR_BAD_ARG, "mg_set_routes: route_off not monotone");
  if (route_off[n_routes] > 0 && !route_lane) return fail(c, GRIDDY_ERR_BAD_ARG, "mg_set_routes: route_lane is null");
  c->mg_route_off.assign(route_off, route_off + n_routes + 1);
  c->mg_route_lane.assign(route_lane, route_lane + route_off[n_routes]);
  c->mg_item_route.assign(item_route, item_route + n_items);
  c->have_mg_routes = true;
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

int griddy_mg_neighbors(void* ctx, int32_t* ranks, int32_t* n) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->mg.on || !n) return fail(c, GRIDDY_ERR_BAD_ARG, "mg_neighbors before mg_configure");
  *n = (int32_t)c->mg.nb.size();
  if (ranks)
    for (size_t k = 0; k < c->mg.nb.size(); ++k) ranks[k] = c->mg.nb[k].rank;
  return GRIDDY_OK;
}

namespace {

// Where `nb` writes in my window, as the 6 offsets that travel with the window's handle.
void window_offsets(const Neighbor& nb, int64_t out[6]) {
  for (int q = 0; q < 2; ++q) {
    out[q] = (int64_t)nb.halo_off[q];
    out[2 + q] = (int64_t)nb.hflag_off[q];
    out[4 + q] = (int64_t)nb.mig_off[q];
  }
}

// The neighbour's window is mapped at `base`: point the kernels at my places in it.
void bind_peer(Ctx* c, Neighbor& nb, uint8_t* base, const int64_t off[6]) {
  nb.peer_base = base;
  for (int q = 0; q < 2; ++q) {
    nb.peer_halo_off[q] = (size_t)off[q];
    nb.peer_hflag_off[q] = (size_t)off[2 + q];
    nb.peer_mig_off[q] = (size_t)off[4 + q];
    c->p.halo_peer[nb.rank][q] = (double*)(base + off[q]);
    c->p.hflag_peer[nb.rank][q] = (int32_t*)(base + off[2 + q]);
    c->p.mig_peer[nb.rank][q] = base + off[4 + q];
  }
  nb.mapped = true;
  bool all = true;
  for (const Neighbor& o : c->mg.nb) all = all && o.mapped;
  c->mg.connected = all;
}

}  // namespace

// One process per GPU: my window is handed to every neighbour as a cudaIpc handle (64 bytes) followed by
// the six byte offsets (int64) of the places that neighbour writes: halo values, halo flag and migrant area, by
// tick parity.  The neighbour maps it with griddy_mg_ipc_import.
int griddy_mg_ipc_export(void* ctx, int32_t from_rank, uint8_t* blob128) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->mg.on || !blob128) return fail(c, GRIDDY_ERR_BAD_ARG, "mg_ipc_export before mg_configure");
  CUDA_TRY(c, cudaSetDevice(c->device));
  for (Neighbor& nb : c->mg.nb)
    if (nb.rank == from_rank) {
      cudaIpcMemHandle_t h;
      CUDA_TRY(c, cudaIpcGetMemHandle(&h, c->mg.window));
      static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
      memset(blob128, 0, 128);
      memcpy(blob128, &h, 64);
      int64_t off[6];
      window_offsets(nb, off);
      memcpy(blob128 + 64, off, sizeof(off));
      const int32_t plan[2] = {nb.n_imp_lanes + nb.n_imp_junc, nb.mig_in};   // what I expect of it: it must agree
      memcpy(blob128 + 112, plan, sizeof(plan));
      return GRIDDY_OK;
    }
  return fail(c, GRIDDY_ERR_BAD_ARG, "rank " + std::to_string(from_rank) + " is not a neighbour");
}

int griddy_mg_ipc_import(void* ctx, int32_t to_rank, const uint8_t* blob128) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->mg.on || !blob128) return fail(c, GRIDDY_ERR_BAD_ARG, "mg_ipc_import before mg_configure");
  CUDA_TRY(c, cudaSetDevice(c->device));
  for (Neighbor& nb : c->mg.nb)
    if (nb.rank == to_rank) {
      cudaIpcMemHandle_t h;
      memcpy(&h, blob128, 64);
      int64_t off[6];
      memcpy(off, blob128 + 64, sizeof(off));
      int32_t plan[2];
      memcpy(plan, blob128 + 112, sizeof(plan));
      if (plan[0] != nb.n_exp_lanes + nb.n_exp_junc || plan[1] != nb.mig_out)
        return fail(c, GRIDDY_ERR_BAD_STATE, "exchange plans of ranks " + std::to_string(c->mg.rank) + " and " + std::to_string(to_rank) + " disagree");
      void* ptr = nullptr;
      CUDA_TRY(c, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
      bind_peer(c, nb, (uint8_t*)ptr, off);
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
    // same process: the neighbours' windows are directly addressable
    for (Neighbor& nb : c->mg.nb) {
      Neighbor* back = nullptr;
      for (Neighbor& b : cs[nb.rank]->mg.nb)
        if (b.rank == c->mg.rank) back = &b;
      if (!back) return fail(c, GRIDDY_ERR_BAD_STATE, "neighbour lists are not symmetric");
      if (nb.n_exp_lanes + nb.n_exp_junc != back->n_imp_lanes + back->n_imp_junc || nb.mig_out != back->mig_in)
        return fail(c, GRIDDY_ERR_BAD_STATE, "exchange plans of two neighbours disagree");
      int64_t off[6];
      window_offsets(*back, off);
      bind_peer(c, nb, cs[nb.rank]->mg.window, off);
    }
  }
  return GRIDDY_OK;
}

// ---- drawing read-back ---------------------------------------------------------------------------
int griddy_upload_geometry(void* ctx, const double* geom) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_net || !geom) return fail(c, GRIDDY_ERR_BAD_ARG, "upload_geometry: no network or null array");
  CUDA_TRY(c, cudaSetDevice(c->device));
  double2* d = nullptr;   // 4 x double2 per item, by global id like every per-item array; geom is compact
  TRY(sparse_alloc(c, 4 * sizeof(double2), (void**)&d));
  TRY(sparse_upload(c, d, (const double2*)geom, 4));
  for (size_t k = 0; k < c->tl.beg.size(); ++k) {   // mark the straight lanes (readback.cuh, k_geom_mark)
    const int64_t m = c->tl.end[k] - c->tl.beg[k];
    if (m > 0) k_geom_mark<<<(unsigned)((m + 255) / 256), 256>>>(c->p.lane + c->tl.beg[k], d + 4 * c->tl.beg[k], m);
  }
  CUDA_TRY(c, cudaDeviceSynchronize());
  c->d_geom = d;
  return GRIDDY_OK;
}

namespace {

// Enqueue (get-pos actor) of every slot on the compute stream, behind the ticks run so far.
int launch_positions(Ctx* c, double2* xy64, float2* xy32, int32_t* gid) {
  const int32_t ns = c->ns_host;
  if (ns <= 0) return GRIDDY_OK;
  const unsigned blocks = (unsigned)((ns + 255) / 256);
  const int par = (int)(c->tick & 1);
  CUDA_TRY(c, cudaMemsetAsync(c->d_ghost, 0, (size_t)ns, c->stream));
  k_flag_ghosts<<<blocks, 256, 0, c->stream>>>(c->p, par, c->d_ghost);
  k_positions<<<blocks, 256, 0, c->stream>>>(c->p, par, c->d_geom, c->d_ghost, xy64, xy32, gid, (int)ns);
  c->launches += 2;
  CUDA_TRY(c, cudaGetLastError());
  return GRIDDY_OK;
}

int positions_ready(Ctx* c, const char* who) {
  if (!c->have_actors) return fail(c, GRIDDY_ERR_BAD_ARG, std::string(who) + " before upload_actors");
  if (!c->d_geom) return fail(c, GRIDDY_ERR_BAD_ARG, std::string(who) + " before upload_geometry");
  return GRIDDY_OK;
}

}  // namespace

int griddy_download_positions(void* ctx, int64_t capacity, int64_t* n, double* xy64, float* xy32, int32_t* gid) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  TRY(positions_ready(c, "download_positions"));
  if (!n) return fail(c, GRIDDY_ERR_BAD_ARG, "download_positions: null n");
  const int64_t ns = c->ns_host;
  *n = ns;
  if (capacity < ns) return fail(c, GRIDDY_ERR_BAD_ARG, "download_positions: arrays too small, need " + std::to_string(ns));
  if (ns == 0) return GRIDDY_OK;
  CUDA_TRY(c, cudaSetDevice(c->device));
  if (!c->d_xy_sync) TRY(dev_alloc(c, c->actor_allocs, &c->d_xy_sync, 28 * (int64_t)c->p.cap));
  double2* d64 = (double2*)c->d_xy_sync;
  float2* d32 = (float2*)(c->d_xy_sync + 16 * (size_t)c->p.cap);
  int32_t* dg = (int32_t*)(c->d_xy_sync + 24 * (size_t)c->p.cap);
  TRY(launch_positions(c, xy64 ? d64 : nullptr, xy32 ? d32 : nullptr, gid ? dg : nullptr));
  if (xy64) CUDA_TRY(c, cudaMemcpyAsync(xy64, d64, (size_t)ns * sizeof(double2), cudaMemcpyDeviceToHost, c->stream));
  if (xy32) CUDA_TRY(c, cudaMemcpyAsync(xy32, d32, (size_t)ns * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
  if (gid) CUDA_TRY(c, cudaMemcpyAsync(gid, dg, (size_t)ns * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return GRIDDY_OK;
}

int griddy_positions_begin(void* ctx, float* xy32, int64_t capacity) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  TRY(positions_ready(c, "positions_begin"));
  if (!xy32) return fail(c, GRIDDY_ERR_BAD_ARG, "positions_begin: null buffer");
  if (c->frames_begun - c->frames_ended >= 2) return fail(c, GRIDDY_ERR_BAD_STATE, "positions_begin: two frames are in flight already (call griddy_positions_end)");
  const int64_t ns = c->ns_host;
  if (capacity < ns) return fail(c, GRIDDY_ERR_BAD_ARG, "positions_begin: buffer too small, need " + std::to_string(ns));
  CUDA_TRY(c, cudaSetDevice(c->device));
  if (!c->draw_stream) {
    CUDA_TRY(c, cudaStreamCreateWithFlags(&c->draw_stream, cudaStreamNonBlocking));
    for (int f = 0; f < 2; ++f) {
      CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_drawn[f], cudaEventDisableTiming));
      CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_copied[f], cudaEventDisableTiming));
    }
  }
  if (!c->d_xy32[0])
    for (int f = 0; f < 2; ++f) TRY(dev_alloc(c, c->actor_allocs, &c->d_xy32[f], c->p.cap));
  const int f = (int)(c->frames_begun & 1);   // its previous user has been waited for by griddy_positions_end
  if (ns > 0) {
    FrameOut o{};
    o.xy32 = c->d_xy32[f];
    k_frame<false><<<(unsigned)((ns + FRAME_TILE - 1) / FRAME_TILE), FRAME_THREADS, 0, c->stream>>>(
        c->p, (int)(c->tick & 1), c->d_geom, o, (int)ns, 0);
    ++c->launches;
    CUDA_TRY(c, cudaGetLastError());
  }
  CUDA_TRY(c, cudaEventRecord(c->ev_drawn[f], c->stream));
  CUDA_TRY(c, cudaStreamWaitEvent(c->draw_stream, c->ev_drawn[f], 0));
  if (ns > 0) CUDA_TRY(c, cudaMemcpyAsync(xy32, c->d_xy32[f], (size_t)ns * sizeof(float2), cudaMemcpyDeviceToHost, c->draw_stream));
  CUDA_TRY(c, cudaEventRecord(c->ev_copied[f], c->draw_stream));
  c->frame_n[f] = ns;
  ++c->frames_begun;
  return GRIDDY_OK;
}

int griddy_positions_end(void* ctx, int64_t* n) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (c->frames_begun == c->frames_ended) return fail(c, GRIDDY_ERR_BAD_STATE, "positions_end without positions_begin");
  const int f = (int)(c->frames_ended & 1);
  CUDA_TRY(c, cudaSetDevice(c->device));
  CUDA_TRY(c, cudaEventSynchronize(c->ev_copied[f]));
  if (n) *n = c->frame_n[f];
  ++c->frames_ended;
  return GRIDDY_OK;
}

// ---- delta frames ----------------------------------------------------------------------------------
namespace {

// Layout of a delta frame in host memory, for `capacity` slots (include/griddy_cuda.h).
struct DeltaLayout {
  int64_t blocks, mask_off, off_off, xy_off, bytes;
  explicit DeltaLayout(int64_t capacity) {
    blocks = (std::max<int64_t>(capacity, 0) + FRAME_TILE - 1) / FRAME_TILE;
    mask_off = 64;
    off_off = mask_off + blocks * (FRAME_TILE / 32) * 4;
    xy_off = (off_off + blocks * 4 + 63) / 64 * 64;
    bytes = xy_off + blocks * FRAME_TILE * 8;
  }
};
constexpr int64_t DELTA_MAGIC = 0x3156444459524747LL;   // "GGRYDDV1"

}  // namespace

int64_t griddy_delta_frame_bytes(int64_t capacity) { return capacity < 0 ? 0 : DeltaLayout(capacity).bytes; }

int griddy_delta_begin(void* ctx, void* frame, int64_t capacity) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  TRY(positions_ready(c, "delta_begin"));
  if (!frame) return fail(c, GRIDDY_ERR_BAD_ARG, "delta_begin: null buffer");
  if (c->deltas_begun - c->deltas_ended >= 2) return fail(c, GRIDDY_ERR_BAD_STATE, "delta_begin: two frames are in flight already (call griddy_delta_end)");
  CUDA_TRY(c, cudaSetDevice(c->device));
  const int64_t cap = c->p.cap;
  if (!c->draw_stream) {
    CUDA_TRY(c, cudaStreamCreateWithFlags(&c->draw_stream, cudaStreamNonBlocking));
    for (int f = 0; f < 2; ++f) {
      CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_drawn[f], cudaEventDisableTiming));
      CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_copied[f], cudaEventDisableTiming));
    }
  }
  if (!c->ev_delta[0])
    for (int f = 0; f < 2; ++f) {
      CUDA_TRY(c, cudaEventCreate(&c->ev_delta[f]));
      CUDA_TRY(c, cudaEventCreate(&c->ev_fmt0[f]));
      CUDA_TRY(c, cudaEventCreate(&c->ev_copy0[f]));
      CUDA_TRY(c, cudaEventCreate(&c->ev_dcopied[f]));
    }
  TRY(delta_flush(c, false));
  if (!c->h_dcnt) CUDA_TRY(c, cudaHostAlloc((void**)&c->h_dcnt, 2 * sizeof(unsigned int), cudaHostAllocMapped));
  if (!c->d_drawn) {
    const DeltaLayout dl(cap);
    TRY(dev_alloc(c, c->actor_allocs, &c->d_drawn, cap));
    for (int f = 0; f < 2; ++f) {
      TRY(dev_alloc(c, c->actor_allocs, &c->d_dvals[f], cap));
      TRY(dev_alloc(c, c->actor_allocs, &c->d_dmask[f], dl.blocks * (FRAME_TILE / 32)));
      TRY(dev_alloc(c, c->actor_allocs, &c->d_doff[f], dl.blocks));
    }
    TRY(dev_alloc(c, c->actor_allocs, &c->d_dcnt, 4));
    CUDA_TRY(c, cudaMemsetAsync(c->d_dcnt, 0, 4 * sizeof(unsigned int), c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->d_drawn, 0xFF, (size_t)cap * sizeof(unsigned long long), c->stream));   // DRAWN_NONE
    c->delta_reorders_seen = -1;   // the first frame is a key frame
    c->delta_cover = 0;
  }
  // a key frame whenever the slot numbering changed since the last frame; the slots beyond it that earlier frames
  // covered go back to "shows nothing", so that an actor arriving in one of them later counts as changed
  const int key = c->delta_reorders_seen != c->reorders;
  const int64_t ns = c->ns_host;
  if (capacity < ns) return fail(c, GRIDDY_ERR_BAD_ARG, "delta_begin: buffer too small, need " + std::to_string(ns) + " slots");
  if (key && c->delta_cover > ns)
    CUDA_TRY(c, cudaMemsetAsync(c->d_drawn + ns, 0xFF, (size_t)(c->delta_cover - ns) * sizeof(unsigned long long), c->stream));
  const int f = (int)(c->deltas_begun & 1);   // its previous user has been waited for by griddy_delta_end
  c->h_dcnt[f] = 0;
  CUDA_TRY(c, cudaEventRecord(c->ev_fmt0[f], c->stream));
  if (ns > 0) {
    FrameOut o{};
    o.drawn = c->d_drawn;
    o.mask = c->d_dmask[f];
    o.blk_off = c->d_doff[f];
    o.vals = c->d_dvals[f];
    o.counter = c->d_dcnt + 2 * f;
    unsigned int* dev_view = nullptr;
    CUDA_TRY(c, cudaHostGetDevicePointer((void**)&dev_view, c->h_dcnt, 0));
    o.host_count = dev_view + f;
    k_frame<true><<<(unsigned)((ns + FRAME_TILE - 1) / FRAME_TILE), FRAME_THREADS, 0, c->stream>>>(
        c->p, (int)(c->tick & 1), c->d_geom, o, (int)ns, key);
    ++c->launches;
    CUDA_TRY(c, cudaGetLastError());
  }
  CUDA_TRY(c, cudaEventRecord(c->ev_delta[f], c->stream));
  c->delta_dst[f] = (uint8_t*)frame;
  c->delta_ns[f] = ns;
  c->delta_cap[f] = capacity;
  c->delta_tick[f] = c->tick;
  c->delta_key[f] = key;
  c->delta_reorders_seen = c->reorders;
  c->delta_cover = key ? (int32_t)ns : std::max(c->delta_cover, (int32_t)ns);
  ++c->deltas_begun;
  return GRIDDY_OK;
}

namespace {

// Start the copies of every frame whose kernel has finished (wait: of every frame begun, waiting for its kernel): the
// size of a frame is known only then.  Called from griddy_step_async and griddy_delta_begin too, so that in a frame
// loop the copy of frame t is on its way before the host enqueues tick t+2 — the copy engine never waits for the host.
int delta_flush(Ctx* c, bool wait) {
  while (c->deltas_flushed < c->deltas_begun) {
    const int f = (int)(c->deltas_flushed & 1);
    if (wait) CUDA_TRY(c, cudaEventSynchronize(c->ev_delta[f]));
    else if (cudaEventQuery(c->ev_delta[f]) != cudaSuccess) { cudaGetLastError(); return GRIDDY_OK; }
    const int64_t n = c->delta_ns[f] > 0 ? (int64_t)*(volatile unsigned int*)&c->h_dcnt[f] : 0;
    const DeltaLayout dl(c->delta_cap[f]);
    uint8_t* dst = c->delta_dst[f];
    const int64_t blocks = (c->delta_ns[f] + FRAME_TILE - 1) / FRAME_TILE;
    CUDA_TRY(c, cudaEventRecord(c->ev_copy0[f], c->draw_stream));
    if (n > 0) CUDA_TRY(c, cudaMemcpyAsync(dst + dl.xy_off, c->d_dvals[f], (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost, c->draw_stream));
    if (blocks > 0) {
      CUDA_TRY(c, cudaMemcpyAsync(dst + dl.mask_off, c->d_dmask[f], (size_t)blocks * (FRAME_TILE / 32) * 4, cudaMemcpyDeviceToHost, c->draw_stream));
      CUDA_TRY(c, cudaMemcpyAsync(dst + dl.off_off, c->d_doff[f], (size_t)blocks * 4, cudaMemcpyDeviceToHost, c->draw_stream));
    }
    CUDA_TRY(c, cudaEventRecord(c->ev_dcopied[f], c->draw_stream));
    int64_t* hdr = (int64_t*)dst;   // (the header is the host's own knowledge: written here, not copied)
    hdr[0] = DELTA_MAGIC;
    hdr[1] = c->delta_ns[f];
    hdr[2] = n;
    hdr[3] = c->delta_tick[f];
    hdr[4] = c->delta_key[f];
    hdr[5] = blocks;
    hdr[6] = c->delta_cap[f];
    hdr[7] = 0;
    c->delta_n[f] = n;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev_fmt0[f], c->ev_delta[f]) == cudaSuccess) c->last_frame_ms = ms;
    ++c->deltas_flushed;
  }
  return GRIDDY_OK;
}

}  // namespace

int griddy_delta_end(void* ctx, int64_t* n_slots, int64_t* n_changed) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (c->deltas_begun == c->deltas_ended) return fail(c, GRIDDY_ERR_BAD_STATE, "delta_end without delta_begin");
  const int f = (int)(c->deltas_ended & 1);
  CUDA_TRY(c, cudaSetDevice(c->device));
  if (c->deltas_flushed == c->deltas_ended) {   // its copies are not on their way yet: wait for the frame to be formatted
    CUDA_TRY(c, cudaEventSynchronize(c->ev_delta[f]));
    TRY(delta_flush(c, false));
  }
  CUDA_TRY(c, cudaEventSynchronize(c->ev_dcopied[f]));
  float copy_ms = 0.f;
  if (cudaEventElapsedTime(&copy_ms, c->ev_copy0[f], c->ev_dcopied[f]) == cudaSuccess) c->last_copy_ms = copy_ms;
  if (n_slots) *n_slots = c->delta_ns[f];
  if (n_changed) *n_changed = c->delta_n[f];
  ++c->deltas_ended;
  return GRIDDY_OK;
}

int griddy_delta_apply(const void* frame, float* xy32, int64_t capacity) {
  if (!frame || !xy32) return GRIDDY_ERR_BAD_ARG;
  const int64_t* hdr = (const int64_t*)frame;
  if (hdr[0] != DELTA_MAGIC || hdr[1] < 0 || hdr[1] > capacity || hdr[6] < hdr[1]) return GRIDDY_ERR_BAD_ARG;
  const DeltaLayout dl(hdr[6]);
  const uint8_t* base = (const uint8_t*)frame;
  const uint32_t* mask = (const uint32_t*)(base + dl.mask_off);
  const uint32_t* blk_off = (const uint32_t*)(base + dl.off_off);
  const float* xy = (const float*)(base + dl.xy_off);
  if (hdr[4]) {   // key frame: slot numbers mean something new
    uint32_t* w = (uint32_t*)xy32;
    for (int64_t k = 0; k < 2 * capacity; ++k) w[k] = 0x7fc00000u;
  }
  int64_t applied = 0;
  for (int64_t b = 0; b < hdr[5]; ++b) {
    const float* src = xy + 2 * (int64_t)blk_off[b];
    for (int w = 0; w < FRAME_TILE / 32; ++w) {
      uint32_t m = mask[b * (FRAME_TILE / 32) + w];
      const int64_t slot0 = b * FRAME_TILE + 32 * w;
      while (m) {
        const int64_t slot = slot0 + __builtin_ctz(m);
        m &= m - 1;
        if (slot >= capacity) return GRIDDY_ERR_BAD_ARG;
        xy32[2 * slot] = src[0];
        xy32[2 * slot + 1] = src[1];
        src += 2;
        ++applied;
      }
    }
  }
  return applied == hdr[2] ? GRIDDY_OK : GRIDDY_ERR_BAD_STATE;
}

void* griddy_host_alloc(int64_t bytes) {
  void* ptr = nullptr;
  if (bytes <= 0 || cudaHostAlloc(&ptr, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return ptr;
}

void griddy_host_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
}

int griddy_debug_trace(void* ctx, uint64_t* stamps) {
  Ctx* c = (Ctx*)ctx;
  if (!c || !stamps) return GRIDDY_ERR_BAD_ARG;
#ifdef GRIDDY_TRACE
  if (!c->p.trace) return fail(c, GRIDDY_ERR_BAD_STATE, "debug_trace: no actors uploaded");
  CUDA_TRY(c, cudaSetDevice(c->device));
  CUDA_TRY(c, cudaMemcpy(stamps, c->p.trace, 64 * TR_COUNT * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return GRIDDY_OK;
#else
  return fail(c, GRIDDY_ERR_BAD_STATE, "debug_trace: the library was not built with -DGRIDDY_TRACE");
#endif
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

int griddy_state_digest(void* ctx, uint64_t* out3) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return GRIDDY_ERR_BAD_ARG;
  if (!c->have_actors || !out3) return fail(c, GRIDDY_ERR_BAD_ARG, "state_digest: no actors or null output");
  CUDA_TRY(c, cudaSetDevice(c->device));
  out3[0] = out3[1] = out3[2] = 0;
  const int32_t ns = c->ns_host;
  if (ns == 0) return GRIDDY_OK;
  const unsigned blocks = (unsigned)((ns + 255) / 256);
  unsigned long long* d_out = (unsigned long long*)c->sort_keys;   // reorder scratch, idle between steps
  CUDA_TRY(c, cudaMemsetAsync(d_out, 0, 3 * sizeof(unsigned long long), c->stream));
  CUDA_TRY(c, cudaMemsetAsync(c->d_ghost, 0, (size_t)ns, c->stream));
  k_flag_ghosts<<<blocks, 256, 0, c->stream>>>(c->p, (int)(c->tick & 1), c->d_ghost);
  k_digest<<<blocks, 256, 0, c->stream>>>(c->p, (int)(c->tick & 1), c->d_ghost, d_out);
  c->launches += 2;
  unsigned long long h[3] = {0, 0, 0};
  CUDA_TRY(c, cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  for (int k = 0; k < 3; ++k) out3[k] = (uint64_t)h[k];
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
double griddy_last_frame_ms(void* ctx) { return ctx ? ((Ctx*)ctx)->last_frame_ms : 0.0; }
double griddy_last_copy_ms(void* ctx) { return ctx ? ((Ctx*)ctx)->last_copy_ms : 0.0; }
int64_t griddy_tick(void* ctx) { return ctx ? ((Ctx*)ctx)->tick : 0; }
int64_t griddy_kernel_launches(void* ctx) { return ctx ? ((Ctx*)ctx)->launches : 0; }
int64_t griddy_slots_in_use(void* ctx) { return ctx ? ((Ctx*)ctx)->ns_host : 0; }
int64_t griddy_reorders(void* ctx) { return ctx ? ((Ctx*)ctx)->reorders : 0; }

}  // extern "C"
