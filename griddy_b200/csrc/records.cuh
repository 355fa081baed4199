// records.cuh — part of griddy_cuda.cu (one translation unit).  Constants, per-slot / per-item records and the kernel parameter block.
#pragma once

namespace griddy {

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
constexpr uint32_t ST_GHOST = 4096u;    // found to be a ghost (its follower holds an equal key): k_unlink of this tick removes it
constexpr uint32_t ST_REDRAW = 8192u;   // container or road side changed since the last delta frame (readback.cuh, k_frame clears it)

constexpr int DEV_ERR_BEGIN_ON_ROAD = 1, DEV_ERR_NO_CLAUSE = 2, DEV_ERR_END_ON_JLANE = 4,
              DEV_ERR_NO_LANE = 8, DEV_ERR_NO_ROUTE = 16, DEV_ERR_INTERNAL = 32, DEV_ERR_MIG_OVERFLOW = 64,
              DEV_ERR_MIG_AGENDA = 128, DEV_ERR_MIG_TIMEOUT = 256, DEV_ERR_LANE_BURST = 512;

// Params::counters[parity][...]
constexpr int CNT_EMIG = 0, CNT_LEFT = 1 /* leaver records */, CNT_SLOW = 2, CNT_ENTERED = 3 /* lanes */, CNT_STRIDE = 4;

// Programmatic dependent launch (sm_90+): the four kernels of a tick are launched with
// programmaticStreamSerializationAllowed, so the blocks of the next one become resident while the last
// blocks of the previous one drain; nothing is read or written before grid_dependency_wait() returns,
// which is when the previous kernel has completed and its writes are visible.
#ifndef GRIDDY_PDL
#define GRIDDY_PDL 1
#endif
// Timeline diagnostics (profiles/r1_mg_timeline.md): compiled in with -DGRIDDY_TRACE only.
enum { TR_ADVANCE, TR_SLOW, TR_UNLINK, TR_LINK, TR_HALO_PACK, TR_HALO_UNPACK, TR_EMIG_PACK, TR_EMIG_FLAG,
       TR_IMMIG_UNPACK, TR_IMMIG_GO, TR_LINK_END, TR_COUNT = 16 };
#ifdef GRIDDY_TRACE
#define TRACE(p, tick, slot)                                                                 \
  do {                                                                                       \
    if ((p).trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {               \
      unsigned long long t_;                                                                 \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                 \
      (p).trace[((tick) & 63) * TR_COUNT + (slot)] = t_;                                     \
    }                                                                                        \
  } while (0)
#else
#define TRACE(p, tick, slot) do { } while (0)
#endif

// L2 residency of the lines a lane change touches in all three mover kernels: k_slow raises the eviction priority
// of the mover's hot record (bit 0 of GRIDDY_L2_KEEP), its old lane's record (bit 1), its new lane's record (bit 2)
// and its new pos-param (bit 3), so that k_unlink / k_link find them in L2 although 300 MB of other traffic passes in
// between.  Measured in profiles/r2/variants.md; more lines than these, or giving the priority back at the last use
// (applypriority), made the tick slower.
#ifndef GRIDDY_L2_KEEP
#define GRIDDY_L2_KEEP 5
#endif
__device__ __forceinline__ void l2_keep(const void* ptr, int bit) {
  if (GRIDDY_L2_KEEP & bit) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(ptr));
}

__device__ __forceinline__ void grid_dependency_wait() {
#if GRIDDY_PDL
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
// device scalars (Params::scal)
constexpr int SC_ERR = 0, SC_NSLOTS = 1, SC_NALIVE = 2, SC_NITEMS = 3, SC_NDEAD = 4, SC_COUNT = 8;   // NDEAD: slots whose actor is gone since the last reorder
constexpr int MAX_RANKS = 8;     // one box
constexpr int MAX_RANGES = 64;   // id ranges of a tiled network (griddy_set_item_ranges)

// The per-slot state every actor reads every tick, one 32-byte sector: k_advance streams it with two 16-byte
// loads, and an actor that changes lane rewrites one sector instead of three.
struct __align__(16) Hot {
  uint32_t st;       // ST_* flags
  int32_t ahead;     // slot of the next actor ahead on the lane (ascending pos-param), -1 at the front
  int32_t behind;    // slot of the next actor behind, -1 at the rear
  int32_t tj;        // on a route: junction of the lane the route head turns onto (-1: none; TJ_REMOTE tag), or, on a
                     // junction lane, -2 - the junction it occupies; asleep (no route): the remaining (sleep-for t) time
  double delta, buf; // fl(speed*dt)*fl(1/len) and 10/len on the current lane (route.scm:62-66), cached at lane entry
};
static_assert(sizeof(Hot) == 32, "one sector per slot");

// Everything about a slot that only matters when its actor does something rare (changes lane, wakes,
// starts or ends a route), in one 64-byte line.  A lane change reads the whole line and writes the first sector;
// k_unlink / k_link touch it only when a lane has several leavers / entrants in one tick, or on a tie.
struct __align__(16) Cold {
  // first sector
  int32_t container;                  // lane (on road) or segment (off road)
  int32_t hop;                        // head of the route: lane of (turn-onto lane), or -1 for (arrive-at _)
  int32_t dest_lane, dest_from_j, dest_to_j;   // grid routing: the route's last lane and its junctions;
                                      // explicit routes: dest_from_j = end of the current route in Params::route_lane
  int32_t mv_old;                     // where a mover came from: lane, or ~segment
  int32_t inbox_next, outbox_next;    // this tick's entrants of a lane, chained (outbox_next: unused)
  // second sector
  double speed, speed_dt;             // max-speed and fl(max-speed * dt)
  int32_t agenda_cur, route_cur;      // agenda head (index into the item pool), route head (explicit routes)
  int32_t ghost_next;                 // (unused)
  int32_t gid;                        // the actor's id (global id on a partitioned world)
};
// The rest of an actor's rarely read state.
struct __align__(16) ColdF {
  double arrive;            // target of the (arrive-at _) that ends the current route
  double land_pos;          // landing stamp (with land_tick, land_lane), see "processing order"
  int32_t agenda_beg, agenda_end;     // the actor's agenda in the item pool
  int32_t land_tick, land_lane;
  int32_t pad[8];
};
static_assert(sizeof(Cold) == 64 && sizeof(ColdF) == 64, "cold records are one 64-byte line each");

// One registered item, one 64-byte line: the static record (first sector: what a lane change reads of its
// target) and, for lanes, the list heads and the row of the grid routing table (second sector).
struct __align__(16) LaneRec {
  double len;             // lanes: length (host-computed, spatial.scm:28-46)
  int32_t link;           // segment lane: its segment; junction lane: the lane it leads onto; segment: outer forw-side lane
  int32_t link2;          // junction lane: its junction; segment: outer back-side lane
  int32_t to_j;           // grid routing: junction a segment lane reaches
  uint32_t meta;          // kind (bits 0-2), direction (bit 3); bit 16: several entrants were pushed this tick (reset by k_link); bit 17: META_TOJ_REMOTE
  double e_pos;           // the new pos-param of this tick's entrant (k_link reads it instead of the entrant's own when there is one)
  int32_t tail;           // rearmost actor; on a partitioned world -2 - k marks a lane another rank owns
  int32_t inbox, outbox, ghosts;   // this tick's entrants (chained through Cold); outbox, ghosts: unused
  int32_t turn[4];        // grid routing: junction lane leading onto the segment that leaves this lane's end with heading h
};
static_assert(sizeof(LaneRec) == 64, "one line per item");
constexpr uint32_t META_KIND = 7u, META_DIR = 8u, META_MULTI = 1u << 16,
                   META_TOJ_REMOTE = 1u << 17;   // partitioned world, segment lanes: another rank owns the junction the lane reaches
__host__ __device__ __forceinline__ int meta_kind(uint32_t m) { return (int)(m & META_KIND); }
__host__ __device__ __forceinline__ uint32_t meta_dir(uint32_t m) { return (m >> 3) & 1u; }

// One agenda item (actor.scm:25-26), one sector.  A travel-to item carries its destination resolved at
// upload time — off-road->on-road of the destination (location.scm:49-57): the outer lane of (segment, side),
// the pos-param on it, the junctions that lane leaves and reaches (grid routing) — so that begin-route$ reads
// nothing of the destination's part of the network, which on a tiled world another rank holds.
struct __align__(16) AgendaItem {
  double dpos;            // travel-to: pos-param on the destination lane
  int32_t sleep;          // sleep-for: the time
  int32_t dlane;          // travel-to: destination lane, -1: the road has no lanes on that side
  int32_t dfrom, dto;     // travel-to: junctions dlane leaves / reaches (-1 without a grid routing table)
  uint32_t kind;          // GRIDDY_SLEEP_FOR / GRIDDY_TRAVEL_TO
  int32_t rid;            // travel-to with explicit routes: its route in Params::route_off (the item's own index on one GPU; an
                          // index into the table every rank holds on a partitioned world — it travels with the item)
};
static_assert(sizeof(AgendaItem) == 32, "one sector per agenda item");

struct Params {
  // network, by registration index
  int64_t n_items;
  LaneRec* lane;
  int32_t* occ;       // actors on a junction's lanes (route.scm:98-104); compact, so that the busy part lives in L2
  const uint32_t* loc_key;   // locality key per item (the reorder's sort key)
  int32_t grid_n;     // > 0: LaneRec::turn / to_j hold the grid routing table
  // slots
  int32_t cap;     // allocated slots
  int32_t* scal;   // SC_*: device error bits, slots in use, live actors counted by the reorder
  Hot* hot;
  double* pos[2];
  Cold* cold;
  ColdF* coldf;
  // agenda items (immutable; actors that arrive from another rank append theirs, a reorder compacts the pool)
  int64_t n_actors;
  AgendaItem* agenda;
  int32_t icap;   // item capacity
  const int64_t* route_off;   // explicit routes, by AgendaItem::rid (null: grid routing table)
  const int32_t* route_lane;
  // per-tick scratch, by slot
  int32_t* touched;   // lanes an actor entered this tick (k_link's work list)
  int4* leavers;      // actors that leave their lane's list this tick: {slot, lane, junction or -1, 0} (k_unlink's work list)
  int4* slow;   // deferred actors: {slot (~slot: turn-onto), st, pos-param lo, hi} as k_advance read them
  int32_t* counters;   // [2 parities][CNT_STRIDE]: emigrants, lanes left, deferred actors, lanes entered
  double dt, two_size;
  // multi-GPU (owner == null on a single GPU): spatial partition, see multi_gpu.cuh.  Every rank owns a
  // WINDOW of device memory that its neighbours write into over NVLink (peer access; cudaIpc between
  // processes): per neighbour and tick parity the halo values with a flag, and the migrants' records and
  // agenda items with a header.  Nothing else crosses: no message, no collective.
  const int32_t* owner;      // rank that owns each item
  int32_t my_rank;
  const double* halo_recv[2];   // by parity: rearmost positive pos-param of the remote lanes my actors can enter
                                // (and the neighbours' junction occupancies); lane.tail = -2 - index into it
  double* halo_peer[MAX_RANKS][2];        // by neighbour and parity: where I write its halo, in ITS window
  int32_t* hflag_peer[MAX_RANKS][2];      //   and the flag I raise there (tick + 1) once it is complete
  const int32_t* hflag_mine[MAX_RANKS][2];   // the flags my neighbours raise in my window
  int32_t halo_first[MAX_RANKS], halo_n[MAX_RANKS];   // my export list, by neighbour: entries [first, first + n)
  const int32_t* halo_items; // my export list: my lane, or ~(my junction), the neighbours' parts back to back
  int32_t halo_total;        //   its length
  int32_t* halo_done;        // halo blocks of k_advance that have finished (the last one raises the flags)
  const int32_t* imp_junc;   // the neighbours' junctions whose occupancy gates my actors, and where in halo_recv each arrives
  const int32_t* imp_where;
  int32_t n_imp_junc;
  int32_t mig_cap[MAX_RANKS];     // by destination rank: records per tick
  int32_t mig_icap[MAX_RANKS];    //   and agenda items per tick
  // by destination rank and tick parity: that rank's receive area for my emigrants, in ITS window:
  // [MigHeader][MigRec x mig_cap][AgendaItem x mig_icap].  emig_pack_block writes records and items there
  // directly, then the header's counts, then raises its flag to tick + 1 — what the receiver's k_immig_unpack waits for.
  uint8_t* mig_peer[MAX_RANKS][2];
  int32_t* mig_mine[MAX_RANKS][2];   // my own receive areas, by source rank
  int32_t mig_in[MAX_RANKS];         // by source rank: records per tick at most
  int32_t* mig_count;                // [2 x MAX_RANKS] places handed out this tick, by destination: records, items
  int32_t* mig_done;                 // blocks of k_slow that have finished (the last one publishes the counts and raises the flags)
  int32_t publish_late;              // one process per GPU: the first block of k_unlink publishes instead (no count, no fence in k_slow)
  unsigned long long* trace;         // -DGRIDDY_TRACE builds: [64 ticks][16] %globaltimer stamps (griddy_debug_trace)
};

}  // namespace griddy
