/* griddy_cuda.h — C ABI of the B200 actor-update path (libgriddy_cuda.so).
 *
 * The reference (eeshugerman/griddy) is pure Guile Scheme and has no FFI of its own.  The seam
 * this library plugs into is the `update` closure of `simulate` (griddy/simulate.scm:141-153):
 * instead of rebuilding the world and calling `advance$` per actor (simulate.scm:93-111,
 * simulate/route.scm:53-141), the caller flattens `(make-skeleton)` + actors once, then calls
 * griddy_step() per tick and reads state back when it wants to draw (simulate.scm:155-160).
 * Guile binds these symbols with (system foreign) `pointer->procedure`; see INTEGRATION.md and
 * guile/griddy/cuda.scm.  Plain pointers and sizes only: no structs by value, no C++ types.
 *
 * Conventions
 *   - Every item id is its REGISTRATION INDEX: the order of `add!` calls in make-skeleton
 *     (world.scm:21-22).  Guile's `static-items` list is consed, so list position i is id N-1-i.
 *   - Actor ids are the order of the `link!` calls that placed them (core.scm:169-178): linking the
 *     actors in id order must rebuild the world (later ids are nearer the head of a segment's list;
 *     on a lane a later id replaces an earlier one with an equal pos-param).
 *   - The caller owns every host buffer; the library copies before returning.  The library owns
 *     all device memory.  One caller thread per context; calls block.
 *   - Every function returns 0 on success or a GRIDDY_ERR_* code; griddy_last_error() describes it.
 *     There is no CPU fallback: without a CUDA device griddy_create() fails with GRIDDY_ERR_CUDA.
 */
#ifndef GRIDDY_CUDA_H
#define GRIDDY_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRIDDY_ABI_VERSION 6

enum {
  GRIDDY_OK = 0,
  GRIDDY_ERR_BAD_ARG = 1,
  GRIDDY_ERR_CUDA = 2,
  GRIDDY_ERR_NCCL = 3,
  GRIDDY_ERR_BAD_STATE = 4, /* the no-matching-clause / no-applicable-method cases of
                               simulate.scm:105-111, 80-85 and location.scm:41-47 */
  GRIDDY_ERR_OOM = 5
};

/* item kinds (road.scm:30-72) */
enum { GRIDDY_JUNCTION = 0, GRIDDY_SEGMENT = 1, GRIDDY_SEGMENT_LANE = 2, GRIDDY_JUNCTION_LANE = 3,
       GRIDDY_OTHER_ITEM = 4 };
/* lane direction / road side ('forw 'back), location kind, agenda item kind, route status */
enum { GRIDDY_FORW = 0, GRIDDY_BACK = 1 };
enum { GRIDDY_OFF_ROAD = 0, GRIDDY_ON_ROAD = 1 };
enum { GRIDDY_SLEEP_FOR = 0, GRIDDY_TRAVEL_TO = 1 };
enum { GRIDDY_ROUTE_NONE = 0, GRIDDY_ROUTE_SOME = 1, GRIDDY_ROUTE_DONE = 2 };

int griddy_abi_version(void);

/* A context holds one world on one GPU (device = CUDA ordinal, -1 = current device). */
int griddy_create(void** ctx, int device);
void griddy_destroy(void* ctx);
const char* griddy_last_error(void* ctx); /* owned by ctx, valid until the next call */

/* constants.scm:34,38: dt = (exact->inexact 1/15), two_size = (* 2 *actor/size*) = 10. */
int griddy_set_constants(void* ctx, double dt, double two_size);

/* The static network of (make-skeleton), one entry per registered item.
 *   kind           item kind
 *   lane_len       get-length of a lane (spatial.scm:28-46), computed by the caller in fp64 — the
 *                  device never recomputes geometry
 *   lane_dir       segment lanes: direction
 *   lane_segment   segment lanes: their segment                       (else -1)
 *   lane_out       junction lanes: get-segment-lane, core.scm:106-111 (else -1)
 *   lane_junction  junction lanes: their junction                     (else -1)
 *   seg_outer_forw / seg_outer_back   segments: get-outer-lane for each side, location.scm:16-24 */
int griddy_upload_network(void* ctx, int64_t n_items, const uint8_t* kind, const double* lane_len,
                          const uint8_t* lane_dir, const int32_t* lane_segment,
                          const int32_t* lane_out, const int32_t* lane_junction,
                          const int32_t* seg_outer_forw, const int32_t* seg_outer_back);

/* Optional, before griddy_upload_network: this context holds only PART of the network (one rank's tile of
 * a partitioned world plus the ring of items next to it): the items whose registration index lies in one
 * of the n_ranges half-open ranges [range_beg[k], range_end[k]) — ascending, disjoint, at most 64.  Every
 * per-item array that crosses this ABI afterwards (griddy_upload_network, griddy_set_locality_keys,
 * griddy_set_routing_grid, griddy_upload_geometry, griddy_mg_configure, griddy_download_occupancy) is then
 * COMPACT: one entry per held item, the ranges back to back; n_items of griddy_upload_network is their
 * total length.  The ids INSIDE the arrays (lane_segment, lane_out, turn4, containers, ...) stay global
 * registration indices.  On the device the per-item arrays keep the global index space (virtual address
 * space for every id, memory mapped behind the held ranges only), so the kernels are the same.
 * A tiled context cannot resolve a travel-to destination that lies outside its tile: pass the destinations
 * with griddy_set_agenda_destinations.  n_ranges = 0 returns to an untiled network. */
int griddy_set_item_ranges(void* ctx, int64_t n_items_global, int32_t n_ranges, const int64_t* range_beg,
                           const int64_t* range_end);

/* Optional.  Actors are kept in device "slots" that are re-sorted every few ticks by (locality key
 * of their lane or segment, pos-param) so that neighbours on the road are neighbours in memory.
 * item_key[n_items] gives each item's place in that order (items that are close on the map should
 * get close keys; keys < 2^32-1).  Default: the registration index.  Performance only — results do
 * not depend on it.  Call after griddy_upload_network and before griddy_upload_actors. */
int griddy_set_locality_keys(void* ctx, const uint32_t* item_key);

/* Ticks between two reorders of the actor slots (default 384; 0 = never).  Performance only.  With a period set,
 * a step also begins with a reorder when more than 1/16 of the slots hold actors that are gone. */
int griddy_set_reorder_period(void* ctx, int32_t ticks);
/* Reorder the slots now (between steps); *ms (may be NULL) receives its device time (CUDA events). */
int griddy_reorder_now(void* ctx, double* ms);
int64_t griddy_reorders(void* ctx);          /* reorders run so far */

/* Device-side routing table for procedural grids (examples/procedural-grid.scm): the first n*n
 * items are the junctions in row-major order; lane_from_j / lane_to_j give the junctions a segment
 * lane leaves and reaches; turn4[4*lane + h] is the junction lane leading from `lane` onto the
 * segment that leaves its end junction with heading h (0 +i, 1 -i, 2 +j, 3 -j), or -1.  With it,
 * travel-to needs no uploaded routes: the next hop is dimension-ordered (i first, then j).
 * Replaces find-route (route.scm:17-51) on synthetic grids only; a stated deviation from A*. */
int griddy_set_routing_grid(void* ctx, int32_t n, const int32_t* lane_from_j,
                            const int32_t* lane_to_j, const int32_t* turn4);

/* The actors placed by (add-actors! world) and their agendas (actor.scm:14-37).
 *   loc_kind/container/pos/side   location.scm:26-39 (container = segment off road, lane on road)
 *   speed      max-speed;  speed_dt = (exact->inexact (* max-speed *simulate/time-step*))
 *   agenda_off [n+1]  CSR offsets into the item_* arrays; item_kind sleep-for / travel-to;
 *              item_sleep the time; item_seg/item_side/item_pos the travel-to destination
 *   route_off  [items+1] / route_lane: for each agenda item the (turn-onto lane) steps find-route
 *              returns for it (route.scm:48-51; the final (arrive-at pos) is implied).  Pass NULL
 *              for both to use the grid routing table instead.
 * Actors start with route 'none (actor.scm:19-21). */
int griddy_upload_actors(void* ctx, int64_t n, const uint8_t* loc_kind, const int32_t* container,
                         const double* pos, const uint8_t* side, const double* speed,
                         const double* speed_dt, const int64_t* agenda_off, const uint8_t* item_kind,
                         const int32_t* item_sleep, const int32_t* item_seg, const uint8_t* item_side,
                         const double* item_pos, const int64_t* route_off, const int32_t* route_lane);

/* Optional: find-route on the device (route.scm:17-51).  With the lane graph uploaded, griddy_upload_actors called
 * with route_off = NULL (and no grid routing table) computes the route of every travel-to item itself: A* from the
 * lane the route starts on — the actor's first location, then the destination of the travel-to before
 * (off-road->on-road, location.scm:49-57) — to the item's destination lane.
 *   nbr_off [n_items+1], nbr   the `neighbors` of route.scm:19-23 per lane, in the reference's list order: a segment
 *                              lane's junction lanes (get-junction-lanes, core.scm:97-104), a junction lane's one
 *                              output lane (get-segment-lane, core.scm:106-111); nothing for other items
 *   cost_out [n_items]         `cost` of leaving a lane (route.scm:25-34): (get-length (ref lane 'segment)) for a
 *                              segment lane, 0 for a junction lane
 *   midpoint_xy [2 x n_items]  (get-midpoint lane), spatial.scm:68-73: `distance` is the l2 of two of them
 * The search is the A* that tests/golden/minischeme.py substitutes for Chickadee's un-vendored `a*` (priority = cost
 * so far + distance to the goal, ties by order of insertion), so the routes equal the fixtures' recorded ones.  A
 * travel-to without a path ('no-path) makes the step that reaches it fail with GRIDDY_ERR_BAD_STATE.  Whole
 * networks only (no griddy_set_item_ranges). */
int griddy_set_route_graph(void* ctx, const int64_t* nbr_off, const int32_t* nbr, const double* cost_out,
                           const double* midpoint_xy);
/* The route lists the last griddy_upload_actors used — given by the caller or found on the device: route_off
 * [agenda items + 1], route_lane (`capacity` entries must hold *n_steps); either array may be NULL. */
int griddy_download_routes(void* ctx, int64_t* route_off, int32_t* route_lane, int64_t capacity, int64_t* n_steps);

/* Optional, before griddy_upload_actors: the destination of every agenda item resolved by the caller, as
 * begin-route$ would (simulate.scm:80-85 -> off-road->on-road, location.scm:49-57): dest_lane = get-outer-lane of
 * (item_seg, item_side) (location.scm:16-24; -1: the road has no lanes there), dest_dir that lane's direction,
 * dest_from_j / dest_to_j the junctions it leaves and reaches (grid routing table; may be NULL).  Entries of
 * sleep-for items are ignored.  Without this call the library resolves them from the uploaded network, which
 * needs every destination segment to be held by this context. */
int griddy_set_agenda_destinations(void* ctx, int64_t n_agenda_items, const int32_t* dest_lane,
                                   const uint8_t* dest_dir, const int32_t* dest_from_j, const int32_t* dest_to_j);

/* Advance the world n_ticks ticks: the body of `update`, simulate.scm:144-151, n_ticks times. */
int griddy_step(void* ctx, int64_t n_ticks);
/* The same without waiting for the device: the ticks are enqueued and the call returns.  For frame loops
 * (`update` then `draw` per tick, with griddy_positions_begin / _end): nothing synchronises the host per tick.
 * An error the device meets (GRIDDY_ERR_BAD_STATE) is reported by a later call — at the latest by the next
 * blocking one, e.g. griddy_step(ctx, 0). */
int griddy_step_async(void* ctx, int64_t n_ticks);

/* Read the world back (caller-allocated arrays of n entries; any pointer may be NULL).
 *   alive        0 once an equal-key insert replaced the actor (core.scm:173-178)
 *   agenda_cur   items popped so far;  sleep_left  the head's time if it is (sleep-for t)
 *   route_head   lane of a (turn-onto lane) head, -1 for (arrive-at p), -2 for 'none or '()
 *   order        the `get-actors` order of the world (world.scm:24-30); *n_order actors in it */
int griddy_download_actors(void* ctx, uint8_t* alive, uint8_t* loc_kind, int32_t* container,
                           uint8_t* side, double* pos, int32_t* agenda_cur, int32_t* sleep_left,
                           uint8_t* route_status, int32_t* route_head, int32_t* order,
                           int64_t* n_order);

/* Actors per container (one entry per held item — n_items of griddy_upload_network; lanes and segments) and
 * per junction (sum over its junction lanes — the quantity junction-full? tests, route.scm:98-104).
 * Counted on the device. */
int griddy_download_occupancy(void* ctx, int32_t* lane_count, int32_t* junction_count);

/* ---- Drawing read-back: (get-pos actor) of every actor, computed on the device --------------------
 * All that `draw` reads of an actor is its position (simulate.scm:155-160 -> draw.scm:74-88 ->
 * spatial.scm:125-149), so a frame needs 8 bytes per actor instead of the whole state.
 *
 * geom: 8 doubles per registered item, computed by the host with the reference's own procedures
 * (they go through Chickadee's vec2, like the lane lengths):
 *   segment lane   beg.x beg.y vec.x vec.y 0 0 0 0            (get-pos lane 'beg), (get-vec lane)      spatial.scm:104-107,156-158
 *   junction lane  p0.x p0.y p1.x p1.y p2.x p2.y p3.x p3.y    (ref lane 'curve)                        core.scm:196-206
 *   segment        of.x of.y vec.x vec.y ob.x ob.y 0 0        (get-pos <off-road at pos-param 0>) for the forw and the
 *                                                             back side, (get-vec segment)             spatial.scm:125-136
 *   anything else  ignored
 * The device evaluates  origin + pos-param * vec  (one rounding per operation and component, the
 * reference's order) and, on junction lanes, bezier-curve-point-at. */
int griddy_upload_geometry(void* ctx, const double* geom /* 8 x n_items */);

/* Positions of all slots, in slot order: *n = griddy_slots_in_use entries of 2 values (x, y); a slot
 * without a living actor holds NaN, NaN.  xy64: as the reference's arithmetic gives them; xy32: the
 * same rounded to float, ready for a vertex buffer; gid: the actor's id per slot (-1: none).  Any of
 * the three may be NULL.  `capacity` entries must fit. */
int griddy_download_positions(void* ctx, int64_t capacity, int64_t* n, double* xy64, float* xy32, int32_t* gid);

/* The same for a frame loop, as a pipeline: _begin formats the frame on the device behind the ticks run
 * so far, starts an asynchronous copy of the float pairs into xy32 (pinned host memory, e.g. from
 * griddy_host_alloc; it must stay valid until the matching _end) and returns at once.  The caller may
 * run griddy_step for the next ticks while the copy is in flight.  _end waits for the oldest frame
 * begun and returns its entry count.  At most two frames are in flight. */
int griddy_positions_begin(void* ctx, float* xy32, int64_t capacity);
int griddy_positions_end(void* ctx, int64_t* n);
/* Delta frames: the same pipeline, but a frame carries only the slots whose drawn position differs from the frame
 * before — most of a traffic jam, every sleeper and every idle actor stand still — so that a `draw` loop moves a
 * fraction of the bytes over PCIe.  The caller keeps one persistent array of float x, y per slot (what it draws from,
 * e.g. a vertex buffer) and patches it with every frame, in order (griddy_delta_apply, or its own loop).
 *
 * A frame buffer of griddy_delta_frame_bytes(capacity) bytes (page-locked, griddy_host_alloc) holds, for T = 1024:
 *   bytes 0-63   int64 x 8: magic, n_slots (slots this frame covers), n_changed, tick, key, blocks = ceil(n_slots / T),
 *                capacity, 0
 *   then         uint32 mask[32 * ceil(capacity / T)]: bit l of word j of block b is set when slot b*T + 32*j + l changed
 *   then         uint32 first[ceil(capacity / T)]: index of block b's first value
 *   then         (64-byte aligned) float x, y per changed slot: a block's changed slots in ascending slot order, the
 *                blocks in no particular order (`first` says where each begins).  NaN, NaN: the slot holds no living actor
 *                any more.
 * key = 1 (the first frame, and the first after every reorder of the slots — griddy_reorders counts them): slot
 * numbers mean something new; the frame holds every living actor and the persistent array starts over from all-NaN.
 * _begin formats the frame on the device behind the ticks run so far and returns at once; _end waits for it, copies
 * exactly its n_changed values, mask and table into `frame` and returns.  At most two frames are in flight; the caller
 * runs griddy_step_async for the next ticks in between.  Full frames (griddy_positions_begin, griddy_download_positions)
 * may be taken in between: they do not change what the delta stream considers sent. */
int64_t griddy_delta_frame_bytes(int64_t capacity);
int griddy_delta_begin(void* ctx, void* frame, int64_t capacity);
int griddy_delta_end(void* ctx, int64_t* n_slots, int64_t* n_changed);
/* Host only (no context, no GPU): patch xy32 (2 x capacity floats, capacity >= the frame's n_slots) with a frame. */
int griddy_delta_apply(const void* frame, float* xy32, int64_t capacity);
void* griddy_host_alloc(int64_t bytes);   /* page-locked host memory; NULL on failure */
void griddy_host_free(void* ptr);

/* ---- Multi-GPU: one context per GPU, the world partitioned spatially -----------------------------
 * The reference is single-threaded (readme.md:24-26); this is new.  Every item has an owner rank; a
 * junction and its junction lanes must share one (griddy_mg_configure checks), and so should a
 * segment and its lanes.  An actor lives on the rank that owns its container and changes rank only
 * while it turns from one lane onto another (route.scm:92-135).  Each tick the ranks exchange (1) the
 * halo: for every lane a neighbour's actors can enter, its rearmost positive pos-param
 * (route.scm:118-121), and for every such junction its occupancy (route.scm:98-104), both of the OLD
 * world; (2) the actors that crossed, each as a 128-byte record plus its agenda items not yet popped
 * (32 bytes each; any number).  Results are identical to a single-GPU run of the same world.  Routes come
 * from the grid routing table or from griddy_mg_set_routes.
 *
 * Data plane: every rank owns a WINDOW of device memory that its neighbours' kernels write into directly
 * (peer stores over NVLink / NVSwitch) and then flag; the reader's next kernel waits for the flag.  No
 * message, no collective and no host work per tick — NCCL is not involved.  Needs peer access between the
 * GPUs of the box.
 *
 * Order of calls on every rank: [griddy_set_item_ranges,] griddy_upload_network, griddy_set_routing_grid,
 * griddy_mg_configure, [griddy_set_agenda_destinations,] griddy_upload_actors (only the actors on this
 * rank's items, in ascending global id), griddy_mg_set_global_ids, then one of the two ways to connect:
 *   one process per GPU (torchrun): every rank calls griddy_mg_ipc_export for each of its neighbours
 *       (griddy_mg_neighbors), the 128-byte blobs are exchanged by any means (torch.distributed, MPI, a
 *       file), every rank calls griddy_mg_ipc_import with what each neighbour exported FOR IT; after a
 *       barrier every rank calls griddy_step with the same n_ticks;
 *   griddy_mg_connect_local: all ranks are contexts of ONE process (tests; works on a single GPU too) and
 *       griddy_mg_step_local drives them tick by tick, ordering the ranks' streams with events.
 * Read back with griddy_download_local. */
int griddy_mg_configure(void* ctx, int32_t rank, int32_t world, const int32_t* owner /* one per held item */,
                        const int32_t* lane_in /* one per held item: junction lanes: their input segment lane */,
                        int32_t mig_cap /* ceiling on boundary crossings per neighbour per tick; the exchange is
                                            sized by min(mig_cap, lanes that lead across), which crossings cannot exceed */,
                        int64_t extra_slots /* room for actors that arrive from other ranks */,
                        int32_t max_agenda_items /* sizes the room for migrants' agenda items: records per tick x this */);
int griddy_mg_set_global_ids(void* ctx, const int32_t* gid /* one per uploaded actor, ascending */);
/* Explicit routes on a partitioned world (instead of the grid routing table): the routes of EVERY travel-to item of the
 * whole world — route_off [n_routes + 1] / route_lane as for griddy_upload_actors — are given to every rank, and
 * item_route [n_items] names, for each agenda item of the actors this rank uploads next, its route in that table
 * (ignored for sleep-for items).  An actor that crosses to another rank takes its items' route indices along, so
 * nothing of a route ever travels.  Call after griddy_mg_configure and before griddy_upload_actors, which is then
 * given route_off = NULL.  Whole networks only (no griddy_set_item_ranges). */
int griddy_mg_set_routes(void* ctx, int64_t n_routes, const int64_t* route_off, const int32_t* route_lane, int64_t n_items,
                         const int32_t* item_route);
int griddy_mg_neighbors(void* ctx, int32_t* ranks /* MAX 8, may be NULL */, int32_t* n);
/* blob128: the cudaIpc handle of my window (64 bytes), the byte offsets of the six places `from_rank` writes in it
 * (int64: halo values, halo flag, migrant area; by tick parity) and the sizes I expect of it (2 x int32). */
int griddy_mg_ipc_export(void* ctx, int32_t from_rank, uint8_t* blob128);
int griddy_mg_ipc_import(void* ctx, int32_t to_rank, const uint8_t* blob128);
int griddy_mg_connect_local(void** ctxs, int32_t n);
int griddy_mg_step_local(void** ctxs, int32_t n, int64_t n_ticks);
/* Host-only helper (no GPU needed; arrays over ALL items): the items of owner_rank that actors on
 * reader_rank read before crossing — lanes they can enter, junctions whose occupancy gates them —
 * ascending.  Pass NULL arrays to get the counts.  Both sides derive identical lists, so none is ever exchanged. */
int griddy_mg_plan(int64_t n_items, const uint8_t* kind, const int32_t* lane_in, const int32_t* lane_out,
                   const int32_t* lane_junction, const int32_t* owner, int32_t owner_rank,
                   int32_t reader_rank, int32_t* lanes, int64_t* n_lanes, int32_t* junctions,
                   int64_t* n_junctions);
/* The same from ONE rank's tile (compact arrays over the id ranges it holds, as for griddy_set_item_ranges): what
 * griddy_mg_configure derives on a tiled network.  Owner and reader each see everything next to their own part, so
 * both get the list the whole network gives.  *n_feeders (may be NULL): lanes of owner_rank with a successor on
 * reader_rank — the most actors that can cross that way in one tick, which sizes the migrant area. */
int griddy_mg_plan_tile(int64_t n_items_global, int32_t n_ranges, const int64_t* range_beg, const int64_t* range_end,
                        const uint8_t* kind, const int32_t* lane_in, const int32_t* lane_out,
                        const int32_t* lane_junction, const int32_t* owner, int32_t owner_rank, int32_t reader_rank,
                        int32_t* lanes, int64_t* n_lanes, int32_t* junctions, int64_t* n_junctions, int64_t* n_feeders);
/* Read back this rank's actors: arrays of `capacity` >= griddy_slots_in_use entries, indexed by slot;
 * gid = the actor's global id; alive = 0 for slots whose actor vanished or left for another rank;
 * order = slot numbers in get-actors order of this rank's containers (a container is owned by one
 * rank, so the world's order is the merge of the ranks' orders by descending container id). */
int griddy_download_local(void* ctx, int64_t capacity, int64_t* n_local, int32_t* gid, uint8_t* alive,
                          uint8_t* loc_kind, int32_t* container, uint8_t* side, double* pos,
                          int32_t* agenda_cur, int32_t* sleep_left, uint8_t* route_status,
                          int32_t* route_head, int32_t* order, int64_t* n_order);

/* Like griddy_step, but brackets every kernel launch with CUDA events on the launching stream and
 * adds the per-kernel device time, in ms summed over the n_ticks, to kernel_ms[0..4] =
 * {k_advance, k_unlink, k_link, reorder (all its launches), k_slow}.  For roofline accounting only: the
 * events serialise launches. */
int griddy_profile_step(void* ctx, int64_t n_ticks, double* kernel_ms);

double griddy_last_step_ms(void* ctx);       /* device time of the last griddy_step (CUDA events) */
double griddy_last_frame_ms(void* ctx);      /* device time of the frame kernel of the delta frame whose copy started last */
double griddy_last_copy_ms(void* ctx);       /* time the copies of the delta frame ended last took on their stream */
int64_t griddy_tick(void* ctx);              /* ticks since upload */
int64_t griddy_kernel_launches(void* ctx);   /* kernels launched by this context so far */
/* Actors in the world right now (an actor replaced by an equal-key insert is gone, core.scm:173-178). */
int griddy_count_alive(void* ctx, int64_t* n_alive);
int64_t griddy_slots_in_use(void* ctx);      /* device slots holding actors (living, or dead and not yet dropped) */
/* Digest of the world's state: out3 = {sum, xor, count} over the living actors of a 64-bit hash of the actor's id
 * and of everything griddy_download_actors reports about it (location kind, container, side, pos-param bits, items
 * popped, sleep time left, route status, route head).  Sum and xor commute: the digests of the ranks of a partitioned
 * world add / xor up to the digest of the same world on one GPU.  griddy_b200/digest.py computes the same from a
 * read-back (device or oracle). */
int griddy_state_digest(void* ctx, uint64_t* out3);

/* Diagnostics for tests/: the reorder's stable LSD radix sort (csrc/radix_sort.cuh) applied to n
 * caller-owned (key, value) pairs in host memory, sorting by the low key_bits bits of the key. */
int griddy_debug_sort_pairs(uint64_t* keys, uint32_t* vals, int64_t n, int32_t key_bits);

/* Diagnostics, builds with -DGRIDDY_TRACE only (BAD_STATE otherwise): %globaltimer stamps (ns) of the last 64
 * ticks, 16 per tick (row = tick mod 64): starts of k_advance, k_slow, k_unlink, k_link, k_halo_pack,
 * k_halo_unpack, the emigrant packing, the moment its flags are raised, start of k_immig_unpack, the moment the last
 * neighbour's flag was seen; the rest unused (with one process per GPU the halo and the immigrants are taken in by blocks
 * of k_advance / k_unlink, so the two unpack stamps stay 0).  How the multi-GPU timelines in profiles/ were obtained. */
int griddy_debug_trace(void* ctx, uint64_t* stamps /* 64 x 16 */);

#ifdef __cplusplus
}
#endif
#endif /* GRIDDY_CUDA_H */
