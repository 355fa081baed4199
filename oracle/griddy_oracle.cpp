// griddy_oracle.cpp — CPU restatement of griddy's per-tick actor update.
//
// *** TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// *** --impl reference legs may load this.  The product path (griddy_b200/) never does.
//
// PARITY PINS: the reference (eeshugerman/griddy, Guile Scheme) ships no tests, golden vectors
// or recorded outputs, and Guile is not installed here, so this restatement cannot be checked
// against a run of the reference.  It is pinned only by (1) the hand-derived closed-form answer for
// examples/simple.scm in tests/test_oracle_kat.py, and (2) tests/golden/, produced by executing the
// reference's own .scm sources under the small Scheme evaluator in tests/golden/ (see README there).
// Load-bearing assumptions about Guile that cannot be verified here: `(* a b c)` folds left to
// right; `bbtree-fold` visits keys in ascending order; chickadee `a*` returns `(start)` when start
// = goal.
//
// Deliberately literal: every lane holds a real ordered map keyed by pos-param (the bbtree of
// road.scm:33-34), every segment a cons-list (road.scm:60-61), a fresh world is filled every tick
// (simulate.scm:141-153) and actors are visited in `get-actors` order (world.scm:24-30).  None of the
// device path's machinery (ahead pointers, stamps, inboxes) appears here.
//
// Build: g++ -O2 -ffp-contract=off -std=c++17 -shared -fPIC (oracle/Makefile).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

namespace {

enum { K_JUNCTION = 0, K_SEGMENT = 1, K_SEGLANE = 2, K_JLANE = 3 };
enum { FORW = 0, BACK = 1 };
enum { ITEM_SLEEP = 0, ITEM_TRAVEL = 1 };
enum { ROUTE_NONE = 0, ROUTE_SOME = 1, ROUTE_DONE = 2 };
enum { ERR_OK = 0, ERR_BAD_ARG = 1, ERR_BAD_STATE = 4 };

struct Actor {
  bool alive = true;
  bool on_road = false;
  int32_t container = -1;   // lane id (on road) or segment id (off road)
  uint8_t side = FORW;      // road-side-direction, off road only
  double pos = 0.0;
  int32_t agenda_cur = 0;   // index of the agenda head in the item arrays
  int32_t sleep_left = 0;   // the head's time when it is (sleep-for time)   simulate.scm:77
  int route_status = ROUTE_NONE;
  size_t route_cur = 0;     // head of the route list held in Oracle::route_lanes / route_arrive
  // lazy grid routes (Oracle::lazy_routes): the list element at index route_cur — a lane, or -1 for the final
  // (arrive-at _) — and the route's last lane; the list itself is never materialised
  int32_t lazy_head = -1, lazy_dest = -1;
};

struct Oracle {
  double dt = 1.0 / 15.0, two_size = 10.0;
  // network, by registration id
  int64_t n_items = 0;
  std::vector<uint8_t> kind, lane_dir;
  std::vector<double> lane_len;
  std::vector<int32_t> lane_segment, lane_out, lane_junction, outer_forw, outer_back;
  // grid routing rule (synthetic configs); explicit routes otherwise
  int grid_n = 0;
  std::vector<int32_t> lane_from_j, lane_to_j, turn4;
  // immutable actor data
  int64_t n = 0;
  std::vector<double> speed, speed_dt;
  std::vector<int64_t> agenda_off, route_off;
  std::vector<uint8_t> item_kind, item_side;
  std::vector<int32_t> item_sleep, item_seg, route_lane;
  std::vector<double> item_pos;
  // world: the containers of the current world and of the one being filled
  std::vector<Actor> actors, actors_next;
  std::vector<std::map<double, int32_t>> lane_actors, lane_actors_next;
  std::vector<std::vector<int32_t>> seg_actors, seg_actors_next;   // back() is the list head
  std::vector<int32_t> filled, filled_next;                         // non-empty containers
  // An actor's route list ((turn-onto lane) ... (arrive-at pos), route.scm:48-51).  The reference
  // deep-copies it every tick (simulate.scm:37-38); each actor is visited once per tick and touches
  // only its own list, so one shared copy behaves identically.
  std::vector<std::vector<int32_t>> route_lanes;
  std::vector<double> route_arrive;
  // With the grid rule the list is a pure function of (lane reached so far, destination): element k + 1
  // follows from element k.  lazy_routes keeps only the head (Actor::lazy_head) and derives the next element
  // when the head is popped — the same sequence find_route unrolls eagerly (tests/test_oracle_properties.py
  // checks the two modes against each other), without 5 kB of lane ids per actor on a 1024 x 1024 grid.
  bool lazy_routes = true;
  // actors on the junction lanes of each junction in the current world (route.scm:98-104), folded
  // once per tick instead of once per query; the old world does not change during a tick.
  std::vector<int32_t> junction_count;
  int64_t ticks = 0, vanished = 0;
  std::string err;
};

// location.scm:41-47
inline double flip_for(const Oracle& o, int32_t lane, double p) {
  return o.lane_dir[lane] == FORW ? p : 1.0 - p;
}
// location.scm:16-24 via the uploaded per-segment table
inline int32_t outer_lane(const Oracle& o, int32_t seg, uint8_t side) {
  return side == FORW ? o.outer_forw[seg] : o.outer_back[seg];
}

// Deterministic dimension-ordered next hop on procedural grids ("Routes on C3-C5").
// Stated deviation from Chickadee A* (not vendored; tie-breaks unknowable).
int32_t grid_next_hop(const Oracle& o, int32_t lane, int32_t dest) {
  const int n = o.grid_n;
  const int32_t J = o.lane_to_j[lane], E = o.lane_from_j[dest];
  const int ji = J / n, jj = J % n, ei = E / n, ej = E % n;
  int h;
  if (J == E) {
    const int32_t T = o.lane_to_j[dest];
    const int ti = T / n, tj = T % n;
    h = ti != ei ? (ti > ei ? 0 : 1) : (tj > ej ? 2 : 3);
  } else if (ji != ei) {
    h = ei > ji ? 0 : 1;
  } else {
    h = ej > jj ? 2 : 3;
  }
  return o.turn4[4 * (int64_t)lane + h];
}

// Element k + 1 of the grid rule's route list given element k (`x`; init = the lane the route starts on, which
// is not part of the list): after a junction lane comes the lane it leads onto, after the destination lane the
// final (arrive-at _) (-1), after any other segment lane the junction lane the rule picks (-2: none).
int32_t grid_route_next(const Oracle& o, int32_t x, int32_t dest_lane) {
  if (o.kind[x] == K_JLANE) return o.lane_out[x];
  if (x == dest_lane) return -1;
  const int32_t jl = grid_next_hop(o, x, dest_lane);
  return jl < 0 ? -2 : jl;
}

// route.scm:17-51 as a route provider: explicit lane lists when uploaded, else the grid rule
// unrolled eagerly into the same list shape.
bool find_route(Oracle& o, int32_t item, int32_t init_lane, int32_t dest_lane,
                std::vector<int32_t>& out) {
  out.clear();
  if (o.grid_n > 0) {
    int32_t lane = init_lane;
    int64_t guard = 0;
    while (lane != dest_lane) {
      const int32_t jl = grid_next_hop(o, lane, dest_lane);
      if (jl < 0 || ++guard > 8 * (int64_t)o.grid_n + 16) return false;
      out.push_back(jl);
      lane = o.lane_out[jl];
      out.push_back(lane);
    }
  } else {
    for (int64_t k = o.route_off[item]; k < o.route_off[item + 1]; ++k)
      out.push_back(o.route_lane[k]);
  }
  return true;
}

// core.scm:173-178: bbtree-set replaces an equal key; the replaced actor is in no container of the
// new world and is never visited again.
void link_on_road(Oracle& o, int32_t a, int32_t lane, double pos) {
  Actor& A = o.actors_next[a];
  A.on_road = true; A.container = lane; A.pos = pos;
  auto& tree = o.lane_actors_next[lane];
  if (tree.empty()) o.filled_next.push_back(lane);
  auto it = tree.find(pos);
  if (it != tree.end()) {
    o.actors_next[it->second].alive = false;
    ++o.vanished;
    it->second = a;
  } else {
    tree.emplace(pos, a);
  }
}

// core.scm:169-171: cons onto the segment's list
void link_off_road(Oracle& o, int32_t a, int32_t seg, uint8_t side, double pos) {
  Actor& A = o.actors_next[a];
  A.on_road = false; A.container = seg; A.side = side; A.pos = pos;
  auto& lst = o.seg_actors_next[seg];
  if (lst.empty()) o.filled_next.push_back(seg);
  lst.push_back(a);
}

// util.scm:152-162: value with the smallest key in (low, high], or -1
int32_t find_min(const std::map<double, int32_t>& tree, double low, double high) {
  auto it = tree.upper_bound(low);
  if (it != tree.end() && it->first <= high) return it->second;
  return -1;
}

// route.scm:53-74.  `delta_speed_time` is (* speed time-step) already evaluated:
// fl(speed * 1/15) on a normal tick, fl(speed * time-remaining) on the carry-over call.
double pos_param_next(const Oracle& o, double current, double delta_speed_time, int32_t lane) {
  const double len = o.lane_len[lane];
  const double buffer = o.two_size / len;
  const double delta = delta_speed_time * (1.0 / len);
  const double next = current + delta;
  const int32_t obstacle = find_min(o.lane_actors[lane], current, next + buffer);
  if (obstacle < 0) return next;
  return o.actors[obstacle].pos - buffer;
}

// load the (sleep-for time) counter when the head changes
void load_head(const Oracle& o, int32_t a, Actor& A) {
  if (A.agenda_cur < o.agenda_off[a + 1] && o.item_kind[A.agenda_cur] == ITEM_SLEEP)
    A.sleep_left = o.item_sleep[A.agenda_cur];
  else
    A.sleep_left = 0;
}

// simulate.scm:93-111
bool advance(Oracle& o, int32_t a) {
  const Actor& old = o.actors[a];
  Actor& nw = o.actors_next[a];
  nw = old;   // (++ actor): copies max-speed, agenda, route   simulate.scm:54-66
  const bool has_head = old.agenda_cur < o.agenda_off[a + 1];
  const int head_kind = has_head ? o.item_kind[old.agenda_cur] : -1;

  auto do_nothing = [&]() {   // simulate.scm:70-72
    if (old.on_road) link_on_road(o, a, old.container, old.pos);
    else link_off_road(o, a, old.container, old.side, old.pos);
  };

  if (!has_head && old.route_status == ROUTE_NONE) { do_nothing(); return true; }

  if (head_kind == ITEM_SLEEP && old.route_status == ROUTE_NONE) {   // simulate.scm:74-78
    if (old.sleep_left == 0) { nw.agenda_cur = old.agenda_cur + 1; load_head(o, a, nw); }
    else nw.sleep_left = old.sleep_left - 1;
    do_nothing();
    return true;
  }

  if (head_kind == ITEM_TRAVEL && old.route_status == ROUTE_NONE) {   // simulate.scm:80-85
    if (old.on_road) { o.err = "begin-route$ on an on-road actor (no applicable method)"; return false; }
    const int32_t it = old.agenda_cur;
    const int32_t init_lane = outer_lane(o, old.container, old.side);
    const int32_t dest_lane = outer_lane(o, o.item_seg[it], o.item_side[it]);
    if (init_lane < 0 || dest_lane < 0) { o.err = "road-has-no-lanes"; return false; }
    const double init_pos = flip_for(o, init_lane, old.pos);
    const double dest_pos = flip_for(o, dest_lane, o.item_pos[it]);
    if (o.grid_n > 0 && o.lazy_routes) {
      nw.lazy_dest = dest_lane;
      nw.lazy_head = grid_route_next(o, init_lane, dest_lane);   // element 0
      if (nw.lazy_head == -2) { o.err = "no route"; return false; }
    } else if (!find_route(o, it, init_lane, dest_lane, o.route_lanes[a])) { o.err = "no route"; return false; }
    o.route_arrive[a] = dest_pos;
    nw.route_cur = 0;
    nw.route_status = ROUTE_SOME;
    link_on_road(o, a, init_lane, init_pos);
    return true;
  }

  if (head_kind == ITEM_TRAVEL && old.route_status == ROUTE_SOME) {   // route.scm:137-141
    const bool lazy = o.grid_n > 0 && o.lazy_routes;
    const std::vector<int32_t>& lanes = o.route_lanes[a];
    const bool head_arrive = lazy ? old.lazy_head == -1 : old.route_cur >= lanes.size();
    const int32_t head_lane = head_arrive ? -1 : (lazy ? old.lazy_head : lanes[old.route_cur]);
    const int32_t lane_current = old.container;
    const double pos_current = old.pos;
    const double next_naive = pos_param_next(o, pos_current, o.speed_dt[a], lane_current);
    auto route_pop = [&]() {   // actor.scm:33-34
      nw.route_cur = old.route_cur + 1;
      if (lazy) {
        if (old.lazy_head == -1) nw.route_status = ROUTE_DONE;
        else nw.lazy_head = grid_route_next(o, old.lazy_head, old.lazy_dest);
      } else if (nw.route_cur > lanes.size()) nw.route_status = ROUTE_DONE;
    };
    if (head_arrive) {     // route.scm:76-90 arrive-at
      const double target_pos = o.route_arrive[a];
      const bool done = next_naive >= target_pos;
      const double pos_next = done ? target_pos : next_naive;
      if (done) route_pop();
      link_on_road(o, a, lane_current, pos_next);
    } else {               // route.scm:92-135 turn-onto
      const int32_t target = head_lane;
      const bool done = next_naive >= 1.0;
      bool waiting = false;
      if (done && o.kind[target] == K_JLANE) {   // junction-full?  route.scm:98-108
        waiting = o.junction_count[o.lane_junction[target]] > 0;
      }
      double pos_next;
      if (waiting) pos_next = pos_current;
      else if (done) {
        const double time_spent = (1.0 - pos_current) / o.speed[a];
        const double time_remaining = o.dt - time_spent;
        pos_next = pos_param_next(o, 0.0, o.speed[a] * time_remaining, target);
      } else pos_next = next_naive;
      const int32_t lane_next = (done && !waiting) ? target : lane_current;
      if (done && !waiting) route_pop();
      if (nw.lazy_head == -2) { o.err = "no route"; return false; }
      link_on_road(o, a, lane_next, pos_next);
    }
    return true;
  }

  if (head_kind == ITEM_TRAVEL && old.route_status == ROUTE_DONE) {   // simulate.scm:87-91
    nw.route_status = ROUTE_NONE;
    o.route_lanes[a].clear(); nw.route_cur = 0;
    nw.agenda_cur = old.agenda_cur + 1;
    load_head(o, a, nw);
    const int32_t lane = old.container;
    if (o.kind[lane] != K_SEGLANE) { o.err = "end-route$ on a junction lane"; return false; }
    link_off_road(o, a, o.lane_segment[lane], o.lane_dir[lane], flip_for(o, lane, old.pos));
    return true;
  }
  o.err = "advance$: no matching clause";   // simulate.scm:105-111
  return false;
}

// world.scm:24-30 + road.scm:102-108: containers in static-items order (reverse registration),
// lanes contribute descending pos-param, segments their list as stored.
void get_actors(const Oracle& o, std::vector<int32_t>& out) {
  out.clear();
  for (int64_t c = o.n_items - 1; c >= 0; --c) {
    if (o.kind[c] == K_SEGMENT) {
      const auto& lst = o.seg_actors[c];
      for (auto it = lst.rbegin(); it != lst.rend(); ++it) out.push_back(*it);
    } else {
      const auto& tree = o.lane_actors[c];
      for (auto it = tree.rbegin(); it != tree.rend(); ++it) out.push_back(it->second);
    }
  }
}

bool tick(Oracle& o) {
  std::vector<int32_t> order;
  get_actors(o, order);
  for (int32_t c : o.filled)
    if (o.kind[c] == K_JLANE) o.junction_count[o.lane_junction[c]] += (int32_t)o.lane_actors[c].size();
  for (int32_t c : o.filled_next) {
    if (o.kind[c] == K_SEGMENT) o.seg_actors_next[c].clear(); else o.lane_actors_next[c].clear();
  }
  o.filled_next.clear();
  // actors that are in no container keep their last record
  for (int32_t a : order)
    if (!advance(o, a)) return false;
  for (int32_t a : order) o.actors[a] = o.actors_next[a];
  for (int32_t c : o.filled)
    if (o.kind[c] == K_JLANE) o.junction_count[o.lane_junction[c]] = 0;
  // an actor replaced during this tick was linked earlier in the same tick: its record moved above
  o.lane_actors.swap(o.lane_actors_next);
  o.seg_actors.swap(o.seg_actors_next);
  o.filled.swap(o.filled_next);
  ++o.ticks;
  return true;
}

}  // namespace

extern "C" {

void* go_create() { return new Oracle(); }
void go_destroy(void* h) { delete (Oracle*)h; }
const char* go_last_error(void* h) { return ((Oracle*)h)->err.c_str(); }

int go_set_constants(void* h, double dt, double two_size) {
  Oracle& o = *(Oracle*)h;
  o.dt = dt; o.two_size = two_size;
  return ERR_OK;
}

int go_load_network(void* h, int64_t n_items, const uint8_t* kind, const double* lane_len,
                    const uint8_t* lane_dir, const int32_t* lane_segment, const int32_t* lane_out,
                    const int32_t* lane_junction, const int32_t* seg_outer_forw,
                    const int32_t* seg_outer_back) {
  Oracle& o = *(Oracle*)h;
  o.n_items = n_items;
  o.kind.assign(kind, kind + n_items);
  o.lane_len.assign(lane_len, lane_len + n_items);
  o.lane_dir.assign(lane_dir, lane_dir + n_items);
  o.lane_segment.assign(lane_segment, lane_segment + n_items);
  o.lane_out.assign(lane_out, lane_out + n_items);
  o.lane_junction.assign(lane_junction, lane_junction + n_items);
  o.outer_forw.assign(seg_outer_forw, seg_outer_forw + n_items);
  o.outer_back.assign(seg_outer_back, seg_outer_back + n_items);
  for (auto* w : {&o.lane_actors, &o.lane_actors_next}) { w->clear(); w->resize(n_items); }
  for (auto* w : {&o.seg_actors, &o.seg_actors_next}) { w->clear(); w->resize(n_items); }
  o.filled.clear(); o.filled_next.clear();
  o.junction_count.assign(n_items, 0);
  return ERR_OK;
}

int go_set_routing_grid(void* h, int n, const int32_t* lane_from_j, const int32_t* lane_to_j,
                        const int32_t* turn4) {
  Oracle& o = *(Oracle*)h;
  o.grid_n = n;
  o.lane_from_j.assign(lane_from_j, lane_from_j + o.n_items);
  o.lane_to_j.assign(lane_to_j, lane_to_j + o.n_items);
  o.turn4.assign(turn4, turn4 + 4 * o.n_items);
  return ERR_OK;
}

// Actors are linked in index order (the order of the `link!` calls in add-actors!).
int go_load_actors(void* h, int64_t n, const uint8_t* loc_kind, const int32_t* container,
                   const double* pos, const uint8_t* side, const double* speed, const double* speed_dt,
                   const int64_t* agenda_off, const uint8_t* item_kind, const int32_t* item_sleep,
                   const int32_t* item_seg, const uint8_t* item_side, const double* item_pos,
                   const int64_t* route_off, const int32_t* route_lane) {
  Oracle& o = *(Oracle*)h;
  o.n = n;
  o.speed.assign(speed, speed + n);
  o.speed_dt.assign(speed_dt, speed_dt + n);
  o.agenda_off.assign(agenda_off, agenda_off + n + 1);
  const int64_t ni = agenda_off[n];
  o.item_kind.assign(item_kind, item_kind + ni);
  o.item_sleep.assign(item_sleep, item_sleep + ni);
  o.item_seg.assign(item_seg, item_seg + ni);
  o.item_side.assign(item_side, item_side + ni);
  o.item_pos.assign(item_pos, item_pos + ni);
  if (route_off) {
    o.route_off.assign(route_off, route_off + ni + 1);
    o.route_lane.assign(route_lane, route_lane + route_off[ni]);
  } else {
    o.route_off.assign(ni + 1, 0);
    o.route_lane.clear();
  }
  o.actors.assign(n, Actor());
  o.actors_next.assign(n, Actor());
  o.route_lanes.assign(n, {});
  o.route_arrive.assign(n, 0.0);
  o.ticks = 0; o.vanished = 0;
  // fill the "next" world, then make it current
  for (int64_t a = 0; a < n; ++a) {
    Actor& A = o.actors_next[a];
    A.agenda_cur = (int32_t)agenda_off[a];
    load_head(o, (int32_t)a, A);
    const int32_t c = container[a];
    if (c < 0 || c >= o.n_items) { o.err = "actor container out of range"; return ERR_BAD_ARG; }
    if (loc_kind[a]) {
      if (o.kind[c] != K_SEGLANE && o.kind[c] != K_JLANE) { o.err = "on-road actor not on a lane"; return ERR_BAD_ARG; }
      link_on_road(o, (int32_t)a, c, pos[a]);
    } else {
      if (o.kind[c] != K_SEGMENT) { o.err = "off-road actor not on a segment"; return ERR_BAD_ARG; }
      link_off_road(o, (int32_t)a, c, side[a], pos[a]);
    }
  }
  o.actors.swap(o.actors_next);
  o.actors_next = o.actors;
  o.lane_actors.swap(o.lane_actors_next);
  o.seg_actors.swap(o.seg_actors_next);
  o.filled.swap(o.filled_next);
  return ERR_OK;
}

// Grid rule only: 1 (default) derives the route head on demand, 0 unrolls the whole list at begin-route$ as
// find-route returns it.  Same results; call before go_load_actors.
int go_set_lazy_routes(void* h, int lazy) {
  ((Oracle*)h)->lazy_routes = lazy != 0;
  return ERR_OK;
}

int go_step(void* h, int64_t n_ticks) {
  Oracle& o = *(Oracle*)h;
  for (int64_t t = 0; t < n_ticks; ++t)
    if (!tick(o)) return ERR_BAD_STATE;
  return ERR_OK;
}

// Per-actor state plus the `get-actors` order of the current world.
// route_head: lane of a (turn-onto lane) head, -1 for an (arrive-at _) head, -2 when the route is
// 'none or '().  order has room for n entries; *n_order receives the number of actors in the world.
int go_get_state(void* h, uint8_t* alive, uint8_t* loc_kind, int32_t* container, uint8_t* side,
                 double* pos, int32_t* agenda_cur, int32_t* sleep_left, uint8_t* route_status,
                 int32_t* route_head, int32_t* order, int64_t* n_order) {
  Oracle& o = *(Oracle*)h;
  for (int64_t a = 0; a < o.n; ++a) {
    const Actor& A = o.actors[a];
    alive[a] = A.alive; loc_kind[a] = A.on_road; container[a] = A.container;
    side[a] = A.on_road ? 0 : A.side; pos[a] = A.pos;
    agenda_cur[a] = (int32_t)(A.agenda_cur - o.agenda_off[a]);
    sleep_left[a] = A.sleep_left;
    route_status[a] = (uint8_t)A.route_status;
    route_head[a] = -2;
    if (A.route_status == ROUTE_SOME) {
      if (o.grid_n > 0 && o.lazy_routes) route_head[a] = A.lazy_head;
      else route_head[a] = A.route_cur < o.route_lanes[a].size() ? o.route_lanes[a][A.route_cur] : -1;
    }
  }
  std::vector<int32_t> ord;
  get_actors(o, ord);
  for (size_t k = 0; k < ord.size(); ++k) order[k] = ord[k];
  *n_order = (int64_t)ord.size();
  return ERR_OK;
}

// road.scm:102-108 lengths: actors per lane and, summed over its junction lanes, per junction.
int go_get_occupancy(void* h, int32_t* lane_count, int32_t* junction_count) {
  Oracle& o = *(Oracle*)h;
  for (int64_t r = 0; r < o.n_items; ++r) { lane_count[r] = 0; junction_count[r] = 0; }
  for (int32_t c : o.filled) {
    if (o.kind[c] == K_SEGMENT) { lane_count[c] = (int32_t)o.seg_actors[c].size(); continue; }
    lane_count[c] = (int32_t)o.lane_actors[c].size();
    if (o.kind[c] == K_JLANE) junction_count[o.lane_junction[c]] += lane_count[c];
  }
  return ERR_OK;
}

// get-pos of an actor (spatial.scm:125-149), stateless: what `draw` computes for every actor
// (draw.scm:74-76, simulate.scm:155-160).  geom holds 8 doubles per item, host-computed by the
// reference's own get-pos / get-vec / curve (include/griddy_cuda.h, griddy_upload_geometry):
//   off road      (+ v-beg v-offset (* pos-param (get-vec segment)))   spatial.scm:125-136; griddy:+ folds
//                 (math.scm:59-62), so this is (v-beg + v-offset) + pos-param * vec, and v-beg + v-offset
//                 is the recorded get-pos at pos-param 0
//   segment lane  (+ (get-pos lane 'beg) (* pos-param (get-vec lane)))   spatial.scm:142-144
//   junction lane (bezier-curve-point-at curve pos-param)   spatial.scm:145-146.  Chickadee is not vendored:
//                 PARITY UNPINNED for this clause beyond the evaluator's stub, whose form is restated —
//                 weights u*u*u, 3*u*u*t, 3*u*t*t, t*t*t (left to right), terms summed from p0 to p3.
// Every vec2 operation is one rounding per component, as in a vec2 of doubles.
int go_get_pos(int64_t n_items, int64_t n, const uint8_t* kind, const double* geom, const uint8_t* loc_kind,
               const int32_t* container, const uint8_t* side, const double* pos, double* xy) {
  for (int64_t a = 0; a < n; ++a) {
    const int32_t c = container[a];
    if (c < 0 || c >= n_items) return ERR_BAD_ARG;
    const double* g = geom + 8 * (int64_t)c;
    const double t = pos[a];
    if (!loc_kind[a]) {
      if (kind[c] != K_SEGMENT) return ERR_BAD_ARG;
      const double* o = side[a] == BACK ? g + 4 : g;
      xy[2 * a] = o[0] + t * g[2];
      xy[2 * a + 1] = o[1] + t * g[3];
    } else if (kind[c] == K_SEGLANE) {
      xy[2 * a] = g[0] + t * g[2];
      xy[2 * a + 1] = g[1] + t * g[3];
    } else if (kind[c] == K_JLANE) {
      const double u = 1.0 - t;
      const double w0 = u * u * u, w1 = 3.0 * u * u * t, w2 = 3.0 * u * t * t, w3 = t * t * t;
      xy[2 * a] = w0 * g[0] + w1 * g[2] + w2 * g[4] + w3 * g[6];
      xy[2 * a + 1] = w0 * g[1] + w1 * g[3] + w2 * g[5] + w3 * g[7];
    } else {
      return ERR_BAD_ARG;
    }
  }
  return ERR_OK;
}

int64_t go_vanished(void* h) { return ((Oracle*)h)->vanished; }
int64_t go_ticks(void* h) { return ((Oracle*)h)->ticks; }

}  // extern "C"
