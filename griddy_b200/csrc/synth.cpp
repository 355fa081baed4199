// synth.cpp — procedural-grid network + actor generator (host side, C ABI).
//
// Restates the reference's examples/procedural-grid.scm:13-95 with the grid size, actor count and
// agenda length as parameters, emitting the flattened structure-of-arrays form that both the CPU
// oracle (oracle/) and the CUDA path (griddy_cuda.cu) consume.  Every id is the item's registration
// index (order of `add!`, world.scm:21-22).
//
//   junctions            procedural-grid.scm:18-25   (array-index-map!, row-major i then j)
//   segments + lanes     procedural-grid.scm:27-63   (dim 'i pass, then dim 'j pass)
//   junction lanes       procedural-grid.scm:65  ->  core.scm:135-148 connect-by-rank!
//                                                ->  core.scm:113-122 connect!  (one add! each)
//   lane lengths         spatial.scm:28-46 (5-chord Bezier), spatial.scm:53-66 (radius),
//                        spatial.scm:75-122 (lane offsets / end points), core.scm:196-206 (curve)
//                        restated in fp64.  The reference computes these through Chickadee vec2
//                        (not vendored); lengths are therefore *inputs* to oracle and device alike
//                         and only need to be self-consistent.
//   actors               procedural-grid.scm:69-95 with a counter-based PRNG (the reference's seed
//                        is `random-source-randomize!`, not reproducible).
//
// Build: g++ -O2 -ffp-contract=off -shared -fPIC  (see __graft_entry__.build()).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct V2 { double x, y; };
inline V2 vadd(V2 a, V2 b) { return {a.x + b.x, a.y + b.y}; }
inline V2 vsub(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }
inline V2 vmul(V2 a, double s) { return {a.x * s, a.y * s}; }
inline double vmag(V2 a) { return std::sqrt(a.x * a.x + a.y * a.y); }

enum { K_JUNCTION = 0, K_SEGMENT = 1, K_SEGLANE = 2, K_JLANE = 3 };
enum { FORW = 0, BACK = 1 };

struct Seg {
  int dim;          // 0 = 'i pass (runs along +j), 1 = 'j pass (runs along +i)
  int i, j;
  int jbeg, jend;   // junction registration ids
  int reg;          // segment registration id
  int cnt[2];       // lanes per direction
  int lane[2][2];   // lane[dir][rank] registration id
};

struct Grid {
  int n;
  std::vector<Seg> segs;
  std::vector<int> di, dj;   // segment ordinal for ('i, i, j) and ('j, i, j), -1 when skipped
  int64_t n_items_before_jlanes;
};

// procedural-grid.scm:27-63: registration order of segments and their lanes.
void build_segments(Grid& g) {
  const int n = g.n;
  g.di.assign((size_t)n * n, -1);
  g.dj.assign((size_t)n * n, -1);
  int64_t reg = (int64_t)n * n;   // junctions were added first
  for (int dim = 0; dim < 2; ++dim)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        if ((dim == 0 && j == n - 1) || (dim == 1 && i == n - 1)) continue;
        Seg s;
        s.dim = dim; s.i = i; s.j = j;
        s.jbeg = i * n + j;
        s.jend = dim == 0 ? i * n + (j + 1) : (i + 1) * n + j;
        s.reg = (int)reg++;
        s.cnt[FORW] = s.cnt[BACK] = 1;
        s.lane[FORW][0] = (int)reg++;   // lane-1
        s.lane[BACK][0] = (int)reg++;   // lane-2
        s.lane[FORW][1] = s.lane[BACK][1] = -1;
        const int dv = dim == 0 ? i : j;
        if (dv % 2 == 0) {
          const int d = (dv % 4 == 0) ? FORW : BACK;   // lane-3 or lane-4
          s.lane[d][1] = (int)reg++;
          s.cnt[d] = 2;
        }
        (dim == 0 ? g.di : g.dj)[(size_t)i * n + j] = (int)g.segs.size();
        g.segs.push_back(s);
      }
  g.n_items_before_jlanes = reg;
}

// (ref junction 'segments): `insert!` conses, so the list is in reverse link! order
// (core.scm:161-165).  link! order is the creation order of the segments.
int junction_segments(const Grid& g, int i, int j, int out[4]) {
  const int n = g.n;
  int c = 0;
  if (i < n - 1) out[c++] = g.dj[(size_t)i * n + j];
  if (i > 0)     out[c++] = g.dj[(size_t)(i - 1) * n + j];
  if (j < n - 1) out[c++] = g.di[(size_t)i * n + j];
  if (j > 0)     out[c++] = g.di[(size_t)i * n + (j - 1)];
  return c;
}

// spatial.scm:53-66: (* 6/5 1/2 max-lane-count 15) is the exact integer 9*m.
double junction_radius(const Grid& g, int i, int j) {
  int segs[4];
  const int c = junction_segments(g, i, j, segs);
  int m = 0;
  for (int k = 0; k < c; ++k) {
    const Seg& s = g.segs[segs[k]];
    const int lc = s.cnt[FORW] + s.cnt[BACK];
    if (lc > m) m = lc;
  }
  return 9.0 * (double)m;
}

struct Geo {
  double spacing, origin;
  const Grid* g;
  std::vector<double> rad;   // junction radius by registration id
  Geo(double sp, double org, const Grid* gr) : spacing(sp), origin(org), g(gr) {
    const int n = gr->n;
    rad.resize((size_t)n * n);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) rad[(size_t)i * n + j] = junction_radius(*gr, i, j);
  }
  V2 jpos(int jreg) const {
    const int n = g->n;
    return {origin + spacing * (double)(jreg / n), origin + spacing * (double)(jreg % n)};
  }
  double jrad(int jreg) const { return rad[(size_t)jreg]; }
  // spatial.scm:160-163
  V2 seg_tangent(const Seg& s) const {
    V2 d = vsub(jpos(s.jend), jpos(s.jbeg));
    double m = vmag(d);
    return {d.x / m, d.y / m};
  }
  // spatial.scm:169-170, math.scm:28-29 (counter-clockwise rotation by pi/2)
  V2 seg_ortho(const Seg& s) const {
    V2 t = seg_tangent(s);
    const double c = std::cos(M_PI / 2.0), sn = std::sin(M_PI / 2.0);
    return {t.x * c - t.y * sn, t.x * sn + t.y * c};
  }
  // spatial.scm:115-122
  V2 seg_pos(const Seg& s, bool end) const {
    V2 t = seg_tangent(s);
    if (!end) return vadd(jpos(s.jbeg), vmul(t, 1.0 * jrad(s.jbeg)));
    return vadd(jpos(s.jend), vmul(t, -1.0 * jrad(s.jend)));
  }
  double seg_len(const Seg& s) const { return vmag(vsub(seg_pos(s, true), seg_pos(s, false))); }
  // spatial.scm:75-102
  V2 lane_offset(const Seg& s, int dir, int rank) const {
    const int from_edge = dir == BACK ? s.cnt[BACK] - rank - 1 : s.cnt[BACK] + rank;
    const int lc = s.cnt[FORW] + s.cnt[BACK];
    V2 o = seg_ortho(s);
    V2 seg_edge = vmul(o, -0.5 * 15.0 * (double)lc);
    V2 lane_edge = vadd(seg_edge, vmul(o, (double)from_edge * 15.0));
    return vadd(lane_edge, vmul(o, 0.5 * 15.0));
  }
  // spatial.scm:104-107
  V2 lane_pos(const Seg& s, int dir, int rank, bool end) const {
    const bool seg_end = dir == FORW ? end : !end;
    return vadd(seg_pos(s, seg_end), lane_offset(s, dir, rank));
  }
  V2 lane_tangent(const Seg& s, int dir) const {
    return vmul(seg_tangent(s), dir == FORW ? 1.0 : -1.0);
  }
};

V2 bezier_at(V2 p0, V2 p1, V2 p2, V2 p3, double t) {
  const double u = 1.0 - t;
  const double a = u * u * u, b = 3.0 * u * u * t, c = 3.0 * u * t * t, d = t * t * t;
  return {a * p0.x + b * p1.x + c * p2.x + d * p3.x, a * p0.y + b * p1.y + c * p2.y + d * p3.y};
}

// spatial.scm:34-46 with the curve of core.scm:196-206.
void jlane_curve(const Geo& geo, int jreg, const Seg& in, int in_dir, int in_rank, const Seg& out,
                 int out_dir, int out_rank, V2 p[4]) {
  // (* 2 radius (/ 20 100)) is an exact rational in the reference; convert once, correctly rounded
  const double offset = (2.0 * geo.jrad(jreg) * 20.0) / 100.0;
  p[0] = geo.lane_pos(in, in_dir, in_rank, true);
  p[3] = geo.lane_pos(out, out_dir, out_rank, false);
  p[1] = vadd(p[0], vmul(geo.lane_tangent(in, in_dir), offset));
  p[2] = vsub(p[3], vmul(geo.lane_tangent(out, out_dir), offset));
}

double jlane_length(const Geo& geo, int jreg, const Seg& in, int in_dir, int in_rank, const Seg& out,
                    int out_dir, int out_rank) {
  V2 cp[4];
  jlane_curve(geo, jreg, in, in_dir, in_rank, out, out_dir, out_rank, cp);
  const V2 p0 = cp[0], p1 = cp[1], p2 = cp[2], p3 = cp[3];
  double acc = 0.0;
  V2 lo = bezier_at(p0, p1, p2, p3, 0.0);
  for (int k = 1; k <= 5; ++k) {
    V2 hi = bezier_at(p0, p1, p2, p3, (double)k / 5.0);
    acc = acc + vmag(vsub(hi, lo));
    lo = hi;
  }
  return acc;
}

// compass heading of a segment lane: 0 = +i, 1 = -i, 2 = +j, 3 = -j
inline int lane_heading(const Seg& s, int dir) {
  return (s.dim == 0 ? 2 : 0) + (dir == FORW ? 0 : 1);
}

// Counter-based PRNG: splitmix64 finaliser over (seed, actor, field).
inline uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
inline uint64_t rnd(uint64_t seed, uint64_t actor, uint64_t field) {
  return mix64(mix64(seed ^ mix64(actor)) + field);
}
inline uint64_t rnd_int(uint64_t seed, uint64_t actor, uint64_t field, uint64_t m) {
  return (rnd(seed, actor, field) >> 11) % m;
}
inline double rnd_real(uint64_t seed, uint64_t actor, uint64_t field) {
  return (double)(rnd(seed, actor, field) >> 11) * (1.0 / 9007199254740992.0);
}

}  // namespace

extern "C" {

// Item counts of an n x n procedural grid.
int synth_grid_sizes(int n, int64_t* n_items, int64_t* n_junctions, int64_t* n_segments,
                     int64_t* n_seg_lanes, int64_t* n_jlanes) {
  if (n < 1) return 1;
  Grid g; g.n = n;
  build_segments(g);
  int64_t sl = 0;
  for (const Seg& s : g.segs) sl += s.cnt[FORW] + s.cnt[BACK];
  int64_t jl = 0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      int segs[4];
      const int c = junction_segments(g, i, j, segs);
      const int jreg = i * n + j;
      for (int a = 0; a < c; ++a)
        for (int b = 0; b < c; ++b) {
          const Seg& in = g.segs[segs[a]];
          const Seg& out = g.segs[segs[b]];
          const int ci = in.cnt[in.jend == jreg ? FORW : BACK];
          const int co = out.cnt[out.jend == jreg ? BACK : FORW];
          jl += ci < co ? ci : co;
        }
    }
  *n_junctions = (int64_t)n * n;
  *n_segments = (int64_t)g.segs.size();
  *n_seg_lanes = sl;
  *n_jlanes = jl;
  *n_items = g.n_items_before_jlanes + jl;
  return 0;
}

// Fill the flattened network.  All per-item arrays have n_items entries; entries that do not apply
// to an item's kind hold -1 (ids) or 0 (numbers).  turn4 has 4 entries per item (headings
// +i, -i, +j, -j): for a segment lane, the junction lane that leads from it onto the segment leaving
// its end junction in that heading (the device-side routing table), else -1.  seg_reg has
// n_segments entries (segment ordinal -> registration id); lane_in gives a junction lane's input lane.
int synth_grid_network(int n, double spacing, double origin, uint8_t* kind, double* lane_len,
                       uint8_t* lane_dir, int32_t* lane_segment, int32_t* lane_out,
                       int32_t* lane_junction, int32_t* seg_outer_forw, int32_t* seg_outer_back,
                       int32_t* lane_in, int32_t* lane_from_j, int32_t* lane_to_j, int32_t* turn4,
                       int32_t* seg_reg) {
  if (n < 1) return 1;
  Grid g; g.n = n;
  build_segments(g);
  Geo geo(spacing, origin, &g);
  int64_t n_items, nj, ns, nsl, njl;
  synth_grid_sizes(n, &n_items, &nj, &ns, &nsl, &njl);
  for (int64_t r = 0; r < n_items; ++r) {
    kind[r] = K_JUNCTION; lane_len[r] = 0.0; lane_dir[r] = 0;
    lane_segment[r] = lane_out[r] = lane_junction[r] = -1;
    seg_outer_forw[r] = seg_outer_back[r] = lane_in[r] = lane_from_j[r] = lane_to_j[r] = -1;
    turn4[4 * r] = turn4[4 * r + 1] = turn4[4 * r + 2] = turn4[4 * r + 3] = -1;
  }
  for (size_t k = 0; k < g.segs.size(); ++k) {
    const Seg& s = g.segs[k];
    seg_reg[k] = s.reg;
    kind[s.reg] = K_SEGMENT;
    // location.scm:16-24 get-outer-lane (both directions are non-empty on this grid)
    seg_outer_forw[s.reg] = s.lane[FORW][s.cnt[FORW] - 1];
    seg_outer_back[s.reg] = s.lane[BACK][s.cnt[BACK] - 1];
    const double len = geo.seg_len(s);
    for (int d = 0; d < 2; ++d)
      for (int r = 0; r < s.cnt[d]; ++r) {
        const int l = s.lane[d][r];
        kind[l] = K_SEGLANE;
        lane_len[l] = len;
        lane_dir[l] = (uint8_t)d;
        lane_segment[l] = s.reg;
        lane_from_j[l] = d == FORW ? s.jbeg : s.jend;
        lane_to_j[l] = d == FORW ? s.jend : s.jbeg;
      }
  }
  // core.scm:135-148 connect-by-rank!, junctions in row-major order (procedural-grid.scm:65)
  int64_t reg = g.n_items_before_jlanes;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      int segs[4];
      const int c = junction_segments(g, i, j, segs);
      const int jreg = i * n + j;
      for (int a = 0; a < c; ++a)
        for (int b = 0; b < c; ++b) {
          const Seg& in = g.segs[segs[a]];
          const Seg& out = g.segs[segs[b]];
          const int din = in.jend == jreg ? FORW : BACK;     // core.scm:69-83 'incoming
          const int dout = out.jend == jreg ? BACK : FORW;   // 'outgoing
          const int m = in.cnt[din] < out.cnt[dout] ? in.cnt[din] : out.cnt[dout];
          for (int k = 0; k < m; ++k) {   // reversed lane lists, zipped (srfi-1 for-each)
            const int rin = in.cnt[din] - 1 - k, rout = out.cnt[dout] - 1 - k;
            const int lin = in.lane[din][rin], lout = out.lane[dout][rout];
            const int jl = (int)reg++;
            kind[jl] = K_JLANE;
            lane_len[jl] = jlane_length(geo, jreg, in, din, rin, out, dout, rout);
            lane_out[jl] = lout;
            lane_in[jl] = lin;
            lane_junction[jl] = jreg;
            turn4[4 * (int64_t)lin + lane_heading(out, dout)] = jl;
          }
        }
    }
  return reg == n_items ? 0 : 2;
}

// Geometry for the drawing read-back (include/griddy_cuda.h, griddy_upload_geometry): 8 doubles per
// item, the fp64 restatement of spatial.scm:104-158 used for the lane lengths above (self-consistent
// only, like them).  Items are visited in the order synth_grid_network registers them.
int synth_grid_geometry(int n, double spacing, double origin, double* geom) {
  if (n < 1) return 1;
  Grid g; g.n = n;
  build_segments(g);
  Geo geo(spacing, origin, &g);
  int64_t n_items, nj, ns, nsl, njl;
  synth_grid_sizes(n, &n_items, &nj, &ns, &nsl, &njl);
  for (int64_t k = 0; k < 8 * n_items; ++k) geom[k] = 0.0;
  auto put = [&](int64_t item, int at, V2 v) { geom[8 * item + at] = v.x; geom[8 * item + at + 1] = v.y; };
  for (const Seg& s : g.segs) {
    const V2 beg = geo.seg_pos(s, false);
    const V2 vec = vsub(geo.seg_pos(s, true), beg);
    // spatial.scm:125-136 with get-width spatial.scm:47-50: v-offset = ortho * (sign * 1/2 * (6/5 * 15 * lanes)
    // * 7/5) — the scalar is the exact rational 63 * lanes / 5, converted once
    const int lc = s.cnt[FORW] + s.cnt[BACK];
    const double w = (63.0 * (double)lc) / 5.0;
    const V2 o = geo.seg_ortho(s);
    put(s.reg, 0, vadd(beg, vmul(o, w)));
    put(s.reg, 2, vec);
    put(s.reg, 4, vadd(beg, vmul(o, -w)));
    for (int d = 0; d < 2; ++d)
      for (int r = 0; r < s.cnt[d]; ++r) {
        const V2 lb = geo.lane_pos(s, d, r, false);
        put(s.lane[d][r], 0, lb);
        put(s.lane[d][r], 2, vsub(geo.lane_pos(s, d, r, true), lb));
      }
  }
  int64_t reg = g.n_items_before_jlanes;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      int segs[4];
      const int c = junction_segments(g, i, j, segs);
      const int jreg = i * n + j;
      for (int a = 0; a < c; ++a)
        for (int b = 0; b < c; ++b) {
          const Seg& in = g.segs[segs[a]];
          const Seg& out = g.segs[segs[b]];
          const int din = in.jend == jreg ? FORW : BACK;
          const int dout = out.jend == jreg ? BACK : FORW;
          const int m = in.cnt[din] < out.cnt[dout] ? in.cnt[din] : out.cnt[dout];
          for (int k = 0; k < m; ++k) {
            V2 cp[4];
            jlane_curve(geo, jreg, in, din, in.cnt[din] - 1 - k, out, dout, out.cnt[dout] - 1 - k, cp);
            for (int q = 0; q < 4; ++q) put(reg, 2 * q, cp[q]);
            ++reg;
          }
        }
    }
  return reg == n_items ? 0 : 2;
}

// procedural-grid.scm:69-95.  agenda_len items per actor alternating sleep-for / travel-to
// (the shipped example is agenda_len = 2).  Actor k of the output is the actor with id ids[k] (ids ==
// null: id k), so that any rank of a partitioned world can generate exactly its own actors: every
// field is a pure function of (seed, actor id).  Arrays: per actor n entries, agenda_off n+1, item
// arrays n*agenda_len.
int synth_actors_ids(int64_t n, const int64_t* ids, uint64_t seed, int agenda_len, int64_t n_segments,
                     const int32_t* seg_reg, uint8_t* loc_kind, int32_t* container, double* pos,
                     uint8_t* side, double* speed, double* speed_dt, int64_t* agenda_off,
                     uint8_t* item_kind, int32_t* item_sleep, int32_t* item_seg, uint8_t* item_side,
                     double* item_pos) {
  if (n_segments < 1 || agenda_len < 0) return 1;
  for (int64_t a = 0; a < n; ++a) {
    const uint64_t ua = (uint64_t)(ids ? ids[a] : a);
    const double s = (double)(30 + (int)rnd_int(seed, ua, 0, 30));
    speed[a] = s;
    speed_dt[a] = s / 15.0;   // exact->inexact of the exact rational s * 1/15
    loc_kind[a] = 0;
    container[a] = seg_reg[rnd_int(seed, ua, 1, (uint64_t)n_segments)];
    side[a] = (uint8_t)(rnd_int(seed, ua, 2, 2) == 0 ? FORW : BACK);
    pos[a] = 0.25 + 0.5 * rnd_real(seed, ua, 3);
    agenda_off[a] = a * agenda_len;
    for (int k = 0; k < agenda_len; ++k) {
      const int64_t it = a * agenda_len + k;
      const uint64_t f = 16 + 8 * (uint64_t)k;
      if (k % 2 == 0) {
        item_kind[it] = 0;
        item_sleep[it] = (int32_t)rnd_int(seed, ua, f, 300);
        item_seg[it] = -1; item_side[it] = 0; item_pos[it] = 0.0;
      } else {
        item_kind[it] = 1;
        item_sleep[it] = 0;
        item_seg[it] = seg_reg[rnd_int(seed, ua, f + 1, (uint64_t)n_segments)];
        item_side[it] = (uint8_t)(rnd_int(seed, ua, f + 2, 2) == 0 ? FORW : BACK);
        item_pos[it] = 0.25 + 0.5 * rnd_real(seed, ua, f + 3);
      }
    }
  }
  agenda_off[n] = n * agenda_len;
  return 0;
}

int synth_actors(int64_t n_actors, uint64_t seed, int agenda_len, int64_t n_segments,
                 const int32_t* seg_reg, uint8_t* loc_kind, int32_t* container, double* pos,
                 uint8_t* side, double* speed, double* speed_dt, int64_t* agenda_off,
                 uint8_t* item_kind, int32_t* item_sleep, int32_t* item_seg, uint8_t* item_side,
                 double* item_pos) {
  return synth_actors_ids(n_actors, nullptr, seed, agenda_len, n_segments, seg_reg, loc_kind, container, pos,
                          side, speed, speed_dt, agenda_off, item_kind, item_sleep, item_seg, item_side, item_pos);
}

// The ids in [0, n_actors) whose start segment belongs to `rank` (owner by item): what one rank of a
// partitioned world uploads.  ids_out has room for n_actors entries; returns the count in *n_out.
int synth_actor_ids_of_rank(int64_t n_actors, uint64_t seed, int64_t n_segments, const int32_t* seg_reg,
                            const int32_t* owner, int32_t rank, int64_t* ids_out, int64_t* n_out) {
  if (n_segments < 1) return 1;
  int64_t k = 0;
  for (int64_t a = 0; a < n_actors; ++a)
    if (owner[seg_reg[rnd_int(seed, (uint64_t)a, 1, (uint64_t)n_segments)]] == rank) ids_out[k++] = a;
  *n_out = k;
  return 0;
}

// The junction an item hangs off (row-major junction id on the grid): a junction itself, a junction
// lane's junction, the beg junction of a segment, and for a segment lane the junction it leaves
// (by_segment = 0, locality) or its segment's beg junction (by_segment = 1, ownership).
static inline int64_t item_junction(int64_t r, const uint8_t* kind, const int32_t* lane_from_j,
                                    const int32_t* lane_junction, const int32_t* seg_outer_forw,
                                    const int32_t* lane_segment, int by_segment) {
  switch (kind[r]) {
    case K_JUNCTION: return r;
    case K_SEGMENT: return lane_from_j[seg_outer_forw[r]];
    case K_SEGLANE: return by_segment ? lane_from_j[seg_outer_forw[lane_segment[r]]] : lane_from_j[r];
    case K_JLANE: return lane_junction[r];
    default: return 0;
  }
}

// Locality key per item (griddy_set_locality_keys): items grouped by the junction they hang off,
// stable within a group.  Counting sort, O(n_items).
int synth_locality_keys(int64_t n_items, int64_t n_junctions, const uint8_t* kind, const int32_t* lane_from_j,
                        const int32_t* lane_junction, const int32_t* seg_outer_forw,
                        const int32_t* lane_segment, uint32_t* key) {
  std::vector<int64_t> start((size_t)n_junctions + 1, 0);
  for (int64_t r = 0; r < n_items; ++r)
    ++start[(size_t)item_junction(r, kind, lane_from_j, lane_junction, seg_outer_forw, lane_segment, 0) + 1];
  for (int64_t j = 0; j < n_junctions; ++j) start[(size_t)j + 1] += start[(size_t)j];
  for (int64_t r = 0; r < n_items; ++r)
    key[r] = (uint32_t)start[(size_t)item_junction(r, kind, lane_from_j, lane_junction, seg_outer_forw, lane_segment, 0)]++;
  return 0;
}

// Owner rank per item: world * strips_per_rank horizontal strips of junction rows, dealt to the ranks
// in turn (with more than one strip per rank every rank gets its share of the busy middle of the grid
// and of its quiet edges).  A junction owns its junction lanes; a segment and all its lanes belong to
// the junction the segment begins at.
int synth_grid_owner(int64_t n_items, int n, int world, int strips_per_rank, const uint8_t* kind,
                     const int32_t* lane_from_j, const int32_t* lane_junction, const int32_t* seg_outer_forw,
                     const int32_t* lane_segment, int32_t* owner) {
  if (n < 1 || world < 1 || strips_per_rank < 1) return 1;
  const int64_t strips = (int64_t)world * strips_per_rank;
  for (int64_t r = 0; r < n_items; ++r) {
    const int64_t j = item_junction(r, kind, lane_from_j, lane_junction, seg_outer_forw, lane_segment, 1);
    owner[r] = (int32_t)(((j / n) * strips / n) % world);   // strips are dealt to the ranks in turn
  }
  return 0;
}

}  // extern "C"
