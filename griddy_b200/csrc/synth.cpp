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
#include <algorithm>
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

// procedural-grid.scm:27-63: the registration order of segments and their lanes is regular enough for
// closed forms, so that any part of the grid can be generated without the rest (a rank of a partitioned
// world builds only its own rows).  Junctions come first (row-major), then the 'i pass (for i, for j:
// the segment from (i,j) to (i,j+1)), then the 'j pass (for i, for j: from (i,j) to (i+1,j)); a segment
// registers itself, lane-1 (forw), lane-2 (back) and, when its dimension value dv (i in the 'i pass, j in
// the 'j pass) is even, a third lane: forw if dv % 4 == 0, else back.
struct Grid {
  int n;
  int64_t D0, D1;                  // first id of the 'i pass and of the 'j pass
  int64_t n_items_before_jlanes;
  std::vector<int64_t> jl_row;     // first junction-lane id of every junction row (n + 1 entries), see count_jlanes()
  explicit Grid(int n_) : n(n_) {
    D0 = (int64_t)n * n;
    D1 = D0 + (int64_t)(n - 1) * (3 * (int64_t)n + (n + 1) / 2);
    n_items_before_jlanes = D1 + (int64_t)(n - 1) * (3 * (int64_t)n + (n + 1) / 2);
  }
  bool has(int dim, int i, int j) const {
    return i >= 0 && j >= 0 && i < n && j < n && (dim == 0 ? j < n - 1 : i < n - 1);
  }
  int64_t n_segments() const { return 2 * (int64_t)n * (n - 1); }
  int64_t ordinal(int dim, int i, int j) const {   // creation order of the segments
    return dim == 0 ? (int64_t)i * (n - 1) + j : (int64_t)n * (n - 1) + (int64_t)i * n + j;
  }
  void unordinal(int64_t k, int* dim, int* i, int* j) const {
    const int64_t h = (int64_t)n * (n - 1);
    if (k < h) { *dim = 0; *i = (int)(k / (n - 1)); *j = (int)(k % (n - 1)); }
    else { *dim = 1; *i = (int)((k - h) / n); *j = (int)((k - h) % n); }
  }
  int64_t seg_reg(int dim, int i, int j) const {
    if (dim == 0) return D0 + (int64_t)(n - 1) * (3 * (int64_t)i + (i + 1) / 2) + (int64_t)j * (3 + (i % 2 == 0 ? 1 : 0));
    return D1 + (int64_t)i * (3 * (int64_t)n + (n + 1) / 2) + 3 * (int64_t)j + (j + 1) / 2;
  }
  // first id of the 'i-pass / 'j-pass items of junction row i (i may be n: the end)
  int64_t d0_row(int i) const { return D0 + (int64_t)(n - 1) * (3 * (int64_t)i + (i + 1) / 2); }
  int64_t d1_row(int i) const { return D1 + (int64_t)std::min(i, n - 1) * (3 * (int64_t)n + (n + 1) / 2); }
  Seg seg(int dim, int i, int j) const {
    Seg s;
    s.dim = dim; s.i = i; s.j = j;
    s.jbeg = i * n + j;
    s.jend = dim == 0 ? i * n + (j + 1) : (i + 1) * n + j;
    s.reg = (int)seg_reg(dim, i, j);
    s.cnt[FORW] = s.cnt[BACK] = 1;
    s.lane[FORW][0] = s.reg + 1;   // lane-1
    s.lane[BACK][0] = s.reg + 2;   // lane-2
    s.lane[FORW][1] = s.lane[BACK][1] = -1;
    const int dv = dim == 0 ? i : j;
    if (dv % 2 == 0) {
      const int d = (dv % 4 == 0) ? FORW : BACK;   // lane-3 or lane-4
      s.lane[d][1] = s.reg + 3;
      s.cnt[d] = 2;
    }
    return s;
  }
  int seg_items(int dim, int i, int j) const { return 3 + ((dim == 0 ? i : j) % 2 == 0 ? 1 : 0); }
  void count_jlanes();
};

// (ref junction 'segments): `insert!` conses, so the list is in reverse link! order
// (core.scm:161-165).  link! order is the creation order of the segments.
int junction_segments(const Grid& g, int i, int j, Seg out[4]) {
  const int n = g.n;
  int c = 0;
  if (i < n - 1) out[c++] = g.seg(1, i, j);
  if (i > 0)     out[c++] = g.seg(1, i - 1, j);
  if (j < n - 1) out[c++] = g.seg(0, i, j);
  if (j > 0)     out[c++] = g.seg(0, i, j - 1);
  return c;
}

// core.scm:135-148 connect-by-rank!: junction lanes of junction (i, j), all pairs (incoming, outgoing) of its
// segments, min(lane counts) each.
int junction_jlanes(const Grid& g, int i, int j) {
  Seg segs[4];
  const int c = junction_segments(g, i, j, segs);
  const int jreg = i * g.n + j;
  int jl = 0;
  for (int a = 0; a < c; ++a)
    for (int b = 0; b < c; ++b) {
      const int ci = segs[a].cnt[segs[a].jend == jreg ? FORW : BACK];
      const int co = segs[b].cnt[segs[b].jend == jreg ? BACK : FORW];
      jl += ci < co ? ci : co;
    }
  return jl;
}

void Grid::count_jlanes() {
  if (!jl_row.empty()) return;
  jl_row.assign((size_t)n + 1, 0);
  jl_row[0] = n_items_before_jlanes;
  for (int i = 0; i < n; ++i) {
    int64_t row = 0;
    for (int j = 0; j < n; ++j) row += junction_jlanes(*this, i, j);
    jl_row[(size_t)i + 1] = jl_row[(size_t)i] + row;
  }
}

// spatial.scm:53-66: (* 6/5 1/2 max-lane-count 15) is the exact integer 9*m.
double junction_radius(const Grid& g, int i, int j) {
  Seg segs[4];
  const int c = junction_segments(g, i, j, segs);
  int m = 0;
  for (int k = 0; k < c; ++k) {
    const Seg& s = segs[k];
    const int lc = s.cnt[FORW] + s.cnt[BACK];
    if (lc > m) m = lc;
  }
  return 9.0 * (double)m;
}

struct Geo {
  double spacing, origin;
  const Grid* g;
  int rad_row0 = 0, rad_row1 = 0;   // junction rows whose radii are cached (a tile's rows)
  std::vector<double> rad;
  Geo(double sp, double org, const Grid* gr, int row0, int row1) : spacing(sp), origin(org), g(gr) {
    rad_row0 = std::max(row0, 0);
    rad_row1 = std::min(row1, gr->n);
    const int n = gr->n;
    rad.resize((size_t)std::max(rad_row1 - rad_row0, 0) * n);
    for (int i = rad_row0; i < rad_row1; ++i)
      for (int j = 0; j < n; ++j) rad[(size_t)(i - rad_row0) * n + j] = junction_radius(*gr, i, j);
  }
  V2 jpos(int jreg) const {
    const int n = g->n;
    return {origin + spacing * (double)(jreg / n), origin + spacing * (double)(jreg % n)};
  }
  double jrad(int jreg) const {
    const int i = jreg / g->n;
    if (i >= rad_row0 && i < rad_row1) return rad[(size_t)(jreg - rad_row0 * g->n)];
    return junction_radius(*g, i, jreg % g->n);
  }
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

// ---- tiles ---------------------------------------------------------------------------------------------
// A tile is a set of junction-row intervals [row_beg[k], row_end[k]) (ascending, disjoint).  Its items are
// everything that hangs off those rows: the junctions, the segments that BEGIN there with their lanes, and the
// junctions' junction lanes — four id ranges per interval, because each of the four registration passes
// visits the rows in order.  Compact arrays list the ranges in ascending id order: all junction ranges, then
// the 'i-pass ranges, the 'j-pass ranges, the junction-lane ranges.
struct Tile {
  std::vector<int> rb, re;
  std::vector<int64_t> beg, end, off;   // id ranges, ascending, and the compact index of each range's first item
  int64_t n_local = 0;
  Tile(Grid& g, int n_iv, const int32_t* row_beg, const int32_t* row_end) {
    g.count_jlanes();
    for (int k = 0; k < n_iv; ++k) { rb.push_back(row_beg[k]); re.push_back(row_end[k]); }
    for (int pass = 0; pass < 4; ++pass)
      for (int k = 0; k < n_iv; ++k) {
        int64_t b, e;
        if (pass == 0) { b = (int64_t)rb[k] * g.n; e = (int64_t)re[k] * g.n; }
        else if (pass == 1) { b = g.d0_row(rb[k]); e = g.d0_row(re[k]); }
        else if (pass == 2) { b = g.d1_row(rb[k]); e = g.d1_row(re[k]); }
        else { b = g.jl_row[(size_t)rb[k]]; e = g.jl_row[(size_t)re[k]]; }
        beg.push_back(b); end.push_back(e); off.push_back(n_local);
        n_local += e - b;
      }
  }
  int64_t local(int64_t gid) const {
    for (size_t k = 0; k < beg.size(); ++k)
      if (gid >= beg[k] && gid < end[k]) return off[k] + (gid - beg[k]);
    return -1;
  }
};

bool tile_rows_ok(int n, int n_iv, const int32_t* row_beg, const int32_t* row_end) {
  int prev = 0;
  for (int k = 0; k < n_iv; ++k) {
    if (row_beg[k] < prev || row_end[k] <= row_beg[k] || row_end[k] > n) return false;
    prev = row_end[k];
  }
  return n_iv >= 1;
}

// The strip of junction rows that row i belongs to, and its owner: world * strips_per_rank horizontal strips
// dealt to the ranks in turn.
inline int row_owner(int i, int n, int world, int strips_per_rank) {
  return (int)((((int64_t)i * world * strips_per_rank) / n) % world);
}

// location.scm:16-24 get-outer-lane on this grid (both directions are non-empty)
inline int outer_lane(const Seg& s, int side) { return s.lane[side][s.cnt[side] - 1]; }

}  // namespace

extern "C" {

// Item counts of an n x n procedural grid.
int synth_grid_sizes(int n, int64_t* n_items, int64_t* n_junctions, int64_t* n_segments,
                     int64_t* n_seg_lanes, int64_t* n_jlanes) {
  if (n < 1) return 1;
  Grid g(n);
  g.count_jlanes();
  int64_t sl = 0;
  for (int i = 0; i < n; ++i) sl += (int64_t)(n - 1) * (g.seg_items(0, i, 0) - 1);
  for (int j = 0; j < n; ++j) sl += (int64_t)(n - 1) * (g.seg_items(1, 0, j) - 1);
  *n_junctions = (int64_t)n * n;
  *n_segments = g.n_segments();
  *n_seg_lanes = sl;
  *n_jlanes = g.jl_row[(size_t)n] - g.n_items_before_jlanes;
  *n_items = g.jl_row[(size_t)n];
  return 0;
}

// The id ranges of a tile (see Tile above): 4 * n_iv ranges, ascending; *n_local = items in them.
int synth_tile_ranges(int n, int n_iv, const int32_t* row_beg, const int32_t* row_end, int64_t* range_beg,
                      int64_t* range_end, int64_t* n_local) {
  if (n < 1 || !tile_rows_ok(n, n_iv, row_beg, row_end)) return 1;
  Grid g(n);
  const Tile t(g, n_iv, row_beg, row_end);
  for (size_t k = 0; k < t.beg.size(); ++k) { range_beg[k] = t.beg[k]; range_end[k] = t.end[k]; }
  *n_local = t.n_local;
  return 0;
}

// Fill the flattened network of a tile (the whole grid: one interval [0, n)).  All per-item arrays are compact
// (one entry per item of the tile, in range order); the ids inside them are global.  Entries that do not
// apply to an item's kind hold -1 (ids) or 0 (numbers).  turn4 has 4 entries per item (headings +i, -i, +j,
// -j): for a segment lane, the junction lane that leads from it onto the segment leaving its end junction in
// that heading (the device-side routing table), else -1 — filled for the lanes whose end junction lies in the
// tile.  lane_in gives a junction lane's input lane.  hang_j: the junction an item hangs off (a junction
// itself, a junction lane's junction, a segment lane: the junction it leaves, a segment: the junction its forw
// lanes leave) — the locality group.  geom (may be null): 8 doubles per item, see synth_grid_geometry.
int synth_tile_network(int n, double spacing, double origin, int n_iv, const int32_t* row_beg, const int32_t* row_end,
                       uint8_t* kind, double* lane_len, uint8_t* lane_dir, int32_t* lane_segment, int32_t* lane_out,
                       int32_t* lane_junction, int32_t* seg_outer_forw, int32_t* seg_outer_back,
                       int32_t* lane_in, int32_t* lane_from_j, int32_t* lane_to_j, int32_t* turn4,
                       int32_t* hang_j, double* geom) {
  if (n < 1 || !tile_rows_ok(n, n_iv, row_beg, row_end)) return 1;
  Grid g(n);
  const Tile t(g, n_iv, row_beg, row_end);
  Geo geo(spacing, origin, &g, row_beg[0] - 1, row_end[n_iv - 1] + 1);
  for (int64_t r = 0; r < t.n_local; ++r) {
    kind[r] = K_JUNCTION; lane_len[r] = 0.0; lane_dir[r] = 0;
    lane_segment[r] = lane_out[r] = lane_junction[r] = -1;
    seg_outer_forw[r] = seg_outer_back[r] = lane_in[r] = lane_from_j[r] = lane_to_j[r] = -1;
    turn4[4 * r] = turn4[4 * r + 1] = turn4[4 * r + 2] = turn4[4 * r + 3] = -1;
    hang_j[r] = -1;
  }
  if (geom)
    for (int64_t k = 0; k < 8 * t.n_local; ++k) geom[k] = 0.0;
  auto put = [&](int64_t l, int at, V2 v) { geom[8 * l + at] = v.x; geom[8 * l + at + 1] = v.y; };
  for (int k = 0; k < n_iv; ++k)
    for (int i = t.rb[k]; i < t.re[k]; ++i) {
      for (int j = 0; j < n; ++j) hang_j[t.local((int64_t)i * n + j)] = i * n + j;
      for (int dim = 0; dim < 2; ++dim)
        for (int j = 0; j < n; ++j) {
          if (!g.has(dim, i, j)) continue;
          const Seg s = g.seg(dim, i, j);
          const int64_t ls = t.local(s.reg);
          kind[ls] = K_SEGMENT;
          seg_outer_forw[ls] = outer_lane(s, FORW);
          seg_outer_back[ls] = outer_lane(s, BACK);
          hang_j[ls] = s.jbeg;
          const double len = geo.seg_len(s);
          if (geom) {
            const V2 beg = geo.seg_pos(s, false);
            const V2 vec = vsub(geo.seg_pos(s, true), beg);
            // spatial.scm:125-136 with get-width spatial.scm:47-50: v-offset = ortho * (sign * 1/2 * (6/5 * 15 * lanes)
            // * 7/5) — the scalar is the exact rational 63 * lanes / 5, converted once
            const int lc = s.cnt[FORW] + s.cnt[BACK];
            const double w = (63.0 * (double)lc) / 5.0;
            const V2 o = geo.seg_ortho(s);
            put(ls, 0, vadd(beg, vmul(o, w)));
            put(ls, 2, vec);
            put(ls, 4, vadd(beg, vmul(o, -w)));
          }
          for (int d = 0; d < 2; ++d)
            for (int r = 0; r < s.cnt[d]; ++r) {
              const int64_t l = ls + (s.lane[d][r] - s.reg);
              kind[l] = K_SEGLANE;
              lane_len[l] = len;
              lane_dir[l] = (uint8_t)d;
              lane_segment[l] = s.reg;
              lane_from_j[l] = d == FORW ? s.jbeg : s.jend;
              lane_to_j[l] = d == FORW ? s.jend : s.jbeg;
              hang_j[l] = lane_from_j[l];
              if (geom) {
                const V2 lb = geo.lane_pos(s, d, r, false);
                put(l, 0, lb);
                put(l, 2, vsub(geo.lane_pos(s, d, r, true), lb));
              }
            }
        }
    }
  // core.scm:135-148 connect-by-rank!, junctions in row-major order (procedural-grid.scm:65)
  for (int k = 0; k < n_iv; ++k) {
    int64_t reg = g.jl_row[(size_t)t.rb[k]];
    for (int i = t.rb[k]; i < t.re[k]; ++i)
      for (int j = 0; j < n; ++j) {
        Seg segs[4];
        const int c = junction_segments(g, i, j, segs);
        const int jreg = i * n + j;
        for (int a = 0; a < c; ++a)
          for (int b = 0; b < c; ++b) {
            const Seg& in = segs[a];
            const Seg& out = segs[b];
            const int din = in.jend == jreg ? FORW : BACK;     // core.scm:69-83 'incoming
            const int dout = out.jend == jreg ? BACK : FORW;   // 'outgoing
            const int m = in.cnt[din] < out.cnt[dout] ? in.cnt[din] : out.cnt[dout];
            for (int q = 0; q < m; ++q) {   // reversed lane lists, zipped (srfi-1 for-each)
              const int rin = in.cnt[din] - 1 - q, rout = out.cnt[dout] - 1 - q;
              const int lin = in.lane[din][rin], lout = out.lane[dout][rout];
              const int64_t jl = t.local(reg);
              kind[jl] = K_JLANE;
              lane_len[jl] = jlane_length(geo, jreg, in, din, rin, out, dout, rout);
              lane_out[jl] = lout;
              lane_in[jl] = lin;
              lane_junction[jl] = jreg;
              hang_j[jl] = jreg;
              const int64_t llin = t.local(lin);   // the feeding lane belongs to a segment that may begin outside the tile
              if (llin >= 0) turn4[4 * llin + lane_heading(out, dout)] = (int32_t)reg;
              if (geom) {
                V2 cp[4];
                jlane_curve(geo, jreg, in, din, rin, out, dout, rout, cp);
                for (int w = 0; w < 4; ++w) put(jl, 2 * w, cp[w]);
              }
              ++reg;
            }
          }
      }
    if (reg != g.jl_row[(size_t)t.re[k]]) return 2;
  }
  return 0;
}

// The whole grid: arrays of n_items entries; seg_reg has n_segments entries (segment ordinal -> registration id).
int synth_grid_network(int n, double spacing, double origin, uint8_t* kind, double* lane_len,
                       uint8_t* lane_dir, int32_t* lane_segment, int32_t* lane_out,
                       int32_t* lane_junction, int32_t* seg_outer_forw, int32_t* seg_outer_back,
                       int32_t* lane_in, int32_t* lane_from_j, int32_t* lane_to_j, int32_t* turn4,
                       int32_t* seg_reg) {
  if (n < 1) return 1;
  const int32_t rb = 0, re = n;
  Grid g(n);
  g.count_jlanes();
  std::vector<int32_t> hang((size_t)g.jl_row[(size_t)n]);
  const int rc = synth_tile_network(n, spacing, origin, 1, &rb, &re, kind, lane_len, lane_dir, lane_segment, lane_out,
                                    lane_junction, seg_outer_forw, seg_outer_back, lane_in, lane_from_j, lane_to_j,
                                    turn4, hang.data(), nullptr);
  if (rc) return rc;
  for (int64_t k = 0; k < g.n_segments(); ++k) {
    int dim, i, j;
    g.unordinal(k, &dim, &i, &j);
    seg_reg[k] = (int32_t)g.seg_reg(dim, i, j);
  }
  return 0;
}

// Geometry for the drawing read-back (include/griddy_cuda.h, griddy_upload_geometry): 8 doubles per
// item, the fp64 restatement of spatial.scm:104-158 used for the lane lengths above (self-consistent
// only, like them).
int synth_grid_geometry(int n, double spacing, double origin, double* geom) {
  if (n < 1) return 1;
  int64_t ni, nj, ns, nsl, njl;
  synth_grid_sizes(n, &ni, &nj, &ns, &nsl, &njl);
  std::vector<uint8_t> u8((size_t)ni * 2);
  std::vector<double> len((size_t)ni);
  std::vector<int32_t> i32((size_t)ni * 13);
  const int32_t rb = 0, re = n;
  int32_t* q = i32.data();
  return synth_tile_network(n, spacing, origin, 1, &rb, &re, u8.data(), len.data(), u8.data() + ni, q, q + ni, q + 2 * ni,
                            q + 3 * ni, q + 4 * ni, q + 5 * ni, q + 6 * ni, q + 7 * ni, q + 8 * ni, q + 12 * ni, geom);
}

// procedural-grid.scm:69-95.  agenda_len items per actor alternating sleep-for / travel-to
// (the shipped example is agenda_len = 2).  Actor k of the output is the actor with id ids[k] (ids ==
// null: id k), so that any rank of a partitioned world can generate exactly its own actors: every
// field is a pure function of (seed, actor id).  Arrays: per actor n entries, agenda_off n+1, item
// arrays n*agenda_len.  seg_reg (segment ordinal -> registration id) may be null when grid_n is given: the ids
// then come from the closed forms.  dest_* (may be null): every travel-to item's destination as begin-route$
// resolves it (griddy_set_agenda_destinations): get-outer-lane of (segment, side), that lane's direction and
// the junctions it leaves / reaches.
int synth_actors_ids(int64_t n, const int64_t* ids, uint64_t seed, int agenda_len, int64_t n_segments,
                     const int32_t* seg_reg, uint8_t* loc_kind, int32_t* container, double* pos,
                     uint8_t* side, double* speed, double* speed_dt, int64_t* agenda_off,
                     uint8_t* item_kind, int32_t* item_sleep, int32_t* item_seg, uint8_t* item_side,
                     double* item_pos, int grid_n, int32_t* dest_lane, uint8_t* dest_dir, int32_t* dest_from_j,
                     int32_t* dest_to_j) {
  if (n_segments < 1 || agenda_len < 0 || (!seg_reg && grid_n < 1) || (dest_lane && grid_n < 1)) return 1;
  const Grid g(grid_n > 0 ? grid_n : 2);
  auto seg_of = [&](uint64_t ordinal) {
    int dim, i, j;
    g.unordinal((int64_t)ordinal, &dim, &i, &j);
    return g.seg(dim, i, j);
  };
  for (int64_t a = 0; a < n; ++a) {
    const uint64_t ua = (uint64_t)(ids ? ids[a] : a);
    const double s = (double)(30 + (int)rnd_int(seed, ua, 0, 30));
    speed[a] = s;
    speed_dt[a] = s / 15.0;   // exact->inexact of the exact rational s * 1/15
    loc_kind[a] = 0;
    const uint64_t k0 = rnd_int(seed, ua, 1, (uint64_t)n_segments);
    container[a] = seg_reg ? seg_reg[k0] : seg_of(k0).reg;
    side[a] = (uint8_t)(rnd_int(seed, ua, 2, 2) == 0 ? FORW : BACK);
    pos[a] = 0.25 + 0.5 * rnd_real(seed, ua, 3);
    agenda_off[a] = a * agenda_len;
    for (int k = 0; k < agenda_len; ++k) {
      const int64_t it = a * agenda_len + k;
      const uint64_t f = 16 + 8 * (uint64_t)k;
      if (dest_lane) { dest_lane[it] = dest_from_j[it] = dest_to_j[it] = -1; dest_dir[it] = 0; }
      if (k % 2 == 0) {
        item_kind[it] = 0;
        item_sleep[it] = (int32_t)rnd_int(seed, ua, f, 300);
        item_seg[it] = -1; item_side[it] = 0; item_pos[it] = 0.0;
      } else {
        const uint64_t kd = rnd_int(seed, ua, f + 1, (uint64_t)n_segments);
        item_kind[it] = 1;
        item_sleep[it] = 0;
        item_seg[it] = seg_reg ? seg_reg[kd] : seg_of(kd).reg;
        item_side[it] = (uint8_t)(rnd_int(seed, ua, f + 2, 2) == 0 ? FORW : BACK);
        item_pos[it] = 0.25 + 0.5 * rnd_real(seed, ua, f + 3);
        if (dest_lane) {
          const Seg sd = seg_of(kd);
          const int d = item_side[it];
          dest_lane[it] = outer_lane(sd, d);
          dest_dir[it] = (uint8_t)d;   // a side's outer lane runs in that side's direction
          dest_from_j[it] = d == FORW ? sd.jbeg : sd.jend;
          dest_to_j[it] = d == FORW ? sd.jend : sd.jbeg;
        }
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
                          side, speed, speed_dt, agenda_off, item_kind, item_sleep, item_seg, item_side, item_pos,
                          0, nullptr, nullptr, nullptr, nullptr);
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

// The same from the closed forms (no global arrays): a segment belongs to the strip of the junction row it begins at.
int synth_actor_ids_of_rank_grid(int64_t n_actors, uint64_t seed, int grid_n, int world, int strips_per_rank, int32_t rank,
                                 int64_t* ids_out, int64_t* n_out) {
  if (grid_n < 2 || world < 1 || strips_per_rank < 1) return 1;
  const Grid g(grid_n);
  const int64_t n_segments = g.n_segments();
  int64_t k = 0;
  for (int64_t a = 0; a < n_actors; ++a) {
    int dim, i, j;
    g.unordinal((int64_t)rnd_int(seed, (uint64_t)a, 1, (uint64_t)n_segments), &dim, &i, &j);
    if (row_owner(i, grid_n, world, strips_per_rank) == rank) ids_out[k++] = a;
  }
  *n_out = k;
  return 0;
}

// Locality key per item (griddy_set_locality_keys): items grouped by the junction they hang off (hang_j, ids below
// n_junctions), stable within a group.  Counting sort, O(n_items + n_junctions).
int synth_keys_from_hang(int64_t n_items, int64_t n_junctions, const int32_t* hang_j, uint32_t* key) {
  std::vector<int64_t> start((size_t)n_junctions + 1, 0);
  for (int64_t r = 0; r < n_items; ++r) {
    if (hang_j[r] < 0 || hang_j[r] >= n_junctions) return 1;
    ++start[(size_t)hang_j[r] + 1];
  }
  for (int64_t j = 0; j < n_junctions; ++j) start[(size_t)j + 1] += start[(size_t)j];
  for (int64_t r = 0; r < n_items; ++r) key[r] = (uint32_t)start[(size_t)hang_j[r]]++;
  return 0;
}

// The junction an item hangs off (row-major junction id on the grid): a junction itself, a junction
// lane's junction, the beg junction of a segment, and for a segment lane the junction it leaves
// (by_segment = 0, locality) or its segment's beg junction (by_segment = 1, ownership).  Arrays over ALL items.
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

int synth_locality_keys(int64_t n_items, int64_t n_junctions, const uint8_t* kind, const int32_t* lane_from_j,
                        const int32_t* lane_junction, const int32_t* seg_outer_forw,
                        const int32_t* lane_segment, uint32_t* key) {
  std::vector<int32_t> hang((size_t)n_items);
  for (int64_t r = 0; r < n_items; ++r)
    hang[(size_t)r] = (int32_t)item_junction(r, kind, lane_from_j, lane_junction, seg_outer_forw, lane_segment, 0);
  return synth_keys_from_hang(n_items, n_junctions, hang.data(), key);
}

// Owner rank per item: world * strips_per_rank horizontal strips of junction rows, dealt to the ranks
// in turn (with more than one strip per rank every rank gets its share of the busy middle of the grid
// and of its quiet edges).  A junction owns its junction lanes; a segment and all its lanes belong to
// the junction the segment begins at.  Arrays over ALL items.
int synth_grid_owner(int64_t n_items, int n, int world, int strips_per_rank, const uint8_t* kind,
                     const int32_t* lane_from_j, const int32_t* lane_junction, const int32_t* seg_outer_forw,
                     const int32_t* lane_segment, int32_t* owner) {
  if (n < 1 || world < 1 || strips_per_rank < 1) return 1;
  for (int64_t r = 0; r < n_items; ++r) {
    const int64_t j = item_junction(r, kind, lane_from_j, lane_junction, seg_outer_forw, lane_segment, 1);
    owner[r] = row_owner((int)(j / n), n, world, strips_per_rank);
  }
  return 0;
}

// The same for the items of a tile (compact): an item of the tile hangs off one of its rows, and that row's
// strip owns it.  Also the junction rows [a, b) a rank owns, as ascending intervals (at most strips_per_rank).
int synth_tile_owner(int n, int n_iv, const int32_t* row_beg, const int32_t* row_end, int world, int strips_per_rank,
                     int32_t* owner) {
  if (n < 1 || world < 1 || strips_per_rank < 1 || !tile_rows_ok(n, n_iv, row_beg, row_end)) return 1;
  Grid g(n);
  const Tile t(g, n_iv, row_beg, row_end);
  for (int pass = 0; pass < 4; ++pass)
    for (int k = 0; k < n_iv; ++k)
      for (int i = t.rb[k]; i < t.re[k]; ++i) {
        int64_t b, e;
        if (pass == 0) { b = (int64_t)i * n; e = b + n; }
        else if (pass == 1) { b = g.d0_row(i); e = g.d0_row(i + 1); }
        else if (pass == 2) { b = g.d1_row(i); e = g.d1_row(i + 1); }
        else { b = g.jl_row[(size_t)i]; e = g.jl_row[(size_t)i + 1]; }
        const int32_t o = row_owner(i, n, world, strips_per_rank);
        for (int64_t id = b; id < e; ++id) owner[t.local(id)] = o;
      }
  return 0;
}

int synth_rank_rows(int n, int world, int strips_per_rank, int rank, int32_t* row_beg, int32_t* row_end, int32_t* n_iv) {
  if (n < 1 || world < 1 || strips_per_rank < 1 || rank < 0 || rank >= world) return 1;
  int k = 0;
  for (int i = 0; i < n; ++i) {
    if (row_owner(i, n, world, strips_per_rank) != rank) continue;
    if (k > 0 && row_end[k - 1] == i) row_end[k - 1] = i + 1;
    else { row_beg[k] = i; row_end[k] = i + 1; ++k; }
  }
  *n_iv = k;
  return 0;
}

}  // extern "C"
