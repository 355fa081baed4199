// routes.cuh — part of griddy_cuda.cu (one translation unit).  find-route on the device: A* over lanes, one search per travel-to item.
#pragma once

namespace griddy {

// ------------------------------------------------------------------------------------------------
// find-route (route.scm:17-51): A* from the lane a route starts on to the destination lane.
//   neighbors   a segment lane -> the junction lanes it feeds (get-junction-lanes, core.scm:97-104), a junction
//               lane -> the lane it leads onto (core.scm:106-111): `nbr`, in the reference's list order
//   cost        of leaving a lane: its segment's length for a segment lane, 0 for a junction lane (route.scm:26-34)
//   distance    l2 of the two lanes' midpoints (route.scm:37-40, math.scm:25-26): vec2-, then vec2-magnitude
// The search itself is Chickadee's `a*` (chickadee data path-finding), which is not vendored; what is restated is
// the algorithm tests/golden/minischeme.py substitutes for it and the fixtures' routes were recorded with: textbook
// A* without a closed set, a priority queue ordered by (cost so far + distance to the goal, order of insertion), a
// neighbour relaxed when unseen or strictly cheaper.  Every pop is decided by that total order, so any correct heap
// yields the same routes; the arithmetic is issued one rounding at a time like the evaluator's.
// One thread per search: a route is computed once per travel-to, the searches of all agenda items run side by side.
struct HeapEnt {
  double f;
  int32_t order, node;
};

__device__ __forceinline__ bool heap_less(const HeapEnt& a, const HeapEnt& b) {
  return a.f < b.f || (a.f == b.f && a.order < b.order);
}

__global__ void __launch_bounds__(32) k_find_routes(int n_search, const int32_t* __restrict__ start, const int32_t* __restrict__ goal,
                                                     int64_t n_items, const int64_t* __restrict__ nbr_off,
                                                     const int32_t* __restrict__ nbr, const double* __restrict__ cost_out,
                                                     const double2* __restrict__ mid, double* __restrict__ best_all,
                                                     int32_t* __restrict__ came_all, HeapEnt* __restrict__ heap_all,
                                                     int64_t heap_cap, int32_t* __restrict__ path_all, int32_t* __restrict__ out_len) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_search) return;
  const int32_t from = start[s], to = goal[s];
  if (from < 0 || to < 0) { out_len[s] = 0; return; }   // a road without lanes: begin-route$ reports it
  double* best = best_all + (int64_t)s * n_items;
  int32_t* came = came_all + (int64_t)s * n_items;
  int32_t* path = path_all + (int64_t)s * n_items;
  HeapEnt* heap = heap_all + (int64_t)s * heap_cap;
  if (from == to) { out_len[s] = 0; return; }   // (a* returns (start): the route is just ((arrive-at p)))
  for (int64_t v = 0; v < n_items; ++v) came[v] = -2;   // unseen
  const double2 gm = mid[to];
  int64_t n_heap = 0;
  int32_t order = 0;
  auto push = [&](HeapEnt e) {
    int64_t i = n_heap++;
    while (i > 0) {
      const int64_t up = (i - 1) / 2;
      if (!heap_less(e, heap[up])) break;
      heap[i] = heap[up];
      i = up;
    }
    heap[i] = e;
  };
  auto pop = [&]() {
    const HeapEnt top = heap[0], last = heap[--n_heap];
    int64_t i = 0;
    for (;;) {
      int64_t c = 2 * i + 1;
      if (c >= n_heap) break;
      if (c + 1 < n_heap && heap_less(heap[c + 1], heap[c])) ++c;
      if (!heap_less(heap[c], last)) break;
      heap[i] = heap[c];
      i = c;
    }
    if (n_heap > 0) heap[i] = last;
    return top;
  };
  came[from] = -1;
  best[from] = 0.0;
  push(HeapEnt{0.0, 0, from});
  bool overflow = false;
  while (n_heap > 0) {
    const HeapEnt cur = pop();
    if (cur.node == to) break;
    const double base = best[cur.node], step = cost_out[cur.node];
    for (int64_t k = nbr_off[cur.node]; k < nbr_off[cur.node + 1]; ++k) {
      const int32_t nb = nbr[k];
      const double nw = __dadd_rn(base, step);
      if (came[nb] == -2 || nw < best[nb]) {
        best[nb] = nw;
        came[nb] = cur.node;
        const double2 m = mid[nb];
        const double dx = __dsub_rn(m.x, gm.x), dy = __dsub_rn(m.y, gm.y);
        const double h = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        if (n_heap >= heap_cap) { overflow = true; break; }
        push(HeapEnt{__dadd_rn(nw, h), ++order, nb});
      }
    }
    if (overflow) break;
  }
  if (overflow) { out_len[s] = -2; return; }
  if (came[to] == -2) { out_len[s] = -1; return; }   // no-path
  int32_t len = 0;
  for (int32_t v = to; v != from; v = came[v]) ++len;   // the route is (cdr lanes): the start lane is not part of it
  int32_t i = len;
  for (int32_t v = to; v != from; v = came[v]) path[--i] = v;
  out_len[s] = len;
}

}  // namespace griddy
