"""Host-side route provider: find-route of griddy/simulate/route.scm:17-51.

Runs once per `travel-to`, never per tick, so it stays on the host: the flattener calls it for every
travel item and uploads the resulting lane lists (include/griddy_cuda.h, route_off / route_lane).
The reference delegates to Chickadee's `a*` (chickadee data path-finding, not vendored); its heap
tie-breaking among equal-cost paths cannot be known here, so ties are broken by insertion order and
that is a stated deviation.  When start = goal the path is just (start), giving ((arrive-at p)).
"""
from __future__ import annotations

import heapq
import itertools
import math

from . import core


def _neighbors(lane):   # route.scm:19-23
    if isinstance(lane, core.RoadLaneSegment):
        return core.get_junction_lanes(lane)
    return [core.get_segment_lane(lane)]


def _cost(lane_1, lane_2):   # route.scm:25-34
    if isinstance(lane_1, core.RoadLaneSegment):
        return core.get_length(lane_1.segment)
    return 0


def _distance(lane_1, lane_2):   # route.scm:36-40
    a, b = core.get_midpoint(lane_1), core.get_midpoint(lane_2)
    return math.sqrt((a[0] - b[0]) ** 2 + (a[1] - b[1]) ** 2)


def a_star(start, goal):
    """Lanes from start to goal inclusive, or None."""
    counter = itertools.count()
    frontier = [(0.0, next(counter), start)]
    came_from = {start: None}
    cost_so_far = {start: 0.0}
    while frontier:
        _, _, current = heapq.heappop(frontier)
        if current is goal:
            break
        for nxt in _neighbors(current):
            new_cost = cost_so_far[current] + _cost(current, nxt)
            if nxt not in cost_so_far or new_cost < cost_so_far[nxt]:
                cost_so_far[nxt] = new_cost
                heapq.heappush(frontier, (new_cost + _distance(goal, nxt), next(counter), nxt))
                came_from[nxt] = current
    if goal not in came_from:
        return None
    path, node = [], goal
    while node is not None:
        path.append(node)
        node = came_from[node]
    return path[::-1]


def find_route(init: core.LocationOnRoad, dest: core.LocationOnRoad):
    """(('turn-onto', lane) ... ('arrive-at', pos-param))   route.scm:42-51"""
    lanes = a_star(init.road_lane, dest.road_lane)
    if lanes is None:
        raise core.GriddyError("no-route", init, dest)
    return [("turn-onto", l) for l in lanes[1:]] + [("arrive-at", dest.pos_param)]
