"""World -> structure-of-arrays buffers for the C ABI (include/griddy_cuda.h).

Ids are registration indices: item at list position i of `static-items` has id N-1-i
(world.scm:21-22 conses at the head).  Actor ids are assigned so that linking the actors in id order
rebuilds every segment's cons-list exactly (core.scm:169-171): the k-th actor of `get-actors` gets id
M-1-k.  The Guile glue in guile/griddy/cuda.scm does the same walk.
"""
from __future__ import annotations

from fractions import Fraction

import numpy as np

from . import core
from .flat import (BACK, FORW, ITEM_SLEEP, ITEM_TRAVEL, K_JLANE, K_JUNCTION, K_OTHER, K_SEGLANE,
                   K_SEGMENT, FlatActors, FlatNetwork)
from .route import find_route

_DIR = {"forw": FORW, "back": BACK}


def registration_index(world: core.World) -> dict:
    n = len(world.static_items)
    return {item: n - 1 - i for i, item in enumerate(world.static_items)}


def flatten_network(world: core.World):
    """-> (FlatNetwork, {item: id})"""
    idx = registration_index(world)
    n = len(idx)
    kind = np.full(n, K_OTHER, np.uint8)
    lane_len = np.zeros(n, np.float64)
    lane_dir = np.zeros(n, np.uint8)
    m1 = lambda: np.full(n, -1, np.int32)
    lane_segment, lane_out, lane_junction, so_f, so_b, lane_in = m1(), m1(), m1(), m1(), m1(), m1()
    for item, r in idx.items():
        if isinstance(item, core.RoadJunction):
            kind[r] = K_JUNCTION
        elif isinstance(item, core.RoadSegment):
            kind[r] = K_SEGMENT
            for side, arr in (("forw", so_f), ("back", so_b)):
                try:
                    arr[r] = idx[core.get_outer_lane(item, side)]
                except core.GriddyError:
                    arr[r] = -1
        elif isinstance(item, core.RoadLaneSegment):
            kind[r] = K_SEGLANE
            lane_len[r] = core.get_length(item)
            lane_dir[r] = _DIR[item.direction]
            lane_segment[r] = idx[item.segment]
        elif isinstance(item, core.RoadLaneJunction):
            kind[r] = K_JLANE
            lane_len[r] = core.get_length(item)
            lane_out[r] = idx[core.get_segment_lane(item)]
            lane_junction[r] = idx[item.junction]
            for in_lane, jls in item.junction.lane_map["inputs"].items():
                if item in jls:
                    lane_in[r] = idx[in_lane]
    # the graph find-route searches (route.scm:19-40): neighbours in the reference's list order, cost of leaving a
    # lane, midpoints — computed here with the API mirror's procedures, by the Guile glue with the reference's own
    by_id = sorted(idx.items(), key=lambda kv: kv[1])
    nbr_off, nbr = [0], []
    cost_out, midpoint = np.zeros(n, np.float64), np.zeros((n, 2), np.float64)
    for item, r in by_id:
        if isinstance(item, core.RoadLaneSegment):
            nbr.extend(idx[jl] for jl in core.get_junction_lanes(item))
            cost_out[r] = float(core.get_length(item.segment))
            midpoint[r] = core.get_midpoint(item)
        elif isinstance(item, core.RoadLaneJunction):
            nbr.append(idx[core.get_segment_lane(item)])
            midpoint[r] = core.get_midpoint(item)
        nbr_off.append(len(nbr))
    return FlatNetwork(kind, lane_len, lane_dir, lane_segment, lane_out, lane_junction, so_f, so_b,
                       lane_in=lane_in, nbr_off=np.array(nbr_off, np.int64), nbr=np.array(nbr, np.int32),
                       cost_out=cost_out, midpoint=midpoint), idx


def flatten_geometry(world: core.World, idx: dict) -> np.ndarray:
    """Drawing geometry, 8 doubles per item (include/griddy_cuda.h, griddy_upload_geometry): what the
    Guile glue computes with the reference's own get-pos / get-vec / curve (guile/griddy/cuda.scm)."""
    geom = np.zeros((len(idx), 8), np.float64)
    for item, r in idx.items():
        if isinstance(item, core.RoadLaneSegment):
            geom[r, 0:2] = core.get_pos(item, "beg")
            geom[r, 2:4] = core.get_vec(item)
        elif isinstance(item, core.RoadLaneJunction):
            geom[r] = [c for pt in item.curve for c in pt]
        elif isinstance(item, core.RoadSegment):
            geom[r, 0:2] = core.get_pos_of_location(core.LocationOffRoad(item, "forw", 0))
            geom[r, 2:4] = core.get_vec(item)
            geom[r, 4:6] = core.get_pos_of_location(core.LocationOffRoad(item, "back", 0))
    return geom


def speed_dt_of(max_speed) -> float:
    """(exact->inexact (* max-speed *simulate/time-step*)) under Guile's numeric tower."""
    if isinstance(max_speed, float):
        return max_speed * float(core.SIMULATE_TIME_STEP)
    return float(Fraction(max_speed) * core.SIMULATE_TIME_STEP)


def flatten_actors(world: core.World, idx: dict, routes: bool = True):
    """-> (FlatActors, [actor objects in id order])"""
    order = core.get_actors(world)
    actors = order[::-1]
    n = len(actors)
    loc_kind = np.zeros(n, np.uint8); container = np.zeros(n, np.int32); pos = np.zeros(n, np.float64)
    side = np.zeros(n, np.uint8); speed = np.zeros(n, np.float64); speed_dt = np.zeros(n, np.float64)
    agenda_off = np.zeros(n + 1, np.int64)
    ik, isl, iseg, iside, ipos = [], [], [], [], []
    route_off, route_lane = [0], []
    for a, actor in enumerate(actors):
        loc = actor.location
        if isinstance(loc, core.LocationOnRoad):
            loc_kind[a], container[a], cur = 1, idx[loc.road_lane], None
        else:
            loc_kind[a], container[a] = 0, idx[loc.road_segment]
            side[a] = _DIR[loc.road_side_direction]
            cur = loc
        pos[a] = float(loc.pos_param)
        speed[a] = float(actor.max_speed)
        speed_dt[a] = speed_dt_of(actor.max_speed)
        if actor.route != "none":
            raise core.GriddyError("unsupported", "actors must be uploaded with route 'none")
        for item in actor.agenda:
            if item[0] == "sleep-for":
                if int(item[1]) != item[1]:
                    raise core.GriddyError("unsupported", "sleep-for needs an exact integer")
                ik.append(ITEM_SLEEP); isl.append(int(item[1])); iseg.append(-1); iside.append(0); ipos.append(0.0)
            elif item[0] == "travel-to":
                dest = item[1]
                ik.append(ITEM_TRAVEL); isl.append(0); iseg.append(idx[dest.road_segment])
                iside.append(_DIR[dest.road_side_direction]); ipos.append(float(dest.pos_param))
                if routes and cur is not None:
                    steps = find_route(core.off_road_to_on_road(cur), core.off_road_to_on_road(dest))
                    route_lane.extend(idx[s[1]] for s in steps[:-1])
                # end-route$ leaves the actor off-road beside the destination lane
                # (location.scm:41-47), which is where the next travel item starts
                cur = core.on_road_to_off_road(core.off_road_to_on_road(dest))
            else:
                raise core.GriddyError("unknown-agenda-item", item)
            route_off.append(len(route_lane))
        agenda_off[a + 1] = len(ik)
    fa = FlatActors(loc_kind, container, pos, side, speed, speed_dt, agenda_off,
                    np.array(ik, np.uint8), np.array(isl, np.int32), np.array(iseg, np.int32),
                    np.array(iside, np.uint8), np.array(ipos, np.float64),
                    route_off=np.array(route_off, np.int64) if routes else None,
                    route_lane=np.array(route_lane, np.int32) if routes else None)
    return fa, actors
