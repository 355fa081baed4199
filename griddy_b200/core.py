"""Host-side mirror of the reference's domain API (griddy/core.scm and griddy/core/*.scm).

The reference is Guile Scheme and Guile is not in this image, so the caller side of the drop-in
boundary is mirrored in Python with the same names (``link!`` -> ``link``, ``add!`` -> ``add``,
``connect-by-rank!`` -> ``connect_by_rank`` ...), the same argument meaning and the same error
behaviour, so that ``make_skeleton`` / ``add_actors`` scripts read like examples/*.scm.  Network
construction runs once on the host and only feeds the flattening (flatten.py); none of this is on
the per-tick path.

Lists that the reference builds with ``insert!`` (cons at the head) are kept head-first here too.
Symbols are strings: 'forw 'back 'beg 'end 'none.
"""
from __future__ import annotations

import math
from fractions import Fraction
from typing import List, Optional

# constants.scm:23-39 (simulation values only)
ROAD_SEGMENT_WIGGLE_ROOM_PCT = 20
ROAD_LANE_WIDTH = 15
ROAD_LANE_APPROX_PTS = 5
ACTOR_SIZE = 5
ACTOR_SPEED = 80
SIMULATE_TIME_STEP = Fraction(1, 15)


class GriddyError(Exception):
    """(throw 'key args ...) of the reference."""

    def __init__(self, key, *args):
        super().__init__(key, *args)
        self.key = key


def flip(x):  # util.scm:26-31
    return {"forw": "back", "back": "forw", "beg": "end", "end": "beg"}[x]


# ----------------------------------------------------------------------------- world.scm
class StaticItem:
    pass


class ActorContainer:
    pass


class World:
    def __init__(self):
        self.static_items: List[StaticItem] = []   # head-first (reverse registration order)


# ----------------------------------------------------------------------------- road.scm
class RoadComponent(StaticItem):
    pass


class RoadLane(RoadComponent, ActorContainer):
    def __init__(self):
        self.actors = {}   # pos-param -> actor; stands for (make-bbtree <)   road.scm:33-34


class RoadLaneSegment(RoadLane):
    def __init__(self, direction):
        super().__init__()
        self.direction = direction
        self.segment: Optional["RoadSegment"] = None
        self.rank: Optional[int] = None


class RoadLaneJunction(RoadLane):
    def __init__(self, junction=None):
        super().__init__()
        self.junction = junction
        self.curve = None   # (p0 p1 p2 p3)


class RoadJunction(RoadComponent):
    def __init__(self, x, y):
        self.pos = (float(x), float(y))
        self.segments: List["RoadSegment"] = []
        self.lane_map = {"inputs": {}, "outputs": {}}


class RoadSegment(RoadComponent, ActorContainer):
    def __init__(self):
        self.actors: List["Actor"] = []   # off-road only, head-first
        self.junction = {"beg": None, "end": None}
        self.lanes = {"forw": [], "back": []}
        self.lane_count = {"forw": 0, "back": 0}


def add(world: World, item: StaticItem):
    """add! — world.scm:21-22, with the segment validation of road.scm:111-120."""
    if isinstance(item, RoadSegment):
        if item.junction["beg"] is None or item.junction["end"] is None:
            raise GriddyError("unlinked-road-segment")
        if not item.lanes["forw"] and not item.lanes["back"]:
            raise GriddyError("road-segment-has-no-lanes")
    world.static_items.insert(0, item)


def get_static_items(world: World, type_=None):   # world.scm:32-37
    if type_ is None:
        return world.static_items
    return [it for it in world.static_items if isinstance(it, type_)]


def get_lanes(obj, direction=None):   # road.scm:81-90
    if isinstance(obj, RoadJunction):
        return list(obj.lane_map["outputs"].keys())
    if direction is None:
        return obj.lanes["forw"] + obj.lanes["back"]
    return obj.lanes[direction]


def get_lane_count(segment: RoadSegment, direction=None):   # road.scm:92-97
    if direction is None:
        return segment.lane_count["forw"] + segment.lane_count["back"]
    return segment.lane_count[direction]


def get_actors(obj):
    """world.scm:24-30 and road.scm:99-108: lanes descending pos-param, segments as stored."""
    if isinstance(obj, World):
        out = []
        for c in obj.static_items:
            if isinstance(c, ActorContainer):
                out.extend(get_actors(c))
        return out
    if isinstance(obj, RoadSegment):
        return list(obj.actors)
    if isinstance(obj, RoadLane):
        return [obj.actors[k] for k in sorted(obj.actors, reverse=True)]
    raise GriddyError("not-implemented", obj)


# ----------------------------------------------------------------------------- actor.scm
class Actor:
    def __init__(self, max_speed=ACTOR_SPEED):
        self.location = None
        self.max_speed = max_speed
        self.route = "none"
        self.agenda: list = []
        self.alive = True


def agenda_push(actor: Actor, item):
    """agenda-push! — item is ('travel-to', <LocationOffRoad>) or ('sleep-for', n)."""
    actor.agenda = actor.agenda + [item]


def agenda_pop(actor: Actor):
    head, actor.agenda = actor.agenda[0], actor.agenda[1:]
    return head


def route_pop(actor: Actor):
    actor.route = actor.route[1:]


def route_reset(actor: Actor):
    actor.route = "none"


# ----------------------------------------------------------------------------- location.scm
class Location:
    def __init__(self, pos_param=0.0):
        self.pos_param = pos_param


class LocationOnRoad(Location):
    def __init__(self, road_lane, pos_param=0.0):
        super().__init__(pos_param)
        self.road_lane = road_lane


class LocationOffRoad(Location):
    def __init__(self, road_segment, road_side_direction, pos_param=0.0):
        super().__init__(pos_param)
        self.road_segment = road_segment
        self.road_side_direction = road_side_direction


def get_outer_lane(segment: RoadSegment, direction):   # location.scm:16-24
    back, forw = segment.lane_count["back"], segment.lane_count["forw"]
    if back == 0 and forw == 0:
        raise GriddyError("road-has-no-lanes")
    if direction == "back" and back == 0:
        return get_lanes(segment, "forw")[0]
    if direction == "forw" and forw == 0:
        return get_lanes(segment, "back")[-1]
    return get_lanes(segment, direction)[-1]


def _match_direction(lane, if_forw, if_back):   # util.scm:135-138
    return if_forw if lane.direction == "forw" else if_back


def on_road_to_off_road(loc: LocationOnRoad) -> LocationOffRoad:   # location.scm:41-47
    lane = loc.road_lane
    return LocationOffRoad(lane.segment, lane.direction,
                           _match_direction(lane, loc.pos_param, 1 - loc.pos_param))


def off_road_to_on_road(loc: LocationOffRoad) -> LocationOnRoad:   # location.scm:49-57
    lane = get_outer_lane(loc.road_segment, loc.road_side_direction)
    return LocationOnRoad(lane, _match_direction(lane, loc.pos_param, 1 - loc.pos_param))


# ----------------------------------------------------------------------------- spatial.scm
# fp64 restatement.  The reference evaluates these through Chickadee vec2 / bezier, which are not
# vendored; lengths cross the boundary as host-computed inputs, so this only has to be
# self-consistent, and it is kept operation-for-operation identical to csrc/synth.cpp.
def _vadd(a, b):
    return (a[0] + b[0], a[1] + b[1])


def _vsub(a, b):
    return (a[0] - b[0], a[1] - b[1])


def _vmul(a, s):
    s = float(s)
    return (a[0] * s, a[1] * s)


def _vmag(a):
    return math.sqrt(a[0] * a[0] + a[1] * a[1])


def get_radius(junction: RoadJunction):   # spatial.scm:53-66
    m = max(get_lane_count(s) for s in junction.segments)
    wiggle = Fraction(ROAD_SEGMENT_WIGGLE_ROOM_PCT, 100) + 1
    return wiggle * Fraction(1, 2) * m * ROAD_LANE_WIDTH


def get_tangent_vec(obj):   # spatial.scm:160-167
    if isinstance(obj, RoadSegment):
        d = _vsub(obj.junction["end"].pos, obj.junction["beg"].pos)
        m = _vmag(d)
        return (d[0] / m, d[1] / m)
    return _vmul(get_tangent_vec(obj.segment), _match_direction(obj, 1.0, -1.0))


def get_ortho_vec(segment: RoadSegment):   # spatial.scm:169-170, math.scm:28-29
    t = get_tangent_vec(segment)
    c, s = math.cos(math.pi / 2.0), math.sin(math.pi / 2.0)
    return (t[0] * c - t[1] * s, t[0] * s + t[1] * c)


def get_offset(lane: RoadLaneSegment):   # spatial.scm:75-102
    seg = lane.segment
    if lane.direction == "back":
        from_edge = get_lane_count(seg, "back") - lane.rank - 1
    else:
        from_edge = get_lane_count(seg, "back") + lane.rank
    o = get_ortho_vec(seg)
    seg_edge = _vmul(o, Fraction(-1, 2) * ROAD_LANE_WIDTH * get_lane_count(seg))
    lane_edge = _vadd(seg_edge, _vmul(o, from_edge * ROAD_LANE_WIDTH))
    return _vadd(lane_edge, _vmul(o, Fraction(1, 2) * ROAD_LANE_WIDTH))


def get_pos(obj, beg_or_end=None):   # spatial.scm:104-122
    if isinstance(obj, RoadLaneSegment):
        return _vadd(get_pos(obj.segment, _match_direction(obj, beg_or_end, flip(beg_or_end))),
                     get_offset(obj))
    if isinstance(obj, RoadLaneJunction):
        return obj.curve[0] if beg_or_end == "beg" else obj.curve[3]
    if isinstance(obj, RoadSegment):
        junction = obj.junction[beg_or_end]
        sign = 1 if beg_or_end == "beg" else -1
        return _vadd(junction.pos, _vmul(get_tangent_vec(obj), sign * get_radius(junction)))
    raise GriddyError("not-implemented", obj)


def get_width(segment: RoadSegment):   # spatial.scm:47-50
    return (1 + Fraction(ROAD_SEGMENT_WIGGLE_ROOM_PCT, 100)) * ROAD_LANE_WIDTH * get_lane_count(segment)


def get_pos_of_location(loc):   # spatial.scm:125-146
    """(get-pos location): where `draw` puts the actor's circle (draw.scm:74-76)."""
    if isinstance(loc, LocationOffRoad):
        seg = loc.road_segment
        sign = 1 if loc.road_side_direction == "forw" else -1
        scale = sign * Fraction(1, 2) * get_width(seg) * (1 + 2 * Fraction(ROAD_SEGMENT_WIGGLE_ROOM_PCT, 100))
        v_offset = _vmul(get_ortho_vec(seg), scale)
        return _vadd(_vadd(get_pos(seg, "beg"), v_offset), _vmul(get_vec(seg), loc.pos_param))
    lane = loc.road_lane
    if isinstance(lane, RoadLaneSegment):
        return _vadd(get_pos(lane, "beg"), _vmul(get_vec(lane), loc.pos_param))
    return _bezier_at(lane.curve, loc.pos_param)


def get_vec(obj):   # spatial.scm:152-158
    return _vsub(get_pos(obj, "end"), get_pos(obj, "beg"))


def _bezier_at(curve, t):
    p0, p1, p2, p3 = curve
    t = float(t)
    u = 1.0 - t
    a, b, c, d = u * u * u, 3.0 * u * u * t, 3.0 * u * t * t, t * t * t
    return (a * p0[0] + b * p1[0] + c * p2[0] + d * p3[0],
            a * p0[1] + b * p1[1] + c * p2[1] + d * p3[1])


def get_length(obj) -> float:   # spatial.scm:28-46
    if isinstance(obj, RoadSegment):
        return _vmag(get_vec(obj))
    if isinstance(obj, RoadLaneSegment):
        return get_length(obj.segment)
    n = ROAD_LANE_APPROX_PTS
    acc = 0.0
    lo = _bezier_at(obj.curve, Fraction(0))
    for k in range(1, n + 1):
        hi = _bezier_at(obj.curve, Fraction(k, n))
        acc = acc + _vmag(_vsub(hi, lo))
        lo = hi
    return acc


def get_midpoint(obj):   # spatial.scm:68-73
    if isinstance(obj, RoadLaneJunction):
        return _bezier_at(obj.curve, Fraction(1, 2))
    return _vadd(get_pos(obj, "beg"), _vmul(get_vec(obj), Fraction(1, 2)))


# ----------------------------------------------------------------------------- core.scm
def get_direction(incoming_or_outgoing, junction, segment):   # core.scm:69-83
    if junction is segment.junction["end"]:
        return "forw" if incoming_or_outgoing == "incoming" else "back"
    if junction is segment.junction["beg"]:
        return "back" if incoming_or_outgoing == "incoming" else "forw"
    raise GriddyError("no-connection", junction, segment)


def get_segment_lanes(incoming_or_outgoing, junction):   # core.scm:85-95
    lanes = []
    for segment in junction.segments:
        lanes = lanes + get_lanes(segment, get_direction(incoming_or_outgoing, junction, segment))
    return lanes


def get_junction_lanes(lane: RoadLaneSegment):   # core.scm:97-104
    junction = lane.segment.junction[_match_direction(lane, "end", "beg")]
    return junction.lane_map["inputs"].get(lane, [])


def get_segment_lane(lane: RoadLaneJunction):   # core.scm:106-111
    return lane.junction.lane_map["outputs"][lane]


def link(*args):
    """link! — dispatches on argument types like the GOOPS generic (core.scm:151-206)."""
    if len(args) == 2 and isinstance(args[0], RoadLaneSegment) and isinstance(args[1], RoadSegment):
        lane, segment = args
        if lane.segment is not None:
            raise GriddyError("lane-already-linked", lane, segment)
        lane.segment = segment
        rank = segment.lane_count[lane.direction]
        lane.rank = rank
        segment.lane_count[lane.direction] = rank + 1
        segment.lanes[lane.direction] = segment.lanes[lane.direction] + [lane]
        return
    if len(args) == 3 and isinstance(args[0], RoadJunction):
        j1, segment, j2 = args
        j1.segments.insert(0, segment)
        j2.segments.insert(0, segment)
        segment.junction["beg"] = j1
        segment.junction["end"] = j2
        return
    if len(args) == 2 and isinstance(args[0], Actor) and isinstance(args[1], LocationOffRoad):
        actor, loc = args
        actor.location = loc
        loc.road_segment.actors.insert(0, actor)
        return
    if len(args) == 2 and isinstance(args[0], Actor) and isinstance(args[1], LocationOnRoad):
        actor, loc = args
        actor.location = loc
        old = loc.road_lane.actors.get(loc.pos_param)
        if old is not None and old is not actor:
            old.alive = False   # bbtree-set replaced it: it is in no container any more
        loc.road_lane.actors[loc.pos_param] = actor
        return
    if len(args) == 3 and isinstance(args[1], RoadLaneJunction):
        in_lane, junction_lane, out_lane = args

        def get_junction(lane, lane_type):
            key = {("in", "forw"): "end", ("in", "back"): "beg",
                   ("out", "forw"): "beg", ("out", "back"): "end"}[(lane_type, lane.direction)]
            return lane.segment.junction[key]

        if get_junction(in_lane, "in") is not get_junction(out_lane, "out"):
            raise GriddyError("no-connection", in_lane, out_lane)
        junction = get_junction(in_lane, "in")
        offset = 2 * get_radius(junction) * Fraction(ROAD_SEGMENT_WIGGLE_ROOM_PCT, 100)
        p0 = get_pos(in_lane, "end")
        p3 = get_pos(out_lane, "beg")
        p1 = _vadd(p0, _vmul(get_tangent_vec(in_lane), offset))
        p2 = _vsub(p3, _vmul(get_tangent_vec(out_lane), offset))
        junction_lane.junction = junction
        junction_lane.curve = (p0, p1, p2, p3)
        return
    raise GriddyError("no-applicable-method", "link!", args)


def connect(in_lane, out_lane, world):   # core.scm:113-122
    junction_lane = RoadLaneJunction()
    link(in_lane, junction_lane, out_lane)
    lane_maps = junction_lane.junction.lane_map
    lane_maps["inputs"][in_lane] = [junction_lane] + lane_maps["inputs"].get(in_lane, [])
    lane_maps["outputs"][junction_lane] = out_lane
    add(world, junction_lane)


def connect_all(junction, world):   # core.scm:124-133
    in_lanes = get_segment_lanes("incoming", junction)
    out_lanes = get_segment_lanes("outgoing", junction)
    for in_lane in in_lanes:
        for out_lane in out_lanes:
            if out_lane is not in_lane:
                connect(in_lane, out_lane, world)


def connect_by_rank(junction, world):   # core.scm:135-148
    for in_segment in list(junction.segments):
        for out_segment in list(junction.segments):
            ins = get_lanes(in_segment, get_direction("incoming", junction, in_segment))[::-1]
            outs = get_lanes(out_segment, get_direction("outgoing", junction, out_segment))[::-1]
            for a, b in zip(ins, outs):   # srfi-1 for-each stops at the shorter list
                connect(a, b, world)
