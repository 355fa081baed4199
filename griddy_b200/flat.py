"""Flattened structure-of-arrays form of a griddy world.

Every id is an item's *registration index*: the order of `add!` calls in `make-skeleton`
(reference world.scm:21-22; Guile sees the list reversed, so index r is list position N-1-r).
These are exactly the buffers that cross the C ABI in include/griddy_cuda.h.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

# item kinds
K_JUNCTION, K_SEGMENT, K_SEGLANE, K_JLANE, K_OTHER = 0, 1, 2, 3, 4
# directions / sides
FORW, BACK = 0, 1
# agenda item kinds
ITEM_SLEEP, ITEM_TRAVEL = 0, 1
# route status
ROUTE_NONE, ROUTE_SOME, ROUTE_DONE = 0, 1, 2


def _arr(x, dtype):
    return np.ascontiguousarray(np.asarray(x, dtype=dtype))


@dataclass
class FlatNetwork:
    """Static road network, one entry per registered item."""

    kind: np.ndarray            # u8  item kind
    lane_len: np.ndarray        # f64 lane length (spatial.scm:28-46), 0 for non-lanes
    lane_dir: np.ndarray        # u8  segment lanes: FORW/BACK
    lane_segment: np.ndarray    # i32 segment lanes: their segment, else -1
    lane_out: np.ndarray        # i32 junction lanes: output segment lane (core.scm:106-111)
    lane_junction: np.ndarray   # i32 junction lanes: their junction
    seg_outer_forw: np.ndarray  # i32 segments: get-outer-lane for side forw (location.scm:16-24)
    seg_outer_back: np.ndarray  # i32 segments: get-outer-lane for side back
    # optional extras (host-side route finding / grid routing table)
    lane_in: Optional[np.ndarray] = None       # i32 junction lanes: input segment lane
    grid_n: int = 0                            # > 0: procedural grid with n x n junctions first
    lane_from_j: Optional[np.ndarray] = None   # i32 segment lanes: junction they leave
    lane_to_j: Optional[np.ndarray] = None     # i32 segment lanes: junction they reach
    turn4: Optional[np.ndarray] = None         # i32 [n_items,4] next-hop table by heading
    seg_reg: Optional[np.ndarray] = None       # i32 segment ordinal -> registration id
    loc_key: Optional[np.ndarray] = None       # u32 locality key per item (griddy_set_locality_keys)
    geom: Optional[np.ndarray] = None          # f64 [n_items,8] drawing geometry (griddy_upload_geometry)
    # the lane graph find-route searches (griddy_set_route_graph): neighbours per item in the reference's order (CSR),
    # the cost of leaving a lane, every lane's midpoint
    nbr_off: Optional[np.ndarray] = None       # i64 [n_items+1]
    nbr: Optional[np.ndarray] = None           # i32
    cost_out: Optional[np.ndarray] = None      # f64 [n_items]
    midpoint: Optional[np.ndarray] = None      # f64 [n_items,2]
    # a TILE of a larger network (griddy_set_item_ranges): every array above then has one entry per held item,
    # the id ranges back to back; the ids inside the arrays stay global
    n_global: Optional[int] = None             # items of the whole network
    range_beg: Optional[np.ndarray] = None     # i64 ascending, disjoint id ranges held
    range_end: Optional[np.ndarray] = None
    owner: Optional[np.ndarray] = None         # i32 owner rank per held item (partitioned worlds)

    def __post_init__(self):
        self.kind = _arr(self.kind, np.uint8)
        self.lane_len = _arr(self.lane_len, np.float64)
        self.lane_dir = _arr(self.lane_dir, np.uint8)
        for name in ("lane_segment", "lane_out", "lane_junction", "seg_outer_forw", "seg_outer_back"):
            setattr(self, name, _arr(getattr(self, name), np.int32))
        for name in ("lane_in", "lane_from_j", "lane_to_j", "turn4", "seg_reg"):
            v = getattr(self, name)
            if v is not None:
                setattr(self, name, _arr(v, np.int32))
        if self.loc_key is not None:
            self.loc_key = _arr(self.loc_key, np.uint32)
        if self.geom is not None:
            self.geom = _arr(self.geom, np.float64).reshape(-1, 8)
            assert self.geom.shape[0] == self.kind.shape[0], "geom needs 8 doubles per item"
        if self.nbr_off is not None:
            self.nbr_off = _arr(self.nbr_off, np.int64)
            self.nbr = _arr(self.nbr, np.int32)
            self.cost_out = _arr(self.cost_out, np.float64)
            self.midpoint = _arr(self.midpoint, np.float64).reshape(-1, 2)
        if self.range_beg is not None:
            self.range_beg = _arr(self.range_beg, np.int64)
            self.range_end = _arr(self.range_end, np.int64)
            assert int((self.range_end - self.range_beg).sum()) == self.kind.shape[0], "ranges do not match the arrays"
        if self.owner is not None:
            self.owner = _arr(self.owner, np.int32)

    @property
    def n_items(self) -> int:
        return int(self.kind.shape[0])

    @property
    def tiled(self) -> bool:
        return self.range_beg is not None

    def local(self, gid) -> np.ndarray:
        """Compact index of global ids (-1 where the tile does not hold the item)."""
        gid = np.asarray(gid, np.int64)
        if not self.tiled:
            return np.where((gid >= 0) & (gid < self.n_items), gid, -1)
        k = np.searchsorted(self.range_end, gid, side="right")
        k = np.minimum(k, len(self.range_beg) - 1)
        inside = (gid >= self.range_beg[k]) & (gid < self.range_end[k])
        off = np.concatenate([[0], np.cumsum(self.range_end - self.range_beg)])[:-1]
        return np.where(inside, off[k] + gid - self.range_beg[k], -1)

    def tile_of(self, range_beg, range_end, owner=None) -> "FlatNetwork":
        """The tile of this (whole) network that holds the given id ranges: what a rank would have generated."""
        rb, re = np.asarray(range_beg, np.int64), np.asarray(range_end, np.int64)
        idx = np.concatenate([np.arange(b, e) for b, e in zip(rb, re)]) if len(rb) else np.zeros(0, np.int64)
        pick = lambda a: None if a is None else a[idx]
        return FlatNetwork(self.kind[idx], self.lane_len[idx], self.lane_dir[idx], self.lane_segment[idx], self.lane_out[idx],
                           self.lane_junction[idx], self.seg_outer_forw[idx], self.seg_outer_back[idx],
                           lane_in=pick(self.lane_in), grid_n=self.grid_n, lane_from_j=pick(self.lane_from_j),
                           lane_to_j=pick(self.lane_to_j), turn4=pick(self.turn4), seg_reg=self.seg_reg,
                           loc_key=pick(self.loc_key), geom=pick(self.geom), n_global=self.n_items, range_beg=rb,
                           range_end=re, owner=None if owner is None else np.ascontiguousarray(owner[idx], np.int32))


@dataclass
class FlatActors:
    """Initial actors in `link!` order, their agendas (CSR) and, optionally, explicit routes.

    route_off / route_lane: for agenda item k (global index), the (turn-onto lane) steps of the
    route `find-route` returns for it (route.scm:48-51) — the trailing (arrive-at pos) is implied by
    the item's destination.  None when the network's grid routing table is used instead.
    """

    loc_kind: np.ndarray    # u8  0 off-road, 1 on-road
    container: np.ndarray   # i32 segment (off-road) or lane (on-road)
    pos: np.ndarray         # f64 pos-param
    side: np.ndarray        # u8  road-side-direction (off-road)
    speed: np.ndarray       # f64 max-speed
    speed_dt: np.ndarray    # f64 (exact->inexact (* max-speed *simulate/time-step*))
    agenda_off: np.ndarray  # i64 [n+1]
    item_kind: np.ndarray   # u8
    item_sleep: np.ndarray  # i32
    item_seg: np.ndarray    # i32
    item_side: np.ndarray   # u8
    item_pos: np.ndarray    # f64
    route_off: Optional[np.ndarray] = None   # i64 [n_agenda_items+1]
    route_lane: Optional[np.ndarray] = None  # i32
    # optional: travel-to destinations resolved by the generator (griddy_set_agenda_destinations), per agenda item
    dest_lane: Optional[np.ndarray] = None   # i32 get-outer-lane of (item_seg, item_side)
    dest_dir: Optional[np.ndarray] = None    # u8  that lane's direction
    dest_from_j: Optional[np.ndarray] = None # i32 junction it leaves (grid routing)
    dest_to_j: Optional[np.ndarray] = None   # i32 junction it reaches
    # optional, one rank's actors of a partitioned world with explicit routes (griddy_mg_set_routes): the WHOLE world's
    # route table and, per agenda item of these actors, its route in it
    route_table_off: Optional[np.ndarray] = None    # i64 [n_routes+1]
    route_table_lane: Optional[np.ndarray] = None   # i32
    item_route: Optional[np.ndarray] = None         # i32 [n_agenda_items]

    def __post_init__(self):
        self.loc_kind = _arr(self.loc_kind, np.uint8)
        self.container = _arr(self.container, np.int32)
        self.pos = _arr(self.pos, np.float64)
        self.side = _arr(self.side, np.uint8)
        self.speed = _arr(self.speed, np.float64)
        self.speed_dt = _arr(self.speed_dt, np.float64)
        self.agenda_off = _arr(self.agenda_off, np.int64)
        self.item_kind = _arr(self.item_kind, np.uint8)
        self.item_sleep = _arr(self.item_sleep, np.int32)
        self.item_seg = _arr(self.item_seg, np.int32)
        self.item_side = _arr(self.item_side, np.uint8)
        self.item_pos = _arr(self.item_pos, np.float64)
        if self.route_off is not None:
            self.route_off = _arr(self.route_off, np.int64)
            self.route_lane = _arr(self.route_lane if self.route_lane is not None else [], np.int32)
        if self.route_table_off is not None:
            self.route_table_off = _arr(self.route_table_off, np.int64)
            self.route_table_lane = _arr(self.route_table_lane if self.route_table_lane is not None else [], np.int32)
            self.item_route = _arr(self.item_route, np.int32)
        if self.dest_lane is not None:
            self.dest_lane = _arr(self.dest_lane, np.int32)
            self.dest_dir = _arr(self.dest_dir, np.uint8)
            self.dest_from_j = _arr(self.dest_from_j, np.int32)
            self.dest_to_j = _arr(self.dest_to_j, np.int32)

    @property
    def n(self) -> int:
        return int(self.loc_kind.shape[0])

    @property
    def n_agenda_items(self) -> int:
        return int(self.agenda_off[-1])


@dataclass
class ActorState:
    """Mutable per-actor state read back from a world (device or oracle)."""

    alive: np.ndarray         # u8  0 once replaced by an equal-key insert (core.scm:173-178)
    loc_kind: np.ndarray      # u8
    container: np.ndarray     # i32
    side: np.ndarray          # u8  (0 while on-road)
    pos: np.ndarray           # f64
    agenda_cur: np.ndarray    # i32 items popped so far
    sleep_left: np.ndarray    # i32 head's remaining time when it is (sleep-for t), else 0
    route_status: np.ndarray  # u8
    route_head: np.ndarray    # i32 lane of a (turn-onto l) head, -1 for (arrive-at p), -2 no head
    order: np.ndarray         # i32 `get-actors` order of the world (world.scm:24-30)

    INT_FIELDS = ("alive", "loc_kind", "container", "side", "agenda_cur", "sleep_left",
                  "route_status", "route_head")

    def diff(self, other: "ActorState", rtol: float = 1e-9) -> list:
        """Names of fields that differ (ints bit-exact, pos within rtol relative).

        A vanished actor is in no container, so only `alive` is compared for it.
        """
        bad = []
        if not np.array_equal(self.alive, other.alive):
            return ["alive"]
        m = self.alive.astype(bool)
        for f in self.INT_FIELDS[1:]:
            if not np.array_equal(getattr(self, f)[m], getattr(other, f)[m]):
                bad.append(f)
        if not np.array_equal(self.order, other.order):
            bad.append("order")
        a, b = self.pos[m], other.pos[m]
        if not np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b))):
            bad.append("pos")
        return bad
