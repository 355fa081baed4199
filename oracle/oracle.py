"""ctypes binding of the CPU oracle (oracle/griddy_oracle.cpp).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libgriddy_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle`")
        _LIB = C.CDLL(path)
        _LIB.go_create.restype = C.c_void_p
        _LIB.go_last_error.restype = C.c_char_p
        _LIB.go_vanished.restype = C.c_int64
        _LIB.go_ticks.restype = C.c_int64
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def get_pos(net, geom, loc_kind, container, side, pos):
    """(get-pos actor) for every actor (spatial.scm:125-149) -> float64 [n, 2]."""
    n = len(pos)
    as_ = lambda a, dt: np.ascontiguousarray(np.asarray(a, dt))
    geom = as_(geom, np.float64)
    assert geom.size == 8 * net.n_items
    lk, ct, sd, ps = as_(loc_kind, np.uint8), as_(container, np.int32), as_(side, np.uint8), as_(pos, np.float64)
    xy = np.zeros((n, 2), np.float64)
    rc = lib().go_get_pos(C.c_int64(net.n_items), C.c_int64(n), _p(net.kind), _p(geom), _p(lk), _p(ct), _p(sd), _p(ps), _p(xy))
    if rc:
        raise RuntimeError(f"oracle get_pos error {rc}")
    return xy


class Oracle:
    """Same call shape as griddy_b200.device.Simulation, so parity tests feed both identically."""

    def __init__(self, net, actors, dt: float = 1.0 / 15.0, two_size: float = 10.0, lazy_routes: bool = True):
        from griddy_b200.flat import ActorState  # data containers only, no product code path
        self._State = ActorState
        L = lib()
        self.h = C.c_void_p(L.go_create())
        self.net, self.actors = net, actors
        self._chk(L.go_set_constants(self.h, C.c_double(dt), C.c_double(two_size)))
        self._chk(L.go_load_network(self.h, C.c_int64(net.n_items), _p(net.kind), _p(net.lane_len),
                                    _p(net.lane_dir), _p(net.lane_segment), _p(net.lane_out),
                                    _p(net.lane_junction), _p(net.seg_outer_forw), _p(net.seg_outer_back)))
        if actors.route_off is None:
            if not net.grid_n:
                raise ValueError("actors carry no routes and the network has no grid routing table")
            self._chk(L.go_set_routing_grid(self.h, C.c_int(net.grid_n), _p(net.lane_from_j),
                                            _p(net.lane_to_j), _p(net.turn4)))
        self._chk(L.go_set_lazy_routes(self.h, C.c_int(1 if lazy_routes else 0)))
        a = actors
        self._chk(L.go_load_actors(self.h, C.c_int64(a.n), _p(a.loc_kind), _p(a.container), _p(a.pos),
                                   _p(a.side), _p(a.speed), _p(a.speed_dt), _p(a.agenda_off),
                                   _p(a.item_kind), _p(a.item_sleep), _p(a.item_seg), _p(a.item_side),
                                   _p(a.item_pos), _p(a.route_off), _p(a.route_lane)))

    def _chk(self, rc):
        if rc:
            raise RuntimeError(f"oracle error {rc}: {lib().go_last_error(self.h).decode()}")

    def step(self, n_ticks: int = 1):
        self._chk(lib().go_step(self.h, C.c_int64(n_ticks)))

    def state(self):
        n = self.actors.n
        u8 = lambda: np.zeros(n, np.uint8)
        i32 = lambda: np.zeros(n, np.int32)
        alive, loc_kind, side, rs = u8(), u8(), u8(), u8()
        container, agenda_cur, sleep_left, route_head, order = i32(), i32(), i32(), i32(), i32()
        pos = np.zeros(n, np.float64)
        n_order = C.c_int64()
        self._chk(lib().go_get_state(self.h, _p(alive), _p(loc_kind), _p(container), _p(side), _p(pos),
                                     _p(agenda_cur), _p(sleep_left), _p(rs), _p(route_head), _p(order),
                                     C.byref(n_order)))
        return self._State(alive, loc_kind, container, side, pos, agenda_cur, sleep_left, rs,
                           route_head, order[: n_order.value].copy())

    def occupancy(self):
        lane_count = np.zeros(self.net.n_items, np.int32)
        junction_count = np.zeros(self.net.n_items, np.int32)
        self._chk(lib().go_get_occupancy(self.h, _p(lane_count), _p(junction_count)))
        return lane_count, junction_count

    @property
    def vanished(self) -> int:
        return int(lib().go_vanished(self.h))

    def close(self):
        if self.h:
            lib().go_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
