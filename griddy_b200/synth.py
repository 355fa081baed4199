"""Procedural-grid generator binding (csrc/synth.cpp): examples/procedural-grid.scm, parametrised."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .flat import FlatActors, FlatNetwork

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DEFAULT_SEED = 0x9E3779B97F4A7C15


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "csrc", "libgriddy_synth.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _LIB = C.CDLL(path)
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def grid_sizes(n: int) -> dict:
    out = [C.c_int64() for _ in range(5)]
    rc = _lib().synth_grid_sizes(C.c_int(n), *[C.byref(o) for o in out])
    if rc:
        raise ValueError(f"synth_grid_sizes({n}) -> {rc}")
    keys = ("n_items", "n_junctions", "n_segments", "n_seg_lanes", "n_jlanes")
    return {k: int(o.value) for k, o in zip(keys, out)}


def grid_network(n: int, spacing: float = 150.0, origin: float = 50.0, geometry: bool = False) -> FlatNetwork:
    """n x n junction grid of procedural-grid.scm:13-67 (n = 5 there)."""
    sz = grid_sizes(n)
    ni = sz["n_items"]
    kind = np.empty(ni, np.uint8); lane_len = np.empty(ni, np.float64); lane_dir = np.empty(ni, np.uint8)
    i32 = lambda k=1: np.empty(ni * k, np.int32)
    lane_segment, lane_out, lane_junction, so_f, so_b, lane_in, lf, lt = (i32() for _ in range(8))
    turn4 = i32(4)
    seg_reg = np.empty(sz["n_segments"], np.int32)
    rc = _lib().synth_grid_network(C.c_int(n), C.c_double(spacing), C.c_double(origin), _p(kind),
                                   _p(lane_len), _p(lane_dir), _p(lane_segment), _p(lane_out),
                                   _p(lane_junction), _p(so_f), _p(so_b), _p(lane_in), _p(lf), _p(lt),
                                   _p(turn4), _p(seg_reg))
    if rc:
        raise RuntimeError(f"synth_grid_network -> {rc}")
    return FlatNetwork(kind, lane_len, lane_dir, lane_segment, lane_out, lane_junction, so_f, so_b,
                       lane_in=lane_in, grid_n=n, lane_from_j=lf, lane_to_j=lt,
                       turn4=turn4.reshape(ni, 4), seg_reg=seg_reg,
                       loc_key=grid_locality_keys(kind, lf, lane_junction, so_f, lane_segment, n * n),
                       geom=grid_geometry(n, spacing, origin) if geometry else None)


def grid_geometry(n: int, spacing: float = 150.0, origin: float = 50.0) -> np.ndarray:
    """Drawing geometry of the n x n grid, [n_items, 8] float64 (griddy_upload_geometry)."""
    geom = np.empty((grid_sizes(n)["n_items"], 8), np.float64)
    rc = _lib().synth_grid_geometry(C.c_int(n), C.c_double(spacing), C.c_double(origin), _p(geom))
    if rc:
        raise RuntimeError(f"synth_grid_geometry -> {rc}")
    return geom


def grid_locality_keys(kind, lane_from_j, lane_junction, seg_outer_forw, lane_segment, n_junctions) -> np.ndarray:
    """Locality key per item for griddy_set_locality_keys: items grouped by the junction they hang
    off (row-major), so that lanes an actor passes through in succession are close in slot order.
    A segment lane hangs off the junction it leaves, a junction lane off its junction, a segment off
    the junction its forw lanes leave."""
    key = np.empty(kind.shape[0], np.uint32)
    rc = _lib().synth_locality_keys(C.c_int64(kind.shape[0]), C.c_int64(n_junctions), _p(kind), _p(lane_from_j),
                                    _p(lane_junction), _p(seg_outer_forw), _p(lane_segment), _p(key))
    if rc:
        raise RuntimeError(f"synth_locality_keys -> {rc}")
    return key


def grid_actors(net: FlatNetwork, n_actors: int, seed: int = DEFAULT_SEED, agenda_len: int = 2, ids=None) -> FlatActors:
    """procedural-grid.scm:69-95: off-road starts, agenda alternating sleep-for / travel-to.
    ids (ascending int64): generate only these actors of the n_actors-actor world (a rank's share)."""
    if ids is not None:
        ids = np.ascontiguousarray(ids, np.int64)
    n, L = (n_actors if ids is None else len(ids)), agenda_len
    loc_kind = np.empty(n, np.uint8); container = np.empty(n, np.int32); pos = np.empty(n, np.float64)
    side = np.empty(n, np.uint8); speed = np.empty(n, np.float64); speed_dt = np.empty(n, np.float64)
    agenda_off = np.empty(n + 1, np.int64)
    item_kind = np.empty(n * L, np.uint8); item_sleep = np.empty(n * L, np.int32)
    item_seg = np.empty(n * L, np.int32); item_side = np.empty(n * L, np.uint8)
    item_pos = np.empty(n * L, np.float64)
    rc = _lib().synth_actors_ids(C.c_int64(n), _p(ids) if ids is not None else None, C.c_uint64(seed), C.c_int(L),
                                 C.c_int64(len(net.seg_reg)), _p(net.seg_reg), _p(loc_kind), _p(container), _p(pos),
                                 _p(side), _p(speed), _p(speed_dt), _p(agenda_off), _p(item_kind), _p(item_sleep),
                                 _p(item_seg), _p(item_side), _p(item_pos))
    if rc:
        raise RuntimeError(f"synth_actors_ids -> {rc}")
    return FlatActors(loc_kind, container, pos, side, speed, speed_dt, agenda_off, item_kind,
                      item_sleep, item_seg, item_side, item_pos)


def actor_ids_of_rank(net: FlatNetwork, n_actors: int, owner: np.ndarray, rank: int, seed: int = DEFAULT_SEED) -> np.ndarray:
    """Ids (ascending) of the actors of the n_actors-actor world that start on `rank`'s items."""
    ids = np.empty(n_actors, np.int64)
    n = C.c_int64(0)
    rc = _lib().synth_actor_ids_of_rank(C.c_int64(n_actors), C.c_uint64(seed), C.c_int64(len(net.seg_reg)), _p(net.seg_reg),
                                        _p(np.ascontiguousarray(owner, np.int32)), C.c_int32(rank), _p(ids), C.byref(n))
    if rc:
        raise RuntimeError(f"synth_actor_ids_of_rank -> {rc}")
    return ids[: n.value].copy()


def grid_owner(net: FlatNetwork, world: int, strips_per_rank: int = 1) -> np.ndarray:
    """Owner rank of every item: world * strips_per_rank horizontal strips of junction rows, dealt to the
    ranks in turn.  A junction owns its junction lanes; a segment and all its lanes belong to the
    junction the segment begins at."""
    owner = np.empty(net.n_items, np.int32)
    rc = _lib().synth_grid_owner(C.c_int64(net.n_items), C.c_int(net.grid_n), C.c_int(world), C.c_int(strips_per_rank),
                                 _p(net.kind), _p(net.lane_from_j), _p(net.lane_junction), _p(net.seg_outer_forw),
                                 _p(net.lane_segment), _p(owner))
    if rc:
        raise RuntimeError(f"synth_grid_owner -> {rc}")
    return owner
