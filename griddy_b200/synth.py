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


def rank_rows(n: int, world: int, strips_per_rank: int, rank: int):
    """Junction rows `rank` owns: world * strips_per_rank horizontal strips dealt to the ranks in turn -> (beg, end) arrays."""
    rb, re = np.zeros(strips_per_rank + 1, np.int32), np.zeros(strips_per_rank + 1, np.int32)
    k = C.c_int32(0)
    rc = _lib().synth_rank_rows(C.c_int(n), C.c_int(world), C.c_int(strips_per_rank), C.c_int(rank), _p(rb), _p(re), C.byref(k))
    if rc:
        raise ValueError(f"synth_rank_rows -> {rc}")
    return rb[: k.value].copy(), re[: k.value].copy()


def tile_rows(n: int, world: int, strips_per_rank: int, rank: int, ring: int = 1):
    """The rows a rank HOLDS: the rows it owns plus `ring` rows either side (everything an actor of the rank reads
    before it crosses a boundary hangs off those), merged into ascending disjoint intervals."""
    rb, re = rank_rows(n, world, strips_per_rank, rank)
    out = []
    for b, e in zip(np.maximum(rb - ring, 0).tolist(), np.minimum(re + ring, n).tolist()):
        if out and b <= out[-1][1]:
            out[-1][1] = max(out[-1][1], e)
        else:
            out.append([b, e])
    return np.array([o[0] for o in out], np.int32), np.array([o[1] for o in out], np.int32)


def tile_ranges(n: int, row_beg, row_end):
    """The id ranges (beg, end arrays, ascending, empty ones dropped) of everything that hangs off the given junction rows."""
    rb, re = np.ascontiguousarray(row_beg, np.int32), np.ascontiguousarray(row_end, np.int32)
    gb, ge = np.zeros(4 * len(rb), np.int64), np.zeros(4 * len(rb), np.int64)
    nl = C.c_int64(0)
    rc = _lib().synth_tile_ranges(C.c_int(n), C.c_int(len(rb)), _p(rb), _p(re), _p(gb), _p(ge), C.byref(nl))
    if rc:
        raise ValueError(f"synth_tile_ranges -> {rc}")
    keep = ge > gb
    return gb[keep], ge[keep]


def grid_tile(n: int, row_beg, row_end, world: int = 1, strips_per_rank: int = 1, spacing: float = 150.0,
              origin: float = 50.0, geometry: bool = False) -> FlatNetwork:
    """The part of the n x n grid that hangs off the junction rows [row_beg[k], row_end[k]): generated from closed forms,
    without the rest of the grid.  Arrays are compact (FlatNetwork.range_beg / range_end), ids inside them global."""
    rb, re = np.ascontiguousarray(row_beg, np.int32), np.ascontiguousarray(row_end, np.int32)
    k = len(rb)
    gb, ge = np.zeros(4 * k, np.int64), np.zeros(4 * k, np.int64)
    nl = C.c_int64(0)
    rc = _lib().synth_tile_ranges(C.c_int(n), C.c_int(k), _p(rb), _p(re), _p(gb), _p(ge), C.byref(nl))
    if rc:
        raise ValueError(f"synth_tile_ranges -> {rc}")
    ni = int(nl.value)
    kind = np.empty(ni, np.uint8); lane_len = np.empty(ni, np.float64); lane_dir = np.empty(ni, np.uint8)
    i32 = lambda m=1: np.empty(ni * m, np.int32)
    lane_segment, lane_out, lane_junction, so_f, so_b, lane_in, lf, lt, hang = (i32() for _ in range(9))
    turn4 = i32(4)
    geom = np.empty((ni, 8), np.float64) if geometry else None
    rc = _lib().synth_tile_network(C.c_int(n), C.c_double(spacing), C.c_double(origin), C.c_int(k), _p(rb), _p(re), _p(kind),
                                   _p(lane_len), _p(lane_dir), _p(lane_segment), _p(lane_out), _p(lane_junction), _p(so_f),
                                   _p(so_b), _p(lane_in), _p(lf), _p(lt), _p(turn4), _p(hang),
                                   _p(geom) if geometry else None)
    if rc:
        raise RuntimeError(f"synth_tile_network -> {rc}")
    key = np.empty(ni, np.uint32)
    rc = _lib().synth_keys_from_hang(C.c_int64(ni), C.c_int64(n * n), _p(hang), _p(key))
    if rc:
        raise RuntimeError(f"synth_keys_from_hang -> {rc}")
    owner = np.empty(ni, np.int32)
    rc = _lib().synth_tile_owner(C.c_int(n), C.c_int(k), _p(rb), _p(re), C.c_int(world), C.c_int(strips_per_rank), _p(owner))
    if rc:
        raise RuntimeError(f"synth_tile_owner -> {rc}")
    keep = ge > gb
    return FlatNetwork(kind, lane_len, lane_dir, lane_segment, lane_out, lane_junction, so_f, so_b, lane_in=lane_in,
                       grid_n=n, lane_from_j=lf, lane_to_j=lt, turn4=turn4.reshape(ni, 4), loc_key=key, geom=geom,
                       n_global=grid_sizes(n)["n_items"], range_beg=gb[keep], range_end=ge[keep], owner=owner)


def grid_actors_of_rank(n: int, n_actors: int, world: int, strips_per_rank: int, rank: int, seed: int = DEFAULT_SEED,
                        agenda_len: int = 2):
    """The actors of the n_actors-actor world on the n x n grid that start on `rank`'s strips, generated from closed
    forms (no network needed), with their travel-to destinations resolved -> (FlatActors, ids)."""
    ids = np.empty(n_actors, np.int64)
    m = C.c_int64(0)
    rc = _lib().synth_actor_ids_of_rank_grid(C.c_int64(n_actors), C.c_uint64(seed), C.c_int(n), C.c_int(world),
                                             C.c_int(strips_per_rank), C.c_int32(rank), _p(ids), C.byref(m))
    if rc:
        raise RuntimeError(f"synth_actor_ids_of_rank_grid -> {rc}")
    ids = ids[: m.value].copy()
    return grid_actors_closed_form(n, n_actors, seed=seed, agenda_len=agenda_len, ids=ids), ids


def grid_actors_closed_form(n: int, n_actors: int, seed: int = DEFAULT_SEED, agenda_len: int = 2, ids=None) -> FlatActors:
    """grid_actors without a network: segment ids from the closed forms, destinations resolved (dest_* filled)."""
    if ids is not None:
        ids = np.ascontiguousarray(ids, np.int64)
    m, L = (n_actors if ids is None else len(ids)), agenda_len
    n_segments = 2 * n * (n - 1)
    loc_kind = np.empty(m, np.uint8); container = np.empty(m, np.int32); pos = np.empty(m, np.float64)
    side = np.empty(m, np.uint8); speed = np.empty(m, np.float64); speed_dt = np.empty(m, np.float64)
    agenda_off = np.empty(m + 1, np.int64)
    item_kind = np.empty(m * L, np.uint8); item_sleep = np.empty(m * L, np.int32)
    item_seg = np.empty(m * L, np.int32); item_side = np.empty(m * L, np.uint8)
    item_pos = np.empty(m * L, np.float64)
    dl = np.empty(m * L, np.int32); dd = np.empty(m * L, np.uint8); df = np.empty(m * L, np.int32); dt = np.empty(m * L, np.int32)
    rc = _lib().synth_actors_ids(C.c_int64(m), _p(ids) if ids is not None else None, C.c_uint64(seed), C.c_int(L),
                                 C.c_int64(n_segments), None, _p(loc_kind), _p(container), _p(pos),
                                 _p(side), _p(speed), _p(speed_dt), _p(agenda_off), _p(item_kind), _p(item_sleep),
                                 _p(item_seg), _p(item_side), _p(item_pos), C.c_int(n), _p(dl), _p(dd), _p(df), _p(dt))
    if rc:
        raise RuntimeError(f"synth_actors_ids -> {rc}")
    return FlatActors(loc_kind, container, pos, side, speed, speed_dt, agenda_off, item_kind,
                      item_sleep, item_seg, item_side, item_pos, dest_lane=dl, dest_dir=dd, dest_from_j=df, dest_to_j=dt)


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
                                 _p(item_seg), _p(item_side), _p(item_pos), C.c_int(0), None, None, None, None)
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
