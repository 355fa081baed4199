"""ctypes binding of libgriddy_cuda.so — the same symbols Guile binds with (system foreign).

There is no fallback: if the CUDA library is not built, or no CUDA device is present, constructing a
Simulation raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .flat import ActorState, FlatActors, FlatNetwork

_HERE = os.path.dirname(os.path.abspath(__file__))
# GRIDDY_CUDA_LIB names another build of the same library (kernel experiments, profiles/variants.md)
LIB_PATH = os.environ.get("GRIDDY_CUDA_LIB") or os.path.join(_HERE, "csrc", "libgriddy_cuda.so")
_LIB = None

ERR_NAMES = {1: "BAD_ARG", 2: "CUDA", 3: "NCCL", 4: "BAD_STATE", 5: "OOM"}

# every symbol include/griddy_cuda.h declares
SYMBOLS = ("griddy_abi_version", "griddy_create", "griddy_destroy", "griddy_last_error",
           "griddy_set_constants", "griddy_upload_network", "griddy_set_routing_grid",
           "griddy_upload_actors", "griddy_step", "griddy_profile_step", "griddy_download_actors",
           "griddy_download_occupancy", "griddy_last_step_ms", "griddy_tick", "griddy_kernel_launches",
           "griddy_set_locality_keys", "griddy_set_reorder_period", "griddy_slots_in_use",
           "griddy_debug_sort_pairs", "griddy_count_alive", "griddy_mg_configure", "griddy_mg_set_global_ids",
           "griddy_set_item_ranges", "griddy_set_agenda_destinations", "griddy_mg_plan_tile", "griddy_step_async",
           "griddy_reorder_now", "griddy_reorders", "griddy_set_route_graph", "griddy_download_routes", "griddy_mg_connect_local", "griddy_mg_step_local",
           "griddy_mg_plan", "griddy_download_local", "griddy_mg_neighbors", "griddy_mg_ipc_export",
           "griddy_mg_ipc_import", "griddy_upload_geometry", "griddy_download_positions", "griddy_positions_begin",
           "griddy_positions_end", "griddy_host_alloc", "griddy_host_free", "griddy_debug_trace", "griddy_state_digest",
           "griddy_delta_frame_bytes", "griddy_delta_begin", "griddy_delta_end", "griddy_delta_apply", "griddy_last_frame_ms",
           "griddy_last_copy_ms", "griddy_mg_set_routes")


class GriddyCudaError(RuntimeError):
    """Nonzero return from the C ABI; mirrors (throw 'griddy-cuda code msg) in the Guile glue."""

    def __init__(self, code, msg):
        super().__init__(f"griddy-cuda {ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is not built (python -c 'import __graft_entry__ as g; g.build()'); "
                               "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.griddy_last_error.restype = C.c_char_p
        L.griddy_last_error.argtypes = [C.c_void_p]
        L.griddy_last_step_ms.restype = C.c_double
        L.griddy_last_step_ms.argtypes = [C.c_void_p]
        L.griddy_last_frame_ms.restype = C.c_double
        L.griddy_last_frame_ms.argtypes = [C.c_void_p]
        L.griddy_last_copy_ms.restype = C.c_double
        L.griddy_last_copy_ms.argtypes = [C.c_void_p]
        L.griddy_tick.restype = C.c_int64
        L.griddy_tick.argtypes = [C.c_void_p]
        L.griddy_kernel_launches.restype = C.c_int64
        L.griddy_kernel_launches.argtypes = [C.c_void_p]
        L.griddy_slots_in_use.restype = C.c_int64
        L.griddy_slots_in_use.argtypes = [C.c_void_p]
        L.griddy_reorders.restype = C.c_int64
        L.griddy_reorders.argtypes = [C.c_void_p]
        L.griddy_destroy.restype = None
        L.griddy_destroy.argtypes = [C.c_void_p]
        L.griddy_host_alloc.restype = C.c_void_p
        L.griddy_host_alloc.argtypes = [C.c_int64]
        L.griddy_host_free.restype = None
        L.griddy_host_free.argtypes = [C.c_void_p]
        L.griddy_delta_frame_bytes.restype = C.c_int64
        L.griddy_delta_frame_bytes.argtypes = [C.c_int64]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Simulation:
    """One world on one GPU: upload once, step, read back.

    With `partition=(rank, world, owner, mig_cap, extra_slots)` the context holds one part of a
    spatially partitioned world (include/griddy_cuda.h, "Multi-GPU"): `actors` are then the actors on
    this rank's items, `gids` their global ids; `owner` has one entry per item of `net`.  `net` may be a
    TILE (FlatNetwork.range_beg / range_end: the rank's own rows plus the ring next to them): its arrays are
    compact and `actors` must carry their resolved destinations (dest_*)."""

    def __init__(self, net: FlatNetwork, actors: FlatActors, device: int = -1,
                 dt: float = 1.0 / 15.0, two_size: float = 10.0, reorder_period: int | None = None,
                 partition=None, gids=None, max_agenda_items=None):
        L = lib()
        self.h = C.c_void_p()
        rc = L.griddy_create(C.byref(self.h), C.c_int(device))
        if rc:
            self.h = None
            raise GriddyCudaError(rc, "griddy_create failed (no CUDA device?)")
        self.net, self.actors = net, actors
        self._chk(L.griddy_set_constants(self.h, C.c_double(dt), C.c_double(two_size)))
        if net.tiled:
            self._chk(L.griddy_set_item_ranges(self.h, C.c_int64(net.n_global), C.c_int32(len(net.range_beg)),
                                               _p(net.range_beg), _p(net.range_end)))
        self._chk(L.griddy_upload_network(self.h, C.c_int64(net.n_items), _p(net.kind), _p(net.lane_len),
                                          _p(net.lane_dir), _p(net.lane_segment), _p(net.lane_out),
                                          _p(net.lane_junction), _p(net.seg_outer_forw),
                                          _p(net.seg_outer_back)))
        if net.loc_key is not None:
            self._chk(L.griddy_set_locality_keys(self.h, _p(net.loc_key)))
        if net.geom is not None:
            self._chk(L.griddy_upload_geometry(self.h, _p(net.geom)))
        self._frames = []
        if reorder_period is not None:
            self._chk(L.griddy_set_reorder_period(self.h, C.c_int32(reorder_period)))
        if actors.route_table_off is not None:
            pass   # a partitioned world with explicit routes: the table goes in after griddy_mg_configure
        elif actors.route_off is None and net.nbr_off is not None and not net.grid_n:
            # find-route on the device: the lane graph of route.scm:19-40
            self._chk(L.griddy_set_route_graph(self.h, _p(net.nbr_off), _p(net.nbr), _p(net.cost_out), _p(net.midpoint)))
        elif actors.route_off is None:
            if not net.grid_n:
                raise ValueError("actors carry no routes and the network has neither a route graph nor a grid routing table")
            self._chk(L.griddy_set_routing_grid(self.h, C.c_int32(net.grid_n), _p(net.lane_from_j),
                                                _p(net.lane_to_j), _p(net.turn4)))
        self.partition = partition
        if partition is not None:
            rank, world, owner, mig_cap, extra_slots = partition
            lens = actors.agenda_off[1:] - actors.agenda_off[:-1]
            max_items = max(int(lens.max()) if len(lens) else 0, int(max_agenda_items or 0), 1)
            self._owner = np.ascontiguousarray(owner, np.int32)
            self._chk(L.griddy_mg_configure(self.h, C.c_int32(rank), C.c_int32(world), _p(self._owner), _p(net.lane_in),
                                            C.c_int32(mig_cap), C.c_int64(extra_slots), C.c_int32(max_items)))
        self.upload_actors(actors)
        if partition is not None:
            self.gids = np.ascontiguousarray(gids, np.int32)
            self._chk(L.griddy_mg_set_global_ids(self.h, _p(self.gids)))

    def upload_actors(self, a: FlatActors):
        self.actors = a
        if a.route_table_off is not None:
            self._chk(lib().griddy_mg_set_routes(self.h, C.c_int64(len(a.route_table_off) - 1), _p(a.route_table_off),
                                                 _p(a.route_table_lane), C.c_int64(len(a.item_route)), _p(a.item_route)))
        if a.dest_lane is not None:   # destinations resolved by the generator (needed on a tiled network)
            self._chk(lib().griddy_set_agenda_destinations(self.h, C.c_int64(a.n_agenda_items), _p(a.dest_lane), _p(a.dest_dir),
                                                           _p(a.dest_from_j), _p(a.dest_to_j)))
        self._chk(lib().griddy_upload_actors(self.h, C.c_int64(a.n), _p(a.loc_kind), _p(a.container),
                                             _p(a.pos), _p(a.side), _p(a.speed), _p(a.speed_dt),
                                             _p(a.agenda_off), _p(a.item_kind), _p(a.item_sleep),
                                             _p(a.item_seg), _p(a.item_side), _p(a.item_pos),
                                             _p(a.route_off), _p(a.route_lane)))

    def _chk(self, rc):
        if rc:
            raise GriddyCudaError(rc, lib().griddy_last_error(self.h).decode())

    def step(self, n_ticks: int = 1):
        self._chk(lib().griddy_step(self.h, C.c_int64(n_ticks)))

    def step_async(self, n_ticks: int = 1):
        """Enqueue n_ticks ticks without waiting for the device (griddy_step_async)."""
        self._chk(lib().griddy_step_async(self.h, C.c_int64(n_ticks)))

    def timed_reorder(self) -> float:
        """Reorder the slots now; device ms (griddy_reorder_now)."""
        ms = C.c_double(0.0)
        self._chk(lib().griddy_reorder_now(self.h, C.byref(ms)))
        return float(ms.value)

    @property
    def reorders(self) -> int:
        return int(lib().griddy_reorders(self.h))

    def profile_step(self, n_ticks: int = 1) -> dict:
        """Per-kernel device ms summed over n_ticks (events around every launch)."""
        ms = (C.c_double * 5)(0.0, 0.0, 0.0, 0.0, 0.0)
        self._chk(lib().griddy_profile_step(self.h, C.c_int64(n_ticks), ms))
        return {"k_advance": ms[0], "k_slow": ms[4], "k_unlink": ms[1], "k_link": ms[2], "reorder": ms[3]}

    @property
    def last_step_ms(self) -> float:
        return float(lib().griddy_last_step_ms(self.h))

    @property
    def last_frame_ms(self) -> float:
        return float(lib().griddy_last_frame_ms(self.h))

    @property
    def last_copy_ms(self) -> float:
        return float(lib().griddy_last_copy_ms(self.h))

    @property
    def tick(self) -> int:
        return int(lib().griddy_tick(self.h))

    def count_alive(self) -> int:
        n = C.c_int64(0)
        self._chk(lib().griddy_count_alive(self.h, C.byref(n)))
        return int(n.value)

    def digest(self) -> dict:
        """Partition-independent digest of this context's living actors (griddy_state_digest; digest.py)."""
        out = (C.c_uint64 * 3)(0, 0, 0)
        self._chk(lib().griddy_state_digest(self.h, out))
        return {"sum": f"{out[0]:016x}", "xor": f"{out[1]:016x}", "alive": int(out[2])}

    @property
    def slots_in_use(self) -> int:
        return int(lib().griddy_slots_in_use(self.h))

    @property
    def kernel_launches(self) -> int:
        return int(lib().griddy_kernel_launches(self.h))

    def state(self, with_order: bool = True) -> ActorState:
        n = self.actors.n
        u8 = lambda: np.zeros(n, np.uint8)
        i32 = lambda: np.zeros(n, np.int32)
        alive, loc_kind, side, rs = u8(), u8(), u8(), u8()
        container, agenda_cur, sleep_left, route_head, order = i32(), i32(), i32(), i32(), i32()
        pos = np.zeros(n, np.float64)
        n_order = C.c_int64(0)
        self._chk(lib().griddy_download_actors(self.h, _p(alive), _p(loc_kind), _p(container), _p(side),
                                               _p(pos), _p(agenda_cur), _p(sleep_left), _p(rs),
                                               _p(route_head), _p(order) if with_order else None,
                                               C.byref(n_order) if with_order else None))
        return ActorState(alive, loc_kind, container, side, pos, agenda_cur, sleep_left, rs, route_head,
                          order[: n_order.value].copy())

    def local_state(self, with_order: bool = True) -> dict:
        """This rank's slots (griddy_download_local): arrays by slot, plus gid and the local order."""
        cap = self.slots_in_use
        u8 = lambda: np.zeros(cap, np.uint8)
        i32 = lambda: np.zeros(cap, np.int32)
        out = {"gid": i32(), "alive": u8(), "loc_kind": u8(), "container": i32(), "side": u8(),
               "pos": np.zeros(cap, np.float64), "agenda_cur": i32(), "sleep_left": i32(), "route_status": u8(),
               "route_head": i32(), "order": i32()}
        n_local, n_order = C.c_int64(0), C.c_int64(0)
        self._chk(lib().griddy_download_local(
            self.h, C.c_int64(cap), C.byref(n_local), _p(out["gid"]), _p(out["alive"]), _p(out["loc_kind"]),
            _p(out["container"]), _p(out["side"]), _p(out["pos"]), _p(out["agenda_cur"]), _p(out["sleep_left"]),
            _p(out["route_status"]), _p(out["route_head"]), _p(out["order"]) if with_order else None,
            C.byref(n_order) if with_order else None))
        out["order"] = out["order"][: n_order.value]
        return out

    def positions(self, f32: bool = False):
        """(get-pos actor) of every slot (spatial.scm:125-149): xy [n, 2] (NaN where the slot holds no living
        actor) and the actor id per slot (-1: none).  Needs a network with `geom`."""
        cap = max(self.slots_in_use, 1)
        xy = np.zeros((cap, 2), np.float32 if f32 else np.float64)
        gid = np.zeros(cap, np.int32)
        n = C.c_int64(0)
        self._chk(lib().griddy_download_positions(self.h, C.c_int64(cap), C.byref(n), None if f32 else _p(xy),
                                                  _p(xy) if f32 else None, _p(gid)))
        return xy[: n.value], gid[: n.value]

    def frame_begin(self, buf: "PinnedFrame"):
        """Start the asynchronous read-back of a frame (float x, y per slot) into pinned host memory."""
        self._chk(lib().griddy_positions_begin(self.h, C.c_void_p(buf.ptr), C.c_int64(buf.capacity)))
        self._frames.append(buf)

    def frame_end(self) -> np.ndarray:
        """Wait for the oldest frame begun; returns its [n, 2] float32 view of the pinned buffer."""
        n = C.c_int64(0)
        self._chk(lib().griddy_positions_end(self.h, C.byref(n)))
        buf = self._frames.pop(0)
        return buf.array[: n.value]

    def delta_begin(self, buf: "DeltaFrame"):
        """Start the read-back of a delta frame: the slots whose drawn position changed since the frame before."""
        self._chk(lib().griddy_delta_begin(self.h, C.c_void_p(buf.ptr), C.c_int64(buf.capacity)))
        self._frames.append(buf)

    def delta_end(self) -> "DeltaFrame":
        """Wait for the oldest delta frame begun; returns its buffer (n_slots, n_changed, key, apply())."""
        ns, nc = C.c_int64(0), C.c_int64(0)
        self._chk(lib().griddy_delta_end(self.h, C.byref(ns), C.byref(nc)))
        buf = self._frames.pop(0)
        buf.n_slots, buf.n_changed = int(ns.value), int(nc.value)
        return buf

    def routes(self):
        """(route_off, route_lane) the world was uploaded with: the caller's, or those find-route computed on the device."""
        n = C.c_int64(0)
        self._chk(lib().griddy_download_routes(self.h, None, None, C.c_int64(0), C.byref(n)))
        off = np.zeros(self.actors.n_agenda_items + 1, np.int64)
        lanes = np.zeros(max(int(n.value), 1), np.int32)
        self._chk(lib().griddy_download_routes(self.h, _p(off), _p(lanes), C.c_int64(len(lanes)), C.byref(n)))
        return off, lanes[: n.value]

    def occupancy(self):
        """Actors per container and per junction, one entry per item of `net` (compact on a tile)."""
        lane_count = np.zeros(self.net.n_items, np.int32)
        junction_count = np.zeros(self.net.n_items, np.int32)
        self._chk(lib().griddy_download_occupancy(self.h, _p(lane_count), _p(junction_count)))
        return lane_count, junction_count

    def close(self):
        if getattr(self, "h", None):
            lib().griddy_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PinnedFrame:
    """Page-locked host memory for one frame of `capacity` float pairs (griddy_host_alloc)."""

    def __init__(self, capacity: int):
        self.capacity = int(capacity)
        self.ptr = lib().griddy_host_alloc(C.c_int64(8 * max(self.capacity, 1)))
        if not self.ptr:
            raise GriddyCudaError(5, "griddy_host_alloc failed")
        self.array = np.ctypeslib.as_array((C.c_float * (2 * self.capacity)).from_address(self.ptr)).reshape(-1, 2)

    def free(self):
        if self.ptr:
            self.array = None
            lib().griddy_host_free(C.c_void_p(self.ptr))
            self.ptr = None


def decode_delta(raw: np.ndarray):
    """(slots, xy) of a delta frame held in the uint8 array `raw`: the changed slots, ascending, and their values."""
    T = DeltaFrame.TILE
    h = raw[:64].view(np.int64)
    nb_cap = (int(h[6]) + T - 1) // T
    blocks = int(h[5])
    mask_off, off_off = 64, 64 + nb_cap * 128
    xy_off = (off_off + nb_cap * 4 + 63) // 64 * 64
    mask = raw[mask_off: mask_off + blocks * 128]
    first = raw[off_off: off_off + blocks * 4].view(np.uint32).astype(np.int64)
    vals = raw[xy_off: xy_off + 8 * int(h[2])].view(np.float32).reshape(-1, 2)
    bits = np.unpackbits(mask, bitorder="little")
    slots = np.flatnonzero(bits)
    per_block = bits.reshape(blocks, T).sum(axis=1).astype(np.int64)
    start = np.cumsum(per_block) - per_block           # changed slots before each block, in slot order
    blk = slots // T
    # value index of the k-th changed slot overall = first[its block] + its rank inside the block
    return slots, vals[first[blk] + (np.arange(len(slots), dtype=np.int64) - start[blk])]


class DeltaFrame:
    """Page-locked host memory for one delta frame over `capacity` slots (include/griddy_cuda.h, "Delta frames")."""
    TILE = 1024

    def __init__(self, capacity: int):
        self.capacity = int(capacity)
        self.bytes = int(lib().griddy_delta_frame_bytes(C.c_int64(self.capacity)))
        self.ptr = lib().griddy_host_alloc(C.c_int64(self.bytes))
        if not self.ptr:
            raise GriddyCudaError(5, "griddy_host_alloc failed")
        self.raw = np.ctypeslib.as_array((C.c_uint8 * self.bytes).from_address(self.ptr))
        self.n_slots = self.n_changed = 0

    @property
    def header(self) -> np.ndarray:
        return self.raw[:64].view(np.int64)

    @property
    def key(self) -> bool:
        return bool(self.header[4])

    @property
    def bytes_copied(self) -> int:
        """What griddy_delta_end moved over PCIe for this frame: mask, block table and the changed values."""
        blocks = (self.n_slots + self.TILE - 1) // self.TILE
        return blocks * (self.TILE // 32) * 4 + blocks * 4 + 8 * self.n_changed

    def apply(self, xy32: np.ndarray):
        """Patch the persistent [capacity, 2] float32 array with this frame (griddy_delta_apply)."""
        assert xy32.dtype == np.float32 and xy32.flags.c_contiguous
        rc = lib().griddy_delta_apply(C.c_void_p(self.ptr), _p(xy32), C.c_int64(xy32.shape[0]))
        if rc:
            raise GriddyCudaError(rc, "griddy_delta_apply: malformed frame or array too small")

    def decode(self):
        """(slots, xy) of this frame in numpy — an independent restatement of the layout, for tests."""
        return decode_delta(self.raw)

    def free(self):
        if self.ptr:
            self.raw = None
            lib().griddy_host_free(C.c_void_p(self.ptr))
            self.ptr = None


def connect_ipc(sim: "Simulation"):
    """One process per GPU: every rank exports its window to each neighbour (cudaIpc handle + the offsets of that
    neighbour's places in it), the blobs travel through torch.distributed (any backend), every rank maps its
    neighbours' windows.  After the closing barrier the ranks call step() in lock-step; the data plane is peer
    stores and flags between the kernels, with no message and no collective."""
    import torch.distributed as dist
    rank = sim.partition[0]
    nbs = np.zeros(8, np.int32)
    n = C.c_int32(0)
    sim._chk(lib().griddy_mg_neighbors(sim.h, _p(nbs), C.byref(n)))
    mine = {}
    for q in nbs[: n.value].tolist():
        h = np.zeros(128, np.uint8)
        sim._chk(lib().griddy_mg_ipc_export(sim.h, C.c_int32(q), _p(h)))
        mine[q] = h.tobytes()
    everyone = [None] * dist.get_world_size()
    dist.all_gather_object(everyone, mine)
    for q in nbs[: n.value].tolist():
        h = np.frombuffer(everyone[q][rank], np.uint8).copy()      # q's window, and my places in it
        sim._chk(lib().griddy_mg_ipc_import(sim.h, C.c_int32(q), _p(h)))
    dist.barrier()


class PartitionedSimulation:
    """All ranks of a partitioned world as contexts of this process (griddy_mg_connect_local): the
    executable form of the multi-GPU path for tests — on one GPU, or spread over the visible ones."""

    def __init__(self, net: FlatNetwork, actors: FlatActors, owner: np.ndarray, world: int, devices=None,
                 mig_cap: int = 4096, extra_slots: int | None = None, reorder_period: int | None = None, tiles=None, **kw):
        """tiles: per rank the (range_beg, range_end) id ranges it holds — every rank then uploads only that part of
        `net` (FlatNetwork.tile_of) and resolves its actors' destinations on the host."""
        self.net, self.actors, self.world = net, actors, world
        owner = np.ascontiguousarray(owner, np.int32)
        extra = actors.n if extra_slots is None else extra_slots
        lens = actors.agenda_off[1:] - actors.agenda_off[:-1]
        kw.setdefault("max_agenda_items", int(lens.max()) if len(lens) else 1)
        if tiles is not None and actors.dest_lane is None:
            actors = with_destinations(net, actors)
        self.ranks = []
        for r in range(world):
            mine = np.flatnonzero(owner[actors.container] == r)
            part = net if tiles is None else net.tile_of(tiles[r][0], tiles[r][1], owner)
            self.ranks.append(Simulation(part, subset_actors(actors, mine), device=devices[r] if devices else -1,
                                         reorder_period=reorder_period,
                                         partition=(r, world, owner if tiles is None else part.owner, mig_cap, extra),
                                         gids=mine, **kw))
        self._hs = (C.c_void_p * world)(*[s.h for s in self.ranks])
        rc = lib().griddy_mg_connect_local(self._hs, C.c_int32(world))
        if rc:
            raise GriddyCudaError(rc, "griddy_mg_connect_local failed")

    def step(self, n_ticks: int = 1):
        rc = lib().griddy_mg_step_local(self._hs, C.c_int32(self.world), C.c_int64(n_ticks))
        if rc:
            msgs = [lib().griddy_last_error(s.h).decode() for s in self.ranks]
            raise GriddyCudaError(rc, " | ".join(m for m in msgs if m))

    @property
    def tick(self) -> int:
        return self.ranks[0].tick

    def state(self) -> ActorState:
        """The world's state by global actor id, merged from every rank's read-back."""
        return merge_local_states([s.local_state() for s in self.ranks], self.actors.n)

    def migrants_seen(self) -> int:
        return sum(s.slots_in_use for s in self.ranks)


def subset_actors(a: FlatActors, idx: np.ndarray) -> FlatActors:
    """The actors `idx` (ascending) with their agendas re-packed."""
    beg, end = a.agenda_off[idx], a.agenda_off[idx + 1]
    lens = end - beg
    off = np.zeros(len(idx) + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    items = np.repeat(beg - off[:-1], lens) + np.arange(off[-1]) if len(idx) else np.zeros(0, np.int64)
    routes = {}
    if a.route_off is not None:   # explicit routes: every rank gets the whole table, its items name their routes in it
        routes = dict(route_table_off=a.route_off, route_table_lane=a.route_lane, item_route=items.astype(np.int32))
    dest = {} if a.dest_lane is None else dict(dest_lane=a.dest_lane[items], dest_dir=a.dest_dir[items],
                                               dest_from_j=a.dest_from_j[items], dest_to_j=a.dest_to_j[items])
    return FlatActors(a.loc_kind[idx], a.container[idx], a.pos[idx], a.side[idx], a.speed[idx], a.speed_dt[idx], off,
                      a.item_kind[items], a.item_sleep[items], a.item_seg[items], a.item_side[items], a.item_pos[items], **dest,
                      **routes)


def with_destinations(net: FlatNetwork, a: FlatActors) -> FlatActors:
    """`a` with every travel-to destination resolved from the WHOLE network `net`, as begin-route$ would
    (off-road->on-road, location.scm:49-57; get-outer-lane, location.scm:16-24): what griddy_set_agenda_destinations takes."""
    travel = a.item_kind == 1
    seg = np.where(travel, a.item_seg, 0)
    lane = np.where(travel, np.where(a.item_side == 1, net.seg_outer_back[seg], net.seg_outer_forw[seg]), -1).astype(np.int32)
    ok = lane >= 0
    l0 = np.where(ok, lane, 0)
    pick = lambda arr, fill: np.where(ok, arr[l0], fill) if arr is not None else np.full(len(lane), fill)
    out = FlatActors(a.loc_kind, a.container, a.pos, a.side, a.speed, a.speed_dt, a.agenda_off, a.item_kind, a.item_sleep,
                     a.item_seg, a.item_side, a.item_pos, a.route_off, a.route_lane, dest_lane=lane,
                     dest_dir=pick(net.lane_dir, 0).astype(np.uint8), dest_from_j=pick(net.lane_from_j, -1).astype(np.int32),
                     dest_to_j=pick(net.lane_to_j, -1).astype(np.int32))
    return out


def merge_local_states(parts, n_total: int) -> ActorState:
    z8, z32 = (lambda: np.zeros(n_total, np.uint8)), (lambda: np.zeros(n_total, np.int32))
    st = ActorState(z8(), z8(), z32(), z8(), np.zeros(n_total, np.float64), z32(), z32(), z8(), z32(),
                    np.zeros(0, np.int32))
    order_gid, order_container = [], []
    for p in parts:
        live = p["alive"].astype(bool)
        g = p["gid"][live]
        for f in ("alive", "loc_kind", "container", "side", "pos", "agenda_cur", "sleep_left", "route_status", "route_head"):
            getattr(st, f)[g] = p[f][live]
        order_gid.append(p["gid"][p["order"]])
        order_container.append(p["container"][p["order"]])
    og, oc = np.concatenate(order_gid), np.concatenate(order_container)
    # a container belongs to one rank: the world's get-actors order is the ranks' orders merged by
    # descending container id (stable, so each rank's order inside a container is kept)
    st.order = og[np.argsort(-oc.astype(np.int64), kind="stable")].astype(np.int32)
    return st
