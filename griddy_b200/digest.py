"""State digests of a world: partition-independent, computed the same way from any read-back.

`commutative(state)` restates in numpy what the device kernel k_digest (csrc/readback.cuh,
griddy_state_digest) computes per living actor and adds up: a 64-bit hash of the actor's id and of
everything griddy_download_actors reports about it.  Sums and xors commute, so the digest of a world
partitioned over several ranks is the sum / xor of the ranks' digests, and equal digests on 1 and N
GPUs mean equal states (up to hash collisions).

`sha_fields(state, ...)` is the strict form used by the full-size parity tests: one SHA-256 per field
over the arrays in actor-id order, plus the get-actors order and the occupancy counts.
"""
from __future__ import annotations

import hashlib

import numpy as np

from .flat import ActorState

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser (the one csrc/synth.cpp and k_digest use), on uint64 arrays."""
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return x ^ (x >> np.uint64(31))


def actor_hashes(gid, loc_kind, container, side, pos, agenda_cur, sleep_left, route_status, route_head) -> np.ndarray:
    u = lambda a: np.asarray(a).astype(np.int64).astype(np.uint64)
    u32 = lambda a: u(a) & np.uint64(0xFFFFFFFF)
    h = mix64(u(gid) + np.uint64(1))
    h = mix64(h ^ (u32(container) | (u(loc_kind) << np.uint64(32)) | (u(side) << np.uint64(33)) | (u(route_status) << np.uint64(34))))
    h = mix64(h ^ np.ascontiguousarray(pos, np.float64).view(np.uint64))
    h = mix64(h ^ (u32(agenda_cur) | (u32(sleep_left) << np.uint64(32))))
    return mix64(h ^ u32(route_head))


def commutative(st: ActorState) -> dict:
    """{"sum", "xor", "alive"} over the living actors of a read-back indexed by actor id."""
    m = st.alive.astype(bool)
    gid = np.flatnonzero(m)
    h = actor_hashes(gid, st.loc_kind[m], st.container[m], st.side[m], st.pos[m], st.agenda_cur[m], st.sleep_left[m],
                     st.route_status[m], st.route_head[m])
    with np.errstate(over="ignore"):
        s = int(np.add.reduce(h, dtype=np.uint64)) if len(h) else 0
    x = int(np.bitwise_xor.reduce(h)) if len(h) else 0
    return {"sum": f"{s:016x}", "xor": f"{x:016x}", "alive": int(m.sum())}


def combine(parts) -> dict:
    """The digest of a partitioned world from its ranks' digests."""
    s = x = n = 0
    for p in parts:
        s = (s + int(p["sum"], 16)) & 0xFFFFFFFFFFFFFFFF
        x ^= int(p["xor"], 16)
        n += int(p["alive"])
    return {"sum": f"{s:016x}", "xor": f"{x:016x}", "alive": n}


def sha_fields(st: ActorState, lane_count=None, junction_count=None) -> dict:
    """One SHA-256 per field (living actors only, in actor-id order; pos-params by their bits)."""
    m = st.alive.astype(bool)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    out = {"alive": sha(st.alive.astype(np.uint8)), "n_alive": int(m.sum())}
    for f in ActorState.INT_FIELDS[1:]:
        out[f] = sha(getattr(st, f)[m].astype(np.int32))
    out["pos"] = sha(st.pos[m].astype(np.float64))
    out["order"] = sha(np.asarray(st.order, np.int32))
    if lane_count is not None:
        out["lane_count"] = sha(np.asarray(lane_count, np.int32))
    if junction_count is not None:
        out["junction_count"] = sha(np.asarray(junction_count, np.int32))
    return out
