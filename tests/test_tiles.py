"""Host side of the tiled network (no GPU): a rank generates only the junction rows it owns plus the ring next to
them, from closed forms; what it gets must be exactly the corresponding part of the whole grid, and the exchange
plans that two neighbours derive from their own tiles must be the ones the whole network gives."""
import ctypes as C

import numpy as np
import pytest

from griddy_b200 import device, synth


def p(a):
    return a.ctypes.data_as(C.c_void_p)


NET_FIELDS = ("kind", "lane_len", "lane_dir", "lane_segment", "lane_out", "lane_junction", "seg_outer_forw", "seg_outer_back",
              "lane_in", "lane_from_j", "lane_to_j", "geom")


@pytest.mark.parametrize("n,world,strips", [(5, 2, 1), (8, 2, 2), (9, 3, 1), (12, 4, 1), (16, 8, 2), (7, 1, 1)])
def test_a_tile_is_the_corresponding_part_of_the_whole_grid(n, world, strips):
    whole = synth.grid_network(n, geometry=True)
    owner = synth.grid_owner(whole, world, strips)
    held_total = 0
    for rank in range(world):
        rb, re = synth.tile_rows(n, world, strips, rank)
        tile = synth.grid_tile(n, rb, re, world, strips, geometry=True)
        gb, ge = synth.tile_ranges(n, rb, re)
        assert np.array_equal(tile.range_beg, gb) and np.array_equal(tile.range_end, ge) and tile.n_global == whole.n_items
        want = whole.tile_of(gb, ge, owner)
        for f in NET_FIELDS:
            assert np.array_equal(getattr(tile, f), getattr(want, f)), (rank, f)
        assert np.array_equal(tile.owner, want.owner), rank
        # the routing table: complete for every lane whose end junction hangs off the tile's rows
        rows = np.zeros(n, bool)
        for b, e in zip(rb, re):
            rows[b:e] = True
        lanes = np.flatnonzero((tile.kind == 2) & rows[np.maximum(tile.lane_to_j, 0) // n])
        assert np.array_equal(tile.turn4[lanes], want.turn4[lanes]), rank
        # everything this rank owns is in the tile, and local() maps ids both ways
        mine = np.flatnonzero(owner == rank)
        assert np.all(tile.local(mine) >= 0)
        idx = np.concatenate([np.arange(b, e) for b, e in zip(gb, ge)])
        assert np.array_equal(tile.local(idx), np.arange(len(idx)))
        outside = np.setdiff1d(np.arange(whole.n_items), idx)
        assert np.all(tile.local(outside) == -1)
        held_total += tile.n_items
    if strips == 1 and n // world >= 3 and world > 1:
        assert held_total < 2 * whole.n_items   # tiles overlap by their rings only


@pytest.mark.parametrize("n,world,strips", [(6, 2, 1), (8, 2, 2), (9, 3, 2), (12, 4, 1), (16, 8, 2)])
def test_neighbours_derive_the_same_exchange_plan_from_their_own_tiles(n, world, strips):
    L = device.lib()
    whole = synth.grid_network(n)
    owner = synth.grid_owner(whole, world, strips)

    def plan(net, own, a, b):
        lanes, junctions = np.empty(net.n_items, np.int32), np.empty(net.n_items, np.int32)
        nl, nj, nf = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        rb = net.range_beg if net.tiled else np.zeros(1, np.int64)
        re = net.range_end if net.tiled else np.array([net.n_items], np.int64)
        rc = L.griddy_mg_plan_tile(C.c_int64(net.n_global or net.n_items), C.c_int32(len(rb)), p(rb), p(re), p(net.kind),
                                   p(net.lane_in), p(net.lane_out), p(net.lane_junction), p(own), C.c_int32(a), C.c_int32(b),
                                   p(lanes), C.byref(nl), p(junctions), C.byref(nj), C.byref(nf))
        assert rc == 0
        return lanes[: nl.value].tolist(), junctions[: nj.value].tolist(), nf.value

    tiles = [synth.grid_tile(n, *synth.tile_rows(n, world, strips, r), world, strips) for r in range(world)]
    crossing = 0
    for a in range(world):
        for b in range(world):
            if a == b:
                continue
            want = plan(whole, owner, a, b)
            assert plan(tiles[a], tiles[a].owner, a, b) == want, (a, b, "as seen by the owner")
            assert plan(tiles[b], tiles[b].owner, a, b) == want, (a, b, "as seen by the reader")
            crossing += len(want[0])
    assert crossing > 0


@pytest.mark.parametrize("n,n_actors,agenda_len", [(5, 300, 2), (9, 2000, 4), (16, 5000, 3)])
def test_actors_from_closed_forms_equal_the_generator_with_the_network(n, n_actors, agenda_len):
    net = synth.grid_network(n)
    a = device.with_destinations(net, synth.grid_actors(net, n_actors, agenda_len=agenda_len))
    b = synth.grid_actors_closed_form(n, n_actors, agenda_len=agenda_len)
    for f in ("loc_kind", "container", "pos", "side", "speed", "speed_dt", "agenda_off", "item_kind", "item_sleep", "item_seg",
              "item_side", "item_pos"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    travel = a.item_kind == 1
    for f in ("dest_lane", "dest_dir", "dest_from_j", "dest_to_j"):
        assert np.array_equal(getattr(a, f)[travel], getattr(b, f)[travel]), f
    for world, strips in ((2, 1), (3, 2)):
        owner = synth.grid_owner(net, world, strips)
        got = []
        for rank in range(world):
            part, ids = synth.grid_actors_of_rank(n, n_actors, world, strips, rank, agenda_len=agenda_len)
            assert np.array_equal(ids, synth.actor_ids_of_rank(net, n_actors, owner, rank))
            assert np.array_equal(part.container, a.container[ids]) and np.array_equal(part.pos, a.pos[ids])
            got.append(ids)
        assert np.array_equal(np.sort(np.concatenate(got)), np.arange(n_actors))
