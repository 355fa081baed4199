"""Shared builders for the parity tests."""
import numpy as np

import examples
from griddy_b200 import flatten, synth
from griddy_b200.flat import FlatActors


def world_to_flat(make_skeleton, add_actors, routes=True):
    world = make_skeleton()
    add_actors(world)
    net, idx = flatten.flatten_network(world)
    actors, objs = flatten.flatten_actors(world, idx, routes=routes)
    return net, actors


def simple_flat():
    return world_to_flat(examples.simple_make_skeleton, examples.simple_add_actors)


def triangle_flat():
    return world_to_flat(examples.triangle_make_skeleton, examples.triangle_add_actors)


def grid_flat(n, n_actors, agenda_len=2, seed=synth.DEFAULT_SEED):
    net = synth.grid_network(n)
    return net, synth.grid_actors(net, n_actors, seed=seed, agenda_len=agenda_len)


def crowded_grid(n, n_actors, agenda_len=4, seed=7, max_sleep=20, n_dest=None):
    """A grid workload squeezed so that lanes jam: short sleeps and few distinct destinations, which
    produces queues, junction waits and equal-key replacements (core.scm:173-178)."""
    net, a = grid_flat(n, n_actors, agenda_len=agenda_len, seed=seed)
    a.item_sleep[:] = a.item_sleep % (max_sleep + 1)
    if n_dest:
        segs = net.seg_reg[:: max(1, len(net.seg_reg) // n_dest)][:n_dest]
        travel = a.item_kind == 1
        a.item_seg[travel] = segs[np.arange(travel.sum()) % len(segs)]
    return net, a


def compare(sim, ref, where=""):
    ds, os_ = sim.state(), ref.state()
    bad = ds.diff(os_)
    assert not bad, f"{where}: device differs from oracle in {bad}"
    dl, dj = sim.occupancy()
    ol, oj = ref.occupancy()
    assert np.array_equal(dl, ol), f"{where}: per-container actor counts differ"
    assert np.array_equal(dj, oj), f"{where}: junction occupancy differs"
    return ds
