"""BASELINE.json configs at (or scaled towards) full size, through the C ABI."""
import numpy as np
import pytest

from griddy_b200 import device
from helpers import compare, grid_flat
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu


def test_c3_256_grid_1m_actors_against_oracle():
    """configs[2]: 256x256 grid, 1M actors.  Full state diffed against the oracle at 5 checkpoints."""
    net, actors = grid_flat(256, 1_000_000)
    sim, ref = device.Simulation(net, actors), Oracle(net, actors)
    for k in range(5):
        sim.step(60); ref.step(60)
        compare(sim, ref, f"tick {60 * (k + 1)}")


def test_c3_invariants_over_1000_ticks():
    """Size-independent properties at full size: conservation, junction occupancy consistent with
    a recount, no two living actors share a (lane, pos-param), determinism across runs."""
    net, actors = grid_flat(256, 1_000_000, seed=99)
    sims = [device.Simulation(net, actors) for _ in range(2)]
    for s in sims:
        s.step(1000)
    a, b = (s.state(with_order=False) for s in sims)
    assert not a.diff(b) and np.array_equal(a.pos, b.pos)
    lane_count, junction_count = sims[0].occupancy()
    alive = a.alive.astype(bool)
    assert int(lane_count.sum()) == int(alive.sum())
    on = alive & a.loc_kind.astype(bool)
    jl = on & (net.kind[a.container] == 3)
    recount = np.bincount(net.lane_junction[a.container[jl]], minlength=net.n_items)
    assert np.array_equal(recount, junction_count)
    keys = np.stack([a.container[on].astype(np.float64), a.pos[on]], axis=1)
    assert len(np.unique(keys, axis=0)) == int(on.sum())
    assert int(on.sum()) > 500_000   # the world is busy, not drained
