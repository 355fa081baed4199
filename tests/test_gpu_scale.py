"""BASELINE.json configs at (or scaled towards) full size, through the C ABI."""
import json
import os

import numpy as np
import pytest

from griddy_b200 import device, digest
from helpers import compare, grid_flat
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu

SCALE_DIGESTS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scale_digests.json")


def check_against_scale_digests(name):
    """Step the device through a full-size world and compare, at every checkpoint, one SHA-256 per field of the
    whole state (actor-id order, pos-param bits, get-actors order, occupancy) and the device's own commutative
    digest with what the oracle produced for the same world (tests/golden/make_scale_digests.py)."""
    with open(SCALE_DIGESTS) as f:
        want = json.load(f)[name]
    net, actors = grid_flat(want["grid"], want["actors"], agenda_len=want["agenda_len"])
    sim = device.Simulation(net, actors)
    tick = 0
    for cp in sorted(int(k) for k in want["checkpoints"]):
        sim.step(cp - tick)
        tick = cp
        w = want["checkpoints"][str(cp)]
        assert sim.digest() == w["commutative"], f"{name} tick {cp}: device digest differs from the oracle's"
        st = sim.state()
        lane_count, junction_count = sim.occupancy()
        got = digest.sha_fields(st, lane_count, junction_count)
        bad = [k for k in w["sha"] if got[k] != w["sha"][k]]
        assert not bad, f"{name} tick {cp}: device differs from the oracle in {bad}"
        assert digest.commutative(st) == w["commutative"]
    return sim


def test_c3_full_size_against_oracle_digests():
    """configs[2] (256x256, 1M actors): every field bit for bit at ticks 100, 200, 300."""
    check_against_scale_digests("c3")


def test_c4_full_size_against_oracle_digests():
    """configs[3] (1024x1024, 16M actors, 4-item agendas — the bench workload): every field bit for bit at ticks
    100, 200, 300 and at tick 325, the end of the driver's timed window."""
    check_against_scale_digests("c4")


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


def test_c4_one_rank_vs_two_ranks_identical_state():
    """configs[3] shape (1 vs 2 B200), at a size the suite can afford: one world on one context and the
    same world partitioned over two ranks must agree bit for bit, including the get-actors order."""
    from griddy_b200 import synth
    net, actors = grid_flat(256, 1_000_000, agenda_len=4)
    one = device.Simulation(net, actors)
    two = device.PartitionedSimulation(net, actors, synth.grid_owner(net, 2), 2, mig_cap=1 << 20)
    for k in range(4):
        one.step(150); two.step(150)
        a, b = one.state(), two.state()
        assert not a.diff(b), f"tick {150 * (k + 1)}"
        m = a.alive.astype(bool)
        assert np.array_equal(a.pos[m], b.pos[m])
        # the device-side digests: the two ranks' add / xor up to the single world's
        assert digest.combine([r.digest() for r in two.ranks]) == one.digest() == digest.commutative(a)
    assert int(a.alive.sum()) < actors.n          # actors vanished, identically on both
    assert int((a.loc_kind[m] == 1).sum()) > 400_000


def test_c4_full_size_invariants():
    """configs[3] at full size (1024x1024, 16M actors): size-independent properties after 400 ticks."""
    net, actors = grid_flat(1024, 16_000_000, agenda_len=4)
    sim = device.Simulation(net, actors)
    sim.step(400)
    a = sim.state(with_order=False)
    alive = a.alive.astype(bool)
    assert sim.count_alive() == int(alive.sum())
    lane_count, junction_count = sim.occupancy()
    assert int(lane_count.sum()) == int(alive.sum())
    on = alive & a.loc_kind.astype(bool)
    jl = on & (net.kind[a.container] == 3)
    assert np.array_equal(np.bincount(net.lane_junction[a.container[jl]], minlength=net.n_items), junction_count)
    # no two living actors share a (lane, pos-param): bbtree keys are unique (core.scm:173-178)
    key = a.container[on].astype(np.int64) * (1 << 20)
    order = np.lexsort((a.pos[on], key))
    c, p = a.container[on][order], a.pos[on][order]
    assert not np.any((c[1:] == c[:-1]) & (p[1:] == p[:-1]))
    # within a lane nobody stands inside the tail-gate buffer of the actor ahead... except entrants,
    # so only the weaker property: positions are finite and the world is busy
    assert np.all(np.isfinite(a.pos[alive])) and int(on.sum()) > 8_000_000
    # off-road actors keep a pos-param in [0, 1]
    off = alive & ~a.loc_kind.astype(bool)
    assert np.all((a.pos[off] >= 0.0) & (a.pos[off] <= 1.0))
