"""Parity of the CUDA path with the oracle, through the C ABI.  Integers bit-exact (lane id, agenda
index, sleep counter, route status/head, actor order, occupancy counts), fp64 positions within 1e-9
relative (ActorState.diff) — and in fact compared for exact equality as well, since every operation
is a correctly rounded IEEE one on both sides."""
import os

import numpy as np
import pytest

import examples
from griddy_b200 import core, device, digest
from griddy_b200.flat import FlatActors
from helpers import compare, crowded_grid, grid_flat, simple_flat, triangle_flat, world_to_flat
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu


def run_pair(net, actors, ticks, every=1, exact_pos=True):
    sim, ref = device.Simulation(net, actors), Oracle(net, actors)
    compare(sim, ref, "tick 0")
    done = 0
    while done < ticks:
        k = min(every, ticks - done)
        sim.step(k); ref.step(k)
        done += k
        ds = compare(sim, ref, f"tick {done}")
        if exact_pos:
            m = ds.alive.astype(bool)
            assert np.array_equal(ds.pos[m], ref.state().pos[m]), f"tick {done}: pos not bit-identical"
    return sim, ref


def test_simple_matches_known_answer():
    net, actors = simple_flat()
    sim = device.Simulation(net, actors)
    from test_oracle_kat import expected_simple
    _, traj = expected_simple()
    for t, (where, pos, rs, head) in enumerate(traj, start=1):
        sim.step(1)
        s = sim.state()
        assert (s.loc_kind[0], s.pos[0], s.route_status[0], s.route_head[0]) == (int(where == "on"), pos, rs, head), t
    run_pair(net, actors, 40)


def test_triangle_10k_ticks_tick_by_tick():
    net, actors = triangle_flat()
    run_pair(net, actors, 10_000)


@pytest.mark.parametrize("n,n_actors,agenda_len,seed", [(5, 80, 2, 1), (3, 200, 4, 2), (8, 3000, 4, 3)])
def test_procedural_grid(n, n_actors, agenda_len, seed):
    net, actors = grid_flat(n, n_actors, agenda_len=agenda_len, seed=seed)
    run_pair(net, actors, 1500, every=25)


@pytest.mark.parametrize("seed", [7, 8, 9])
def test_crowded_grid_with_replacements(seed):
    """Jammed lanes: queues, junction waits, several entrants per lane per tick, equal-key losers."""
    net, actors = crowded_grid(4, 800, n_dest=3, seed=seed)
    sim, ref = run_pair(net, actors, 600, every=1)
    assert ref.vanished > 0
    assert int(sim.state().alive.sum()) == actors.n - ref.vanished


def test_crowded_grid_long():
    net, actors = crowded_grid(6, 5000, n_dest=6, seed=21, agenda_len=6)
    run_pair(net, actors, 3000, every=50)


def test_explicit_routes_from_host_a_star():
    """Routes uploaded as lane lists (route.scm:48-51) instead of the grid routing table."""
    import random
    rng = random.Random(5)

    def add_actors(world):
        segs = core.get_static_items(world, core.RoadSegment)
        loc = lambda: core.LocationOffRoad(rng.choice(segs), rng.choice(["forw", "back"]), 0.25 + 0.5 * rng.random())
        for _ in range(60):
            a = core.Actor(max_speed=30 + rng.randrange(30))
            core.link(a, loc())
            core.agenda_push(a, ("sleep-for", rng.randrange(20)))
            core.agenda_push(a, ("travel-to", loc()))
            core.agenda_push(a, ("travel-to", loc()))
    net, actors = world_to_flat(lambda: examples.grid_make_skeleton(4), add_actors)
    assert actors.route_off is not None and len(actors.route_lane) > 0
    run_pair(net, actors, 1200, every=10)


@pytest.mark.parametrize("world,strips", [(2, 1), (3, 2)])
def test_partitioned_world_with_explicit_routes(world, strips):
    """Routes from the host's A* on a partitioned world: every rank holds the whole route table (griddy_mg_set_routes),
    an actor that crosses takes its items' route indices and its place in the current route along."""
    import random
    from griddy_b200 import synth
    rng = random.Random(11)
    n = 6

    def add_actors(w):
        segs = core.get_static_items(w, core.RoadSegment)
        loc = lambda: core.LocationOffRoad(rng.choice(segs), rng.choice(["forw", "back"]), 0.25 + 0.5 * rng.random())
        for _ in range(400):
            a = core.Actor(max_speed=30 + rng.randrange(30))
            core.link(a, loc())
            core.agenda_push(a, ("sleep-for", rng.randrange(12)))
            for _ in range(3):
                core.agenda_push(a, ("travel-to", loc()))
    net, actors = world_to_flat(lambda: examples.grid_make_skeleton(n), add_actors)
    assert actors.route_off is not None and len(actors.route_lane) > 0
    owner = synth.grid_owner(synth.grid_network(n), world, strips)   # (the generator's ids are the API mirror's: test_host.py)
    sim = device.PartitionedSimulation(net, actors, owner, world, reorder_period=7)
    ref = Oracle(net, actors)
    crossed = 0
    for t in range(700):
        sim.step(1); ref.step(1)
        ds, os_ = sim.state(), ref.state()
        bad = ds.diff(os_)
        assert not bad, f"world {world} tick {t + 1}: partitioned device with explicit routes differs from oracle in {bad}"
        m = ds.alive.astype(bool)
        assert np.array_equal(ds.pos[m], os_.pos[m]), f"tick {t + 1}: pos not bit-identical"
        crossed += int((owner[ds.container[m]] != owner[actors.container[m]]).sum() > 0)
    assert crossed > 0, "no actor ever crossed a partition boundary"
    assert digest.combine([r.digest() for r in sim.ranks]) == digest.commutative(os_)


def test_empty_world_and_empty_agendas():
    net, actors = grid_flat(3, 0)
    sim = device.Simulation(net, actors)
    sim.step(5)
    assert sim.state().order.size == 0
    net, actors = grid_flat(3, 40, agenda_len=0)
    sim, ref = run_pair(net, actors, 7)
    assert np.array_equal(sim.state().pos, actors.pos)
    # off-road lists reverse every tick (core.scm:169-171): after an odd count the order flips
    o0 = device.Simulation(net, actors).state().order
    assert not np.array_equal(o0, sim.state().order)


def test_on_road_idle_actors_are_obstacles_and_dedupe_on_upload():
    """Actors linked on-road with nothing to do never move but block followers (route.scm:68-74);
    equal keys at upload keep the later link! (core.scm:173-178)."""
    net, a = grid_flat(3, 30, agenda_len=2, seed=4)
    lane = int(net.seg_outer_forw[a.container[0]])
    seg = int(a.container[0])
    n_extra = 4
    cat = lambda x, y: np.concatenate([x, np.asarray(y, x.dtype)])
    # everyone starts beside the same lane, behind the parked actors
    a.container[:] = seg; a.side[:] = 0; a.pos[:] = np.linspace(0.05, 0.30, a.n)
    acts = FlatActors(cat(a.loc_kind, [1] * n_extra), cat(a.container, [lane] * n_extra),
                      cat(a.pos, [0.6, 0.6, 0.8, 0.95]), cat(a.side, [0] * n_extra),
                      cat(a.speed, [80.0] * n_extra), cat(a.speed_dt, [80 / 15] * n_extra),
                      cat(a.agenda_off, [a.agenda_off[-1]] * n_extra), a.item_kind, a.item_sleep % 5,
                      a.item_seg, a.item_side, a.item_pos)
    sim, ref = run_pair(net, acts, 400)
    s = sim.state()
    assert s.alive[a.n] == 0 and s.alive[a.n + 1] == 1
    assert s.pos[a.n + 2] == 0.8 and s.pos[a.n + 3] == 0.95


def test_arrive_behind_start_jumps_back():
    """Destination on the start lane but behind the actor: the route is ((arrive-at p)) and the
    actor lands on p in one tick, passing other residents (route.scm:81-85)."""
    net, a = grid_flat(3, 12, agenda_len=2, seed=6)
    seg = int(a.container[0])
    a.container[:] = seg; a.side[:] = 1
    a.pos[:] = np.linspace(0.30, 0.74, a.n)
    a.item_sleep[:] = np.arange(a.n * 2)[: len(a.item_sleep)] % 3
    a.item_seg[:] = seg; a.item_side[:] = 1
    a.item_pos[:] = np.where(np.arange(len(a.item_pos)) % 4 == 1, 0.9, 0.28)   # half ahead, half behind
    run_pair(net, a, 200)


def test_begin_route_collision_same_position():
    """Two actors beside the same lane at the same pos-param begin their routes in the same tick:
    the later one in the segment's list replaces the other (core.scm:173-178)."""
    net, a = grid_flat(3, 6, agenda_len=2, seed=10)
    a.container[:] = a.container[0]; a.side[:] = 0
    a.pos[:] = [0.4, 0.4, 0.5, 0.5, 0.5, 0.6]
    a.item_sleep[:] = 0
    a.item_sleep[0::2] = [3, 3, 4, 4, 4, 2]
    sim, ref = run_pair(net, a, 100)
    assert ref.vanished == 3


def test_bad_state_is_reported():
    net, a = grid_flat(3, 2, agenda_len=2)
    lane = int(net.seg_outer_forw[a.container[0]])
    a.loc_kind[0] = 1; a.container[0] = lane; a.item_sleep[0] = 1   # on-road actor that will travel-to
    sim = device.Simulation(net, a)
    with pytest.raises(device.GriddyCudaError) as e:
        sim.step(10)
    assert e.value.code == 4 and "begin-route$" in str(e.value)
    with pytest.raises(RuntimeError):
        o = Oracle(net, a); o.step(10)


def test_bad_arguments_are_rejected():
    net, a = grid_flat(3, 2)
    a.container[0] = 0    # a junction
    with pytest.raises(device.GriddyCudaError) as e:
        device.Simulation(net, a)
    assert e.value.code == 1


def test_radix_sort_matches_stable_argsort():
    """The reorder's hand-written LSD radix sort against numpy's stable sort, ragged sizes included."""
    import ctypes as C
    L = device.lib()
    rng = np.random.default_rng(11)
    for n, bits in [(1, 8), (4095, 16), (4096, 24), (4097, 41), (100_000, 41), (1_000_003, 33)]:
        keys = rng.integers(0, 1 << bits, size=n, dtype=np.uint64)
        if n > 10:
            keys[: n // 3] = keys[n // 2: n // 2 + n // 3]      # plenty of duplicates: stability matters
        vals = np.arange(n, dtype=np.uint32)
        want = np.argsort(keys, kind="stable").astype(np.uint32)
        k, v = keys.copy(), vals.copy()
        rc = L.griddy_debug_sort_pairs(k.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p),
                                       C.c_int64(n), C.c_int32(bits))
        assert rc == 0
        assert np.array_equal(v, want), (n, bits)
        assert np.array_equal(k, keys[want])


@pytest.mark.parametrize("period", [0, 1, 7])
def test_reorder_period_does_not_change_results(period):
    """Slots are re-sorted every `period` ticks (0 = never): state, order and occupancy stay identical
    to the oracle's, and dead actors are dropped from the slots."""
    net, actors = crowded_grid(4, 800, n_dest=3, seed=8)
    sim, ref = device.Simulation(net, actors, reorder_period=period), Oracle(net, actors)
    for t in range(300):
        sim.step(1); ref.step(1)
        compare(sim, ref, f"period {period} tick {t + 1}")
    assert ref.vanished > 0
    if period:
        sim.step(period); ref.step(period)
        assert sim.slots_in_use <= actors.n - ref.vanished + 8   # only a few ghosts awaiting removal
    else:
        assert sim.slots_in_use == actors.n


def test_default_locality_keys_without_grid():
    """Networks without uploaded locality keys (registration index as key), explicit routes, reorder on."""
    net, actors = triangle_flat()
    sim, ref = device.Simulation(net, actors, reorder_period=2), Oracle(net, actors)
    for t in range(60):
        sim.step(1); ref.step(1)
        compare(sim, ref, f"tick {t + 1}")


@pytest.mark.parametrize("world,n,n_actors,strips,tiled", [(2, 6, 2500, 1, False), (4, 8, 4000, 1, False), (3, 5, 600, 1, False),
                                                           (3, 9, 3000, 2, False), (2, 8, 3000, 4, False),
                                                           (2, 8, 2500, 1, True), (4, 12, 6000, 1, True), (3, 12, 5000, 2, True)])
def test_partitioned_world_matches_oracle(world, n, n_actors, strips, tiled):
    """The multi-GPU path (spatial partition, halo + migrant hand-over every tick) with all ranks as
    contexts of this process: state, order and population identical to the oracle's single world.
    tiled: every rank uploads only its own rows of the network plus the ring next to them."""
    from griddy_b200 import synth
    net, actors = crowded_grid(n, n_actors, n_dest=max(4, n), seed=31 + world, agenda_len=4)
    owner = synth.grid_owner(net, world, strips)   # strips > 1: interleaved strips, every rank borders every other
    assert len(set(owner.tolist())) == world
    tiles = [synth.tile_ranges(n, *synth.tile_rows(n, world, strips, r)) for r in range(world)] if tiled else None
    if tiled:
        held = [int((e - b).sum()) for b, e in tiles]
        assert max(held) < net.n_items, "every rank holds the whole network: nothing is tiled"
    sim = device.PartitionedSimulation(net, actors, owner, world, reorder_period=5, tiles=tiles)
    ref = Oracle(net, actors)
    moved = 0
    for t in range(400):
        sim.step(1); ref.step(1)
        ds, os_ = sim.state(), ref.state()
        bad = ds.diff(os_)
        assert not bad, f"world {world} tick {t + 1}: partitioned device differs from oracle in {bad}"
        m = ds.alive.astype(bool)
        assert np.array_equal(ds.pos[m], os_.pos[m]), f"tick {t + 1}: pos not bit-identical"
        moved += int((owner[ds.container[m]] != owner[actors.container[m]]).sum() > 0)
    assert moved > 0, "no actor ever crossed a partition boundary"
    assert ref.vanished > 0
    assert digest.combine([r.digest() for r in sim.ranks]) == digest.commutative(os_)
    if tiled:   # occupancy of a tile: compact, this rank's items only
        ol, oj = ref.occupancy()
        for r, s_ in enumerate(sim.ranks):
            idx = np.concatenate([np.arange(b, e) for b, e in zip(*tiles[r])])
            own = s_.net.owner == r
            dl, dj = s_.occupancy()
            assert np.array_equal(dl[own], ol[idx][own]) and np.array_equal(dj[own], oj[idx][own]), f"rank {r}: tile occupancy differs"


def test_migrants_do_not_exhaust_slots_or_the_agenda_pool():
    """Actors that cross ranks again and again take a fresh slot and fresh room for their agenda items at every
    arrival; a reorder gives both back.  Far more cumulative arrivals than extra_slots (and long agendas, of which
    only the items not yet popped travel) must therefore run without an overflow, and stay equal to the oracle."""
    from griddy_b200 import synth
    net, actors = crowded_grid(6, 1500, n_dest=6, seed=5, agenda_len=12, max_sleep=3)
    owner = synth.grid_owner(net, 2, 3)   # six strips: an actor crosses often
    sim = device.PartitionedSimulation(net, actors, owner, 2, reorder_period=8, extra_slots=700)
    ref = Oracle(net, actors)
    arrivals = 0
    for k in range(40):
        before = [s_.slots_in_use for s_ in sim.ranks]
        sim.step(50); ref.step(50)
        ds, os_ = sim.state(), ref.state()
        assert not ds.diff(os_), f"tick {50 * (k + 1)}"
        arrivals += sum(max(0, s_.slots_in_use - b) for s_, b in zip(sim.ranks, before))
    assert ref.vanished < actors.n


def run_ipc_check(world, *extra, gpus_needed=1, timeout=900):
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < gpus_needed:
        pytest.skip(f"needs {gpus_needed} GPUs (run with gpurun --gpus {gpus_needed})")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(29517 + world),
                        os.path.join(root, "tests", "mg_ipc_check.py"), *extra],
                       capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0 and "mg_ipc_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_partitioned_world_over_ipc_windows_sharing_one_gpu():
    """The production transport — one process per rank, windows shared through cudaIpc, flags between kernels of
    DIFFERENT processes — on whatever GPUs the box has: with one GPU the ranks share it (time-sliced)."""
    run_ipc_check(2, "--grid", "8", "--actors", "4000", "--ticks", "100", "--every", "25", "--tiled")


def test_partitioned_world_over_ipc_windows_two_gpus():
    """Two ranks on two GPUs (NVLink peer stores), interleaved strips, tiled network, against the oracle."""
    run_ipc_check(2, "--strips", "2", "--tiled", gpus_needed=2)
    run_ipc_check(2, "--strips", "3", "--tiled", "--async-steps", "--every", "100", gpus_needed=2)
