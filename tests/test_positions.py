"""Drawing read-back: (get-pos actor) of every actor (spatial.scm:125-149), which is all `draw` reads
(simulate.scm:155-160 -> draw.scm:74-88).

CPU: the oracle's restatement against `(get-pos actor)` values recorded by running the reference's own
spatial.scm under tests/golden/minischeme.py (every tick for simple / triangle, every 10th tick for the
grids), bit for bit; the generator's geometry against the API mirror's.
GPU: griddy_download_positions and the griddy_positions_begin / _end pipeline against the same golden
values and against the oracle on crowded synthetic grids, bit for bit in fp64; the float frame is the fp64
result rounded once."""
import numpy as np
import pytest

import examples
from griddy_b200 import flatten, synth
from griddy_b200.flat import K_JLANE, K_SEGLANE, K_SEGMENT
from helpers import crowded_grid
from oracle import oracle
from test_golden import FIXTURES, expected_states, load


def recorded_xy(s):
    return np.asarray(s["xy"], np.float64).reshape(-1, 2)


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_get_pos_matches_the_reference(name):
    fx, net, _ = load(name)
    geom = np.asarray(fx["network"]["geom"], np.float64)
    checked = on_jlane = 0
    for tick, s in enumerate(expected_states(fx)):
        if "xy" not in s:
            continue
        pos = np.array([float.fromhex(h) for h in s["pos"]])
        xy = oracle.get_pos(net, geom, s["loc_kind"], s["container"], s["side"], pos)
        assert np.array_equal(xy, recorded_xy(s)), f"{name} tick {tick}: get-pos differs"
        checked += 1
        on_jlane += int((net.kind[np.asarray(s["container"])] == K_JLANE).sum())
    assert checked >= fx["ticks"] // 10
    if name.startswith("grid"):
        assert on_jlane > 0, "no recorded frame caught an actor on a junction lane"


@pytest.mark.parametrize("n", [2, 3, 5])
def test_generator_geometry_matches_api_mirror(n):
    world = examples.grid_make_skeleton(n)
    _, idx = flatten.flatten_network(world)
    assert np.array_equal(flatten.flatten_geometry(world, idx), synth.grid_geometry(n))


def test_generator_geometry_is_consistent_with_the_network():
    net = synth.grid_network(6, geometry=True)
    g = net.geom
    sl = np.nonzero(net.kind == K_SEGLANE)[0]
    assert np.allclose(np.hypot(g[sl, 2], g[sl, 3]), net.lane_len[sl], rtol=1e-12)
    jl = np.nonzero(net.kind == K_JLANE)[0]
    lin, lout = net.lane_in[jl], net.lane_out[jl]
    assert np.allclose(g[jl, 0:2], g[lin, 0:2] + g[lin, 2:4], atol=1e-9)      # curve starts where the input lane ends
    assert np.allclose(g[jl, 6:8], g[lout, 0:2], atol=1e-9)                   # and ends where the output lane begins
    seg = np.nonzero(net.kind == K_SEGMENT)[0]
    mid = 0.5 * (g[seg, 0:2] + g[seg, 4:6])                                   # the two road sides straddle the axis
    lane = net.seg_outer_forw[seg]
    assert np.allclose(np.hypot(*(g[seg, 2:4]).T), net.lane_len[lane], rtol=1e-12)
    assert np.all(np.hypot(*(g[seg, 0:2] - mid).T) > 0)


def by_actor(xy, gid, n):
    out = np.full((n, 2), np.nan)
    live = gid >= 0
    out[gid[live]] = xy[live]
    assert not np.isnan(xy[live]).any() and np.isnan(xy[~live]).all()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
def test_device_positions_match_the_reference(name):
    from griddy_b200 import device
    fx, net, actors = load(name)
    net.geom = np.asarray(fx["network"]["geom"], np.float64).reshape(-1, 8)
    sim = device.Simulation(net, actors, dt=fx["dt"], two_size=fx["two_size"], reorder_period=16)
    stride = 50 if name == "triangle" else 1
    pending = checked = 0
    for tick, s in enumerate(expected_states(fx)):
        if tick:
            pending += 1
        if "xy" not in s or tick % stride:
            continue
        sim.step(pending)
        pending = 0
        xy, gid = sim.positions()
        order = sim.state().order
        assert np.array_equal(by_actor(xy, gid, actors.n)[order], recorded_xy(s)), f"{name} tick {tick}"
        xy32, gid32 = sim.positions(f32=True)
        assert np.array_equal(gid32, gid) and np.array_equal(xy32[gid >= 0], xy[gid >= 0].astype(np.float32))
        checked += 1
    assert checked >= 10


@pytest.mark.gpu
def test_device_positions_match_oracle_on_a_crowded_grid():
    from griddy_b200 import device
    net, actors = crowded_grid(6, 1500, agenda_len=4, seed=5, max_sleep=8, n_dest=6)
    net.geom = synth.grid_geometry(6)
    sim = device.Simulation(net, actors, reorder_period=32)
    jl = 0
    for _ in range(12):
        sim.step(25)
        st = sim.state()
        alive = st.alive.astype(bool)          # what a download holds for a vanished actor is unspecified
        want = oracle.get_pos(net, net.geom, st.loc_kind[alive], st.container[alive], st.side[alive], st.pos[alive])
        xy, gid = sim.positions()
        got = by_actor(xy, gid, actors.n)
        assert np.array_equal(got[alive], want)
        assert np.isnan(got[~alive]).all() and int((gid >= 0).sum()) == int(alive.sum()) == sim.count_alive()
        jl += int((net.kind[st.container[alive]] == K_JLANE).sum())
    assert jl > 0 and not st.alive.all(), "the workload must put actors on junction lanes and lose some to equal keys"


@pytest.mark.gpu
def test_frame_pipeline_overlaps_steps_and_returns_each_frame():
    from griddy_b200 import device
    net, actors = crowded_grid(8, 5000, agenda_len=4, seed=9, max_sleep=10)
    net.geom = synth.grid_geometry(8)
    sim = device.Simulation(net, actors, reorder_period=16)
    bufs = [device.PinnedFrame(actors.n), device.PinnedFrame(actors.n)]
    sim.step(40)
    want = []
    sim.frame_begin(bufs[0])
    want.append(sim.positions(f32=True)[0].copy())
    for k in range(1, 9):                       # step, begin the next frame, collect the previous one
        sim.step(3)
        sim.frame_begin(bufs[k & 1])
        want.append(sim.positions(f32=True)[0].copy())
        frame = sim.frame_end()
        assert np.array_equal(frame, want[k - 1], equal_nan=True), f"frame {k - 1}"
    assert np.array_equal(sim.frame_end(), want[8], equal_nan=True)
    # contract: two frames in flight at most, _end needs a _begin
    with pytest.raises(device.GriddyCudaError) as e:
        sim.frame_end()
    assert e.value.code == 4
    sim.frame_begin(bufs[0]); sim.frame_begin(bufs[1])
    with pytest.raises(device.GriddyCudaError) as e:
        sim.frame_begin(bufs[0])
    assert e.value.code == 4
    sim.frame_end(); sim.frame_end()
    for b in bufs:
        b.free()


@pytest.mark.gpu
def test_positions_need_geometry():
    from griddy_b200 import device
    net, actors = crowded_grid(3, 20)
    sim = device.Simulation(net, actors)
    with pytest.raises(device.GriddyCudaError) as e:
        sim.positions()
    assert e.value.code == 1


@pytest.mark.parametrize("name,make", [("simple", examples.simple_make_skeleton),
                                       ("triangle", examples.triangle_make_skeleton),
                                       ("grid-5x5-80", lambda: examples.grid_make_skeleton(5)),
                                       ("grid-3x3-crowded", lambda: examples.grid_make_skeleton(3)),
                                       ("grid-2x2-jam", lambda: examples.grid_make_skeleton(2))])
def test_api_mirror_geometry_matches_the_reference(name, make):
    """The Python mirror of spatial.scm (griddy_b200/core.py) — and through test_generator_geometry_matches_api_mirror
    the C++ generator — yields, bit for bit, the lane lengths and the drawing geometry that the reference's own
    spatial.scm produced when the fixture was recorded (fp64 vec2 stub, tests/golden/README.md)."""
    fx, net, _ = load(name)
    world = make()
    mirror, idx = flatten.flatten_network(world)
    assert np.array_equal(mirror.kind, net.kind)
    assert np.array_equal(mirror.lane_len, net.lane_len)
    assert np.array_equal(flatten.flatten_geometry(world, idx), np.asarray(fx["network"]["geom"]).reshape(-1, 8))
