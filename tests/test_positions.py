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


# ---- delta frames (griddy_delta_begin / _end / _apply) -------------------------------------------------------------
def synthetic_delta_frame(capacity, n_slots, changed_slots, values, key, first_of_block=None):
    """A frame laid out by hand as include/griddy_cuda.h documents it (T = 1024), in plain host memory."""
    T = 1024
    nb_cap, blocks = (capacity + T - 1) // T, (n_slots + T - 1) // T
    mask_off, off_off = 64, 64 + nb_cap * 128
    xy_off = (off_off + nb_cap * 4 + 63) // 64 * 64
    raw = np.zeros(xy_off + nb_cap * T * 8, np.uint8)
    hdr = raw[:64].view(np.int64)
    hdr[:] = [0x3156444459524747, n_slots, len(changed_slots), 7, int(key), blocks, capacity, 0]
    bits = np.zeros(nb_cap * T, np.uint8)
    bits[changed_slots] = 1
    raw[mask_off:mask_off + nb_cap * 128] = np.packbits(bits, bitorder="little")
    per_block = bits.reshape(nb_cap, T).sum(axis=1).astype(np.int64)
    order = list(range(blocks)) if first_of_block is None else first_of_block   # the order the blocks reserved room in
    first = np.zeros(nb_cap, np.uint32)
    at = 0
    for b in order:
        first[b] = at
        at += int(per_block[b])
    raw[off_off:off_off + nb_cap * 4] = first.view(np.uint8)
    vals = raw[xy_off:].view(np.float32).reshape(-1, 2)
    rank_in_block = np.arange(len(changed_slots)) - np.concatenate([[0], np.cumsum(per_block)])[np.asarray(changed_slots) // T]
    vals[first.astype(np.int64)[np.asarray(changed_slots) // T] + rank_in_block] = values
    return raw


def test_delta_apply_follows_the_documented_layout():
    import ctypes as C
    from griddy_b200 import device
    L = device.lib()
    rng = np.random.default_rng(3)
    cap, n_slots = 5000, 4100
    assert L.griddy_delta_frame_bytes(C.c_int64(cap)) == 64 + 5 * 128 + 64 + 5 * 1024 * 8   # header, mask, table (padded), values
    slots = np.sort(rng.choice(n_slots, 900, replace=False))
    vals = rng.standard_normal((900, 2)).astype(np.float32)
    vals[::7] = np.nan                                   # slots that lost their actor
    xy = rng.standard_normal((cap, 2)).astype(np.float32)
    want = xy.copy()
    want[slots] = vals
    frame = synthetic_delta_frame(cap, n_slots, slots, vals, key=False, first_of_block=[3, 0, 4, 2, 1])
    assert L.griddy_delta_apply(frame.ctypes.data_as(C.c_void_p), xy.ctypes.data_as(C.c_void_p), C.c_int64(cap)) == 0
    assert np.array_equal(xy, want, equal_nan=True)
    got_slots, got_vals = device.decode_delta(frame)     # the numpy restatement the GPU tests check apply against
    assert np.array_equal(got_slots, slots) and np.array_equal(got_vals, vals, equal_nan=True)
    # a key frame starts from all-NaN
    frame = synthetic_delta_frame(cap, n_slots, slots, vals, key=True)
    want = np.full((cap, 2), np.nan, np.float32)
    want[slots] = vals
    assert L.griddy_delta_apply(frame.ctypes.data_as(C.c_void_p), xy.ctypes.data_as(C.c_void_p), C.c_int64(cap)) == 0
    assert np.array_equal(xy, want, equal_nan=True)
    # malformed: wrong magic, an array smaller than the frame, a count that disagrees with the mask
    bad = frame.copy(); bad[0] ^= 1
    assert L.griddy_delta_apply(bad.ctypes.data_as(C.c_void_p), xy.ctypes.data_as(C.c_void_p), C.c_int64(cap)) == 1
    assert L.griddy_delta_apply(frame.ctypes.data_as(C.c_void_p), xy.ctypes.data_as(C.c_void_p), C.c_int64(n_slots - 1)) == 1
    bad = frame.copy(); bad[:64].view(np.int64)[2] += 1
    assert L.griddy_delta_apply(bad.ctypes.data_as(C.c_void_p), xy.ctypes.data_as(C.c_void_p), C.c_int64(cap)) == 4


def check_delta_stream(sim, bufs, full, strides, label=""):
    """Steps `sim` by the given strides with two delta frames in flight; after every frame the patched array must
    equal griddy_download_positions of the same tick, bit for bit.  Returns per-frame (key, n_changed, living)."""
    want, log = [], []

    def begin(k):
        sim.delta_begin(bufs[k & 1])
        want.append(sim.positions(f32=True)[0].copy())

    def end(k):
        fr = sim.delta_end()
        slots, vals = fr.decode()                    # the layout, restated in numpy
        assert len(slots) == fr.n_changed and fr.n_slots == len(want[k])
        before = full.copy()
        fr.apply(full)
        ref = np.full_like(full, np.nan) if fr.key else before.copy()
        ref[slots] = vals
        assert np.array_equal(full, ref, equal_nan=True), f"{label}frame {k}: apply and decode disagree"
        assert np.array_equal(full[: fr.n_slots], want[k], equal_nan=True), f"{label}frame {k} (key={fr.key})"
        assert np.isnan(full[fr.n_slots:]).all()
        if not fr.key:   # a delta frame holds what changed: a pos-param that moved by an ulp may round to the same floats
            same = np.all((before[slots] == vals) | (np.isnan(before[slots]) & np.isnan(vals)), axis=1)
            assert same.sum() <= 0.02 * len(slots) + 2, f"{label}frame {k}: {int(same.sum())} of {len(slots)} slots sent are unchanged"
        log.append((fr.key, fr.n_changed, int((~np.isnan(want[k][:, 0])).sum())))

    begin(0)
    for k, stride in enumerate(strides, start=1):
        sim.step_async(stride)
        begin(k)
        end(k - 1)
    end(len(strides))
    sim.step(0)
    return log


@pytest.mark.gpu
def test_delta_frames_rebuild_every_full_frame():
    from griddy_b200 import device
    net, actors = crowded_grid(8, 6000, agenda_len=4, seed=11, max_sleep=10)
    net.geom = synth.grid_geometry(8)
    sim = device.Simulation(net, actors, reorder_period=16)
    bufs = [device.DeltaFrame(actors.n), device.DeltaFrame(actors.n)]
    full = np.full((actors.n, 2), np.nan, np.float32)
    sim.step(30)
    reorders0 = sim.reorders
    log = check_delta_stream(sim, bufs, full, [1, 1, 1, 2, 1, 3, 1, 1, 5, 1, 1, 1, 7, 1, 1, 1, 1, 2, 1, 1] * 4)
    keys = [k for k, _, _ in log]
    assert keys[0] and 2 <= sum(keys) - 1 <= sim.reorders - reorders0, "reorders must show up as key frames"
    deltas = [(c, n) for k, c, n in log if not k]
    assert sum(c for c, _ in deltas) < 0.8 * sum(n for _, n in deltas), "delta frames should be sparser than full frames"
    assert sim.count_alive() < actors.n, "the workload must lose actors to equal keys (slots that turn NaN)"
    # contract: two frames in flight at most, _end needs a _begin
    with pytest.raises(device.GriddyCudaError) as e:
        sim.delta_end()
    assert e.value.code == 4
    sim.delta_begin(bufs[0]); sim.delta_begin(bufs[1])
    with pytest.raises(device.GriddyCudaError) as e:
        sim.delta_begin(bufs[0])
    assert e.value.code == 4
    sim.delta_end(); sim.delta_end()
    for b in bufs:
        b.free()


@pytest.mark.gpu
def test_delta_frames_of_a_standing_world_are_empty():
    """Idle actors and sleepers cost nothing on PCIe: after the key frame, frames of a world in which nobody moves
    carry no values at all."""
    from griddy_b200 import device
    net, actors = crowded_grid(4, 300, agenda_len=0)
    net.geom = synth.grid_geometry(4)
    sim = device.Simulation(net, actors)
    bufs = [device.DeltaFrame(actors.n), device.DeltaFrame(actors.n)]
    full = np.full((actors.n, 2), np.nan, np.float32)
    log = check_delta_stream(sim, bufs, full, [1, 1, 4, 1])
    assert log[0][0] and log[0][1] == log[0][2] == sim.count_alive()
    assert all(not k and c == 0 for k, c, _ in log[1:])
    for b in bufs:
        b.free()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_delta_frames_on_a_partitioned_world(world):
    """Every rank streams its own slots: actors that leave for another rank turn NaN, arrivals appear in fresh slots."""
    from griddy_b200 import device
    net, actors = crowded_grid(9, 4000, agenda_len=4, seed=21, max_sleep=8)
    net.geom = synth.grid_geometry(9)
    owner = synth.grid_owner(net, world, 2)
    ps = device.PartitionedSimulation(net, actors, owner, world, reorder_period=24)
    cap = 2 * actors.n
    bufs = [[device.DeltaFrame(cap), device.DeltaFrame(cap)] for _ in range(world)]
    full = [np.full((cap, 2), np.nan, np.float32) for _ in range(world)]
    ps.step(20)
    arrivals = 0
    covered = [0] * world
    for k in range(40):
        for r, sim in enumerate(ps.ranks):
            sim.delta_begin(bufs[r][k & 1])
        want = [sim.positions(f32=True)[0].copy() for sim in ps.ranks]
        ps.step(1 if k % 5 else 3)
        for r, sim in enumerate(ps.ranks):
            fr = sim.delta_end()
            fr.apply(full[r])
            assert np.array_equal(full[r][: fr.n_slots], want[r], equal_nan=True), f"rank {r} frame {k} (key={fr.key})"
            assert np.isnan(full[r][fr.n_slots:]).all()
            arrivals += int(not fr.key and fr.n_slots > covered[r])
            covered[r] = fr.n_slots
    assert arrivals > 0, "actors must have crossed between the ranks (fresh slots between two reorders)"
    for bb in bufs:
        for b in bb:
            b.free()
