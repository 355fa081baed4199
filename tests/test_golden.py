"""The oracle (and, on a GPU, the CUDA path) against golden vectors recorded by EXECUTING THE REFERENCE'S
OWN SOURCES (tests/golden/generate.py runs griddy/*.scm and examples/*.scm under tests/golden/minischeme.py;
see tests/golden/README.md for what that does and does not prove).

Compared after every tick, for every actor in `get-actors` order (world.scm:24-30): location kind, lane or
segment id, road side, pos-param (bit-exact), remaining agenda length, remaining sleep time, route status,
route head, max-speed.  The fixtures also fix the world's size, so a vanished actor is noticed."""
import glob
import gzip
import json
import os

import numpy as np
import pytest

from griddy_b200.flat import FlatActors, FlatNetwork
from oracle.oracle import Oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = sorted(os.path.basename(p)[: -len(".json.gz")] for p in glob.glob(os.path.join(GOLDEN, "*.json.gz")))


def load(name):
    with gzip.open(os.path.join(GOLDEN, name + ".json.gz"), "rt") as f:
        fx = json.load(f)
    n, a = fx["network"], fx["actors"]
    net = FlatNetwork(n["kind"], n["lane_len"], n["lane_dir"], n["lane_segment"], n["lane_out"], n["lane_junction"],
                      n["seg_outer_forw"], n["seg_outer_back"])
    actors = FlatActors(a["loc_kind"], a["container"], a["pos"], a["side"], a["speed"], a["speed_dt"], a["agenda_off"],
                        a["item_kind"], a["item_sleep"], a["item_seg"], a["item_side"], a["item_pos"],
                        route_off=a["route_off"], route_lane=a["route_lane"])
    return fx, net, actors


def expected_states(fx):
    for s in fx["states"]:
        for _ in range(s["repeat"]):
            yield s


def check(world, fx, actors, tick, s):
    """world.state() in get-actors order against one recorded state."""
    st = world.state()
    o = st.order
    where = f'{fx["scenario"]} tick {tick}'
    assert len(o) == len(s["pos"]), f"{where}: {len(o)} actors in the world, the reference has {len(s['pos'])}"
    assert bool(st.alive[o].all()), where
    left = (actors.agenda_off[1:] - actors.agenda_off[:-1])[o] - st.agenda_cur[o]
    got = {"loc_kind": st.loc_kind[o], "container": st.container[o], "side": st.side[o], "agenda_left": left,
           "sleep_left": st.sleep_left[o], "route_status": st.route_status[o], "route_head": st.route_head[o]}
    for k, v in got.items():
        assert np.array_equal(np.asarray(v, np.int64), np.asarray(s[k], np.int64)), f"{where}: {k} differs"
    assert np.array_equal(actors.speed[o], np.asarray(s["speed"])), f"{where}: actor identity (max-speed) differs"
    want = np.array([float.fromhex(h) for h in s["pos"]])
    assert np.array_equal(st.pos[o], want), f"{where}: pos-param not bit-identical"


def replay(make_world, name, stride=1):
    fx, net, actors = load(name)
    world = make_world(net, actors, fx)
    states = expected_states(fx)
    check(world, fx, actors, 0, next(states))
    pending = 0
    for tick, s in enumerate(states, start=1):
        pending += 1
        if tick % stride == 0 or tick == fx["ticks"]:
            world.step(pending)
            pending = 0
            check(world, fx, actors, tick, s)
    return world


def test_fixtures_exist():
    assert {"simple", "triangle", "grid-5x5-80", "grid-3x3-crowded", "grid-2x2-jam"} <= set(FIXTURES)


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_reproduces_the_reference_run(name):
    replay(lambda net, actors, fx: Oracle(net, actors, dt=fx["dt"], two_size=fx["two_size"]), name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
def test_device_reproduces_the_reference_run(name):
    from griddy_b200 import device
    stride = 20 if name == "triangle" else 1          # triangle idles for 9,970 of its 10,000 ticks
    replay(lambda net, actors, fx: device.Simulation(net, actors, dt=fx["dt"], two_size=fx["two_size"],
                                                      reorder_period=16), name, stride=stride)


def test_guile_dump_checker(tmp_path):
    """tests/golden/check_guile_dump.py (the comparison a maintainer with a real Guile runs on the output of
    guile/tools/headless-reference.scm) accepts a dump that agrees with the fixture and flags one that does not."""
    import subprocess
    import sys
    fx, _, _ = load("simple")
    lines = []
    for tick, s in enumerate(expected_states(fx)):
        lines.append(f"tick {tick} {len(s['pos'])}")
        for k in range(len(s["pos"])):
            route = {0: "none", 2: "done"}.get(s["route_status"][k], "arrive" if s["route_head"][k] == -1 else s["route_head"][k])
            lines.append(" ".join(str(v) for v in (
                s["loc_kind"][k], s["container"][k], s["side"][k], repr(float.fromhex(s["pos"][k])), s["agenda_left"][k],
                s["sleep_left"][k], route, s["speed"][k], s["xy"][2 * k], s["xy"][2 * k + 1])))
    good = tmp_path / "good.dump"
    good.write_text("\n".join(lines) + "\n")
    script = os.path.join(GOLDEN, "check_guile_dump.py")
    fixture = os.path.join(GOLDEN, "simple.json.gz")
    ok = subprocess.run([sys.executable, script, str(good), fixture, "--xy"], capture_output=True, text=True)
    assert ok.returncode == 0, ok.stdout
    bad = tmp_path / "bad.dump"
    lines[5] = lines[5].replace(lines[5].split()[3], "0.5", 1)
    bad.write_text("\n".join(lines) + "\n")
    ko = subprocess.run([sys.executable, script, str(bad), fixture], capture_output=True, text=True)
    assert ko.returncode == 1 and "pos-param" in ko.stdout


GRID_FIXTURES = {"grid-2x2-jam": 2, "grid-3x3-crowded": 3, "grid-5x5-80": 5}


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GRID_FIXTURES))
def test_device_finds_the_routes_the_reference_recorded(name):
    """find-route on the device (route.scm:17-51; griddy_set_route_graph): given only the lane graph — neighbours in
    the reference's order, segment lengths, lane midpoints, all from the API mirror, whose geometry
    tests/test_positions.py pins against the reference's — the device's A* returns, for every travel-to item, exactly
    the (turn-onto lane) ... list the reference's own find-route returned when the fixture was recorded, and the run
    that uses those routes reproduces the reference's, tick by tick."""
    import dataclasses

    import examples
    from griddy_b200 import device, flatten
    fx, fnet, factors = load(name)
    net, _ = flatten.flatten_network(examples.grid_make_skeleton(GRID_FIXTURES[name]))
    for f in ("kind", "lane_len", "lane_dir", "lane_segment", "lane_out", "lane_junction", "seg_outer_forw", "seg_outer_back"):
        assert np.array_equal(getattr(net, f), getattr(fnet, f)), f"the API mirror builds another network than the fixture's: {f}"
    bare = dataclasses.replace(factors, route_off=None, route_lane=None)
    sim = device.Simulation(net, bare, dt=fx["dt"], two_size=fx["two_size"], reorder_period=16)
    off, lanes = sim.routes()
    assert np.array_equal(off, factors.route_off), "route lengths differ from the reference's find-route"
    assert np.array_equal(lanes, factors.route_lane), "routes differ from the reference's find-route"
    assert len(lanes) > 0
    states = expected_states(fx)
    check(sim, fx, factors, 0, next(states))
    for tick, s in enumerate(states, start=1):
        sim.step(1)
        check(sim, fx, factors, tick, s)
