"""The Guile glue (guile/griddy/cuda.scm, guile/griddy/simulate-cuda.scm) EXECUTED — without Guile and without a
GPU: under the Scheme evaluator that also records the golden fixtures, with (system foreign) replaced by a
Python stand-in for libgriddy_cuda.so that records every call (tests/golden/minischeme_ffi.py).

The reference's own example scripts are run with the ONE change INTEGRATION.md tells a user to make —
`(griddy simulate)` becomes `(griddy simulate-cuda)` in their use-modules — and the test checks that
  - `load` flattens the world the reference's code built into exactly the arrays the fixtures hold (which the
    device tests upload): network, drawing geometry, actors, agendas, routes from the reference's find-route;
  - `update` is one griddy_step(ctx, 1);
  - `draw` fetches the device-computed positions and paints one circle per living actor;
  - `cuda-download-world!` rebuilds, from a read-back, a world whose get-actors order, locations and speeds are
    the reference's at that tick;
  - a nonzero return code becomes (throw 'griddy-cuda code message).
Needs /root/reference (build container); skipped elsewhere."""
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "griddy")), reason="the reference is not mounted here")

from test_golden import expected_states, load  # noqa: E402


class FakeLibrary:
    """Plays libgriddy_cuda.so for the glue: records calls, serves read-backs from a fixture state."""

    def __init__(self, ffi):
        self.ffi = ffi
        self.calls = []          # (name, decoded arguments)
        self.fail = {}           # name -> return code to hand back
        self.state = None        # fixture state served by griddy_download_positions
        self.by_id = None        # per-actor-id state served by griddy_download_actors
        self.geom = None
        self.net = None
        self.n_actors = 0

    # decoding of what crosses the boundary
    def arr(self, p, dtype, count):
        if p.obj is None:
            return None
        return np.frombuffer(bytes(p.obj.data), dtype=dtype)[:count].copy()

    def call(self, name, a):
        if name in self.fail:
            self.calls.append((name, None))
            return self.fail[name]
        h = getattr(self, name, None)
        if h is None:
            raise AssertionError(f"the glue called {name}, which this stand-in does not know")
        return h(*a)

    def griddy_create(self, slot, device):
        struct.pack_into("<Q", slot.obj.data, 0, 0xC0FFEE)
        self.calls.append(("griddy_create", {"device": device}))
        return 0

    def griddy_destroy(self, ctx):
        self.calls.append(("griddy_destroy", {"ctx": ctx.address}))

    def griddy_last_error(self, ctx):
        return self.ffi.Pointer("stand-in says no")

    def griddy_set_constants(self, ctx, dt, two_size):
        self.calls.append(("griddy_set_constants", {"ctx": ctx.address, "dt": dt, "two_size": two_size}))
        return 0

    def griddy_upload_network(self, ctx, n, kind, length, direction, seg, out, jun, of, ob):
        n = int(n)
        self.calls.append(("griddy_upload_network", {
            "n": n, "kind": self.arr(kind, np.uint8, n), "lane_len": self.arr(length, np.float64, n),
            "lane_dir": self.arr(direction, np.uint8, n), "lane_segment": self.arr(seg, np.int32, n),
            "lane_out": self.arr(out, np.int32, n), "lane_junction": self.arr(jun, np.int32, n),
            "seg_outer_forw": self.arr(of, np.int32, n), "seg_outer_back": self.arr(ob, np.int32, n)}))
        self.n_items = n
        return 0

    def griddy_upload_geometry(self, ctx, geom):
        self.calls.append(("griddy_upload_geometry", {"geom": self.arr(geom, np.float64, 8 * self.n_items)}))
        return 0

    def griddy_upload_actors(self, ctx, n, loc_kind, container, pos, side, speed, speed_dt, agenda_off, item_kind,
                             item_sleep, item_seg, item_side, item_pos, route_off, route_lane):
        n = int(n)
        off = self.arr(agenda_off, np.int64, n + 1)
        ni = int(off[n])
        no_routes = route_off.obj is None     # routes 'device: find-route runs on the device
        roff = np.zeros(ni + 1, np.int64) if no_routes else self.arr(route_off, np.int64, ni + 1)
        self.calls.append(("griddy_upload_actors", {
            "no_routes": no_routes,
            "n": n, "loc_kind": self.arr(loc_kind, np.uint8, n), "container": self.arr(container, np.int32, n),
            "pos": self.arr(pos, np.float64, n), "side": self.arr(side, np.uint8, n),
            "speed": self.arr(speed, np.float64, n), "speed_dt": self.arr(speed_dt, np.float64, n),
            "agenda_off": off, "item_kind": self.arr(item_kind, np.uint8, ni),
            "item_sleep": self.arr(item_sleep, np.int32, ni), "item_seg": self.arr(item_seg, np.int32, ni),
            "item_side": self.arr(item_side, np.uint8, ni), "item_pos": self.arr(item_pos, np.float64, ni),
            "route_off": roff, "route_lane": np.zeros(0, np.int32) if no_routes else self.arr(route_lane, np.int32, int(roff[ni]))}))
        self.n_actors = n
        return 0

    def griddy_set_route_graph(self, ctx, nbr_off, nbr, cost_out, midpoint_xy):
        n = self.n_items
        off = self.arr(nbr_off, np.int64, n + 1)
        self.calls.append(("griddy_set_route_graph", {
            "nbr_off": off, "nbr": self.arr(nbr, np.int32, int(off[n])), "cost_out": self.arr(cost_out, np.float64, n),
            "midpoint": self.arr(midpoint_xy, np.float64, 2 * n).reshape(-1, 2)}))
        return 0

    def griddy_step_async(self, ctx, n_ticks):
        self.calls.append(("griddy_step_async", {"ctx": ctx.address, "n_ticks": int(n_ticks)}))
        return 0

    def griddy_step(self, ctx, n_ticks):
        self.calls.append(("griddy_step", {"ctx": ctx.address, "n_ticks": int(n_ticks)}))
        return 0

    def griddy_last_step_ms(self, ctx):
        return 0.25

    def griddy_slots_in_use(self, ctx):
        return self.n_actors

    # read-backs, served from self.state (a fixture state: actors in get-actors order; actor id = M-1-k)
    def ids(self):
        m = len(self.state["pos"])
        return m, [m - 1 - k for k in range(m)]

    def griddy_download_actors(self, ctx, alive, loc_kind, container, side, pos, agenda_cur, sleep_left, route_status,
                               route_head, order, n_order):
        st = self.by_id            # an ActorState indexed by actor id, as the device returns it (here: the oracle's)
        for a in range(self.n_actors):
            alive.obj.data[a] = int(st.alive[a])
            loc_kind.obj.data[a] = int(st.loc_kind[a])
            struct.pack_into("<i", container.obj.data, 4 * a, int(st.container[a]))
            side.obj.data[a] = int(st.side[a])
            struct.pack_into("<d", pos.obj.data, 8 * a, float(st.pos[a]))
        for k, a in enumerate(st.order):
            struct.pack_into("<i", order.obj.data, 4 * k, int(a))
        struct.pack_into("<q", n_order.obj.data, 0, len(st.order))
        self.calls.append(("griddy_download_actors", {"nulls": [p.obj is None for p in (agenda_cur, sleep_left, route_status, route_head)]}))
        return 0

    def griddy_download_positions(self, ctx, capacity, n, xy64, xy32, gid):
        m, ids = self.ids()
        assert int(capacity) >= m and xy64.obj is None and gid.obj is None
        xy = np.asarray(self.state["xy"], np.float64).reshape(-1, 2)
        for k, a in enumerate(ids):          # slot = actor id here
            struct.pack_into("<ff", xy32.obj.data, 8 * a, xy[k, 0], xy[k, 1])
        struct.pack_into("<q", n.obj.data, 0, m)
        self.calls.append(("griddy_download_positions", {"n": m}))
        return 0


def run_example(name, example, subs=(), seed=1):
    import minischeme as ms
    import minischeme_ffi as ffi
    lib = FakeLibrary(ffi)
    I = ms.Interp([REF, os.path.join(ROOT, "guile")], seed=seed)
    ffi.install(I, lib.call)
    text = open(os.path.join(REF, example)).read()
    assert text.count("(griddy simulate)") == 1
    text = text.replace("(griddy simulate)", "(griddy simulate-cuda)")     # the documented one-line switch
    for old, new in subs:
        assert text.count(old) == 1
        text = text.replace(old, new)
    path = os.path.join("/tmp", f"glue_{name}.scm")
    with open(path, "w") as f:
        f.write(text)
    I.load_file(path)                         # ends in (simulate make-skeleton add-actors!) -> run-game captured
    return I, lib, ms


def the_call(lib, name):
    found = [c for c in lib.calls if c[0] == name]
    assert len(found) == 1, (name, [c[0] for c in lib.calls])
    return found[0][1]


SCENARIOS = {
    "simple": ("examples/simple.scm", (), 1),
    "triangle": ("examples/triangle.scm", (), 1),
    "grid-2x2-jam": ("examples/procedural-grid.scm",
                     (("(n 5)", "(n 2)"), ("(list-tabulate 80 make-actor)", "(list-tabulate 60 make-actor)"),
                      ("(random-integer 300)", "(random-integer 4)")), 13),
}


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_glue_flattens_the_world_like_the_fixtures(name):
    example, subs, seed = SCENARIOS[name]
    I, lib, ms = run_example(name, example, subs, seed)
    fx, net, actors = load(name)
    game = I.captured_game
    assert {"load", "update", "draw", "quit"} <= set(game) and game["update-hz"] == 15
    I.apply(game["load"], [])

    c = the_call(lib, "griddy_set_constants")
    assert c["ctx"] == 0xC0FFEE and c["dt"] == fx["dt"] and c["two_size"] == fx["two_size"]
    got = the_call(lib, "griddy_upload_network")
    assert got["n"] == net.n_items
    for k in ("kind", "lane_len", "lane_dir", "lane_segment", "lane_out", "lane_junction", "seg_outer_forw", "seg_outer_back"):
        assert np.array_equal(got[k], getattr(net, k)), k
    geom = the_call(lib, "griddy_upload_geometry")["geom"]
    assert np.array_equal(geom, np.asarray(fx["network"]["geom"], np.float64))
    got = the_call(lib, "griddy_upload_actors")
    assert got["n"] == actors.n
    for k in ("loc_kind", "container", "pos", "side", "speed", "speed_dt", "agenda_off", "item_kind", "item_sleep",
              "item_seg", "item_side", "item_pos", "route_off", "route_lane"):
        assert np.array_equal(got[k], getattr(actors, k)), k
    order = [c[0] for c in lib.calls]
    assert order.index("griddy_upload_network") < order.index("griddy_upload_geometry") < order.index("griddy_upload_actors")

    # update = one tick on the device
    before = len(lib.calls)
    I.apply(game["update"], [1.0 / 15.0])
    assert lib.calls[before:] == [("griddy_step", {"ctx": 0xC0FFEE, "n_ticks": 1})]

    # draw = the skeleton canvas, then one circle per living actor at the positions the device computed
    states = list(expected_states(fx))
    lib.state = next(s for s in states[1:] if "xy" in s)
    I.drawn.clear()
    I.apply(game["draw"], [0.0])
    assert the_call(lib, "griddy_download_positions")["n"] == actors.n
    canvas = I.drawn[-1]
    assert canvas[0] == "canvas" and canvas[1][0] == "superimpose"
    circles = sorted((c[1], c[2]) for c in canvas[1][1])
    want = np.asarray(lib.state["xy"], np.float64).reshape(-1, 2).astype(np.float32)
    assert circles == sorted((float(x), float(y)) for x, y in want)
    assert all(c[0] == "circle" and c[3] == 5 for c in canvas[1][1])        # *actor/size*, constants.scm:34


def test_glue_uploads_the_lane_graph_for_find_route_on_the_device():
    """(cuda-upload-world! ctx world #:routes 'device): no route lists cross the boundary; instead the glue uploads the
    graph find-route searches — neighbours in the reference's order, costs, midpoints, all computed by the reference's
    own get-junction-lanes / get-segment-lane / get-length / get-midpoint — which must equal what the API mirror
    (griddy_b200/flatten.py) builds and tests/test_golden.py feeds to the device's A*."""
    import examples
    from griddy_b200 import flatten
    example, subs, seed = SCENARIOS["grid-2x2-jam"]
    subs = subs + (("(simulate make-skeleton add-actors! #:width 700 #:height 700 #:duration 800)",
                    "(use-modules (griddy cuda))\n(define world (make-skeleton))\n(add-actors! world)\n"
                    "(define ctx (cuda-open))\n(cuda-upload-world! ctx world #:routes 'device)\n(cuda-step-async! ctx 3)"),)
    I, lib, ms = run_example("grid-2x2-device-routes", example, subs, seed)
    fx, net, actors = load("grid-2x2-jam")
    got = the_call(lib, "griddy_set_route_graph")
    want, _ = flatten.flatten_network(examples.grid_make_skeleton(2))
    for k in ("nbr_off", "nbr", "cost_out", "midpoint"):
        assert np.array_equal(got[k], getattr(want, k)), k
    up = the_call(lib, "griddy_upload_actors")
    assert up["no_routes"] and np.array_equal(up["item_seg"], actors.item_seg) and np.array_equal(up["pos"], actors.pos)
    order = [c[0] for c in lib.calls]
    assert order.index("griddy_upload_network") < order.index("griddy_set_route_graph") < order.index("griddy_upload_actors")
    assert the_call(lib, "griddy_step_async")["n_ticks"] == 3


@pytest.mark.parametrize("name", ["simple", "grid-2x2-jam"])
def test_glue_rebuilds_the_world_from_a_read_back(name):
    example, subs, seed = SCENARIOS[name]
    I, lib, ms = run_example(name, example, subs, seed)
    fx, net, actors = load(name)
    S = ms.S
    cuda = None
    I.apply(I.captured_game["load"], [])
    cuda = I.modules[("griddy", "cuda")]
    sim = I.modules[("griddy", "simulate-cuda")]
    call = lambda fname, *args, module=cuda: I.apply(I.lookup(S(fname), module.env), list(args))
    # the handle `load` stored, and the user's make-skeleton
    handle = I.lookup(S("handle"), I.captured_game["load"].env) if hasattr(I.captured_game["load"], "env") else None
    make_skeleton = I.lookup(S("make-skeleton"), I.user.env)
    ctx = I.lookup(S("ctx"), I.captured_game["load"].env)
    states = list(expected_states(fx))
    tick = 25
    lib.state = states[tick]
    # what the device would hand back at that tick, by actor id: the oracle's state (itself checked against the
    # fixtures tick by tick in test_golden.py)
    from oracle.oracle import Oracle
    ref = Oracle(net, actors, dt=fx["dt"], two_size=fx["two_size"])
    ref.step(tick)
    lib.by_id = ref.state()
    world = call("cuda-download-world!", ctx, handle, make_skeleton)
    assert the_call(lib, "griddy_download_actors")["nulls"] == [True] * 4
    core = I.modules[("griddy", "core")]
    got = ms.plist(I.apply(I.lookup(S("get-actors"), sim.env), [world]))
    items = ms.plist(world.slots[S("static-items")])
    index = {id(it): len(items) - 1 - i for i, it in enumerate(items)}
    s = lib.state
    assert len(got) == len(s["pos"])
    for k, a in enumerate(got):
        loc = a.slots[S("location")]
        on = loc.cls.name == "<location/on-road>"
        assert int(on) == s["loc_kind"][k]
        assert index[id(loc.slots[S("road-lane" if on else "road-segment")])] == s["container"][k]
        assert float(loc.slots[S("pos-param")]) == float.fromhex(s["pos"][k])
        if not on:
            assert (0 if loc.slots[S("road-side-direction")] is S("forw") else 1) == s["side"][k]
        assert float(a.slots[S("max-speed")]) == s["speed"][k]


def test_glue_turns_error_codes_into_throws():
    I, lib, ms = run_example("simple", "examples/simple.scm")
    lib.fail["griddy_upload_network"] = 1
    with pytest.raises(ms.SchemeThrow) as e:
        I.apply(I.captured_game["load"], [])
    assert e.value.key is ms.S("griddy-cuda") and e.value.args_ == [1, "stand-in says no"]


def test_glue_headless_driver_and_frame_pipeline():
    """simulate/headless (the fixed-tick-count driver BASELINE.json's configs ask for) and the asynchronous frame
    read-back, as the glue calls them."""
    I, lib, ms = run_example("simple", "examples/simple.scm")
    fx, net, actors = load("simple")
    from oracle.oracle import Oracle
    ref = Oracle(net, actors, dt=fx["dt"], two_size=fx["two_size"])
    ref.step(25)
    lib.by_id = ref.state()
    ev = lambda src: [I.eval(f, I.user.env) for f in ms.parse_all(src)][-1]
    world = ev("(simulate/headless make-skeleton add-actors! #:ticks 25)")
    names = [c[0] for c in lib.calls]
    assert names == ["griddy_create", "griddy_set_constants", "griddy_upload_network", "griddy_upload_geometry",
                     "griddy_upload_actors", "griddy_step", "griddy_download_actors", "griddy_destroy"]
    assert the_call(lib, "griddy_step")["n_ticks"] == 25
    (actor,) = ms.plist(I.apply(I.lookup(ms.S("get-actors"), I.modules[("griddy", "simulate-cuda")].env), [world]))
    want = list(expected_states(fx))[25]
    assert float(actor.slots[ms.S("location")].slots[ms.S("pos-param")]) == float.fromhex(want["pos"][0])

    # frame pipeline: pinned buffer from the library, begin / end
    import minischeme_ffi as ffi
    begun = []
    lib.griddy_host_alloc = lambda n: ffi.Pointer(ffi.ByteVector(int(n)))
    lib.griddy_positions_begin = lambda ctx, buf, cap: (begun.append((buf.obj, int(cap))), 0)[1]

    def positions_end(ctx, n):
        struct.pack_into("<q", n.obj.data, 0, 7)
        return 0
    lib.griddy_positions_end = positions_end
    cuda = I.modules[("griddy", "cuda")]
    call = lambda fname, *args: I.apply(I.lookup(ms.S(fname), cuda.env), list(args))
    ctx = call("cuda-open")
    buf = call("make-frame-buffer", 100)
    assert len(buf.data) == 800
    call("cuda-frame-begin!", ctx, buf)
    assert begun == [(buf, 100)]
    assert call("cuda-frame-end!", ctx) == 7


@pytest.mark.parametrize("name,example,ticks", [("simple", "examples/simple.scm", 40), ("triangle", "examples/triangle.scm", 60)])
def test_headless_reference_runner_and_dump_checker(name, example, ticks, tmp_path):
    """guile/tools/headless-reference.scm — the script a maintainer with a real Guile runs to get the true baseline
    and golden vectors — executed under the evaluator; its dump passes tests/golden/check_guile_dump.py against
    the fixture, positions included."""
    import subprocess
    import minischeme as ms
    import minischeme_ffi as ffi
    I = ms.Interp([REF, os.path.join(ROOT, "guile")], seed=1)
    ffi.install(I, lambda fn, args: (_ for _ in ()).throw(AssertionError("no foreign calls expected")))
    I.argv = ["headless-reference.scm", os.path.join(REF, example), str(ticks)]
    I.load_file(os.path.join(ROOT, "guile", "tools", "headless-reference.scm"))
    dump = tmp_path / f"{name}.dump"
    dump.write_text("".join(I.stdout))
    assert "".join(I.stdout).count("tick ") == ticks + 1
    assert "actor-updates" in I.stderr[-1]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "check_guile_dump.py"), str(dump),
                        os.path.join(ROOT, "tests", "golden", f"{name}.json.gz"), "--xy"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert f"{ticks + 1} ticks compared, 0 differences" in r.stdout


def test_glue_bindings_match_the_header():
    """Every (c-fn "griddy_…" ret args…) in guile/griddy/cuda.scm has the return type, the number and the types of
    arguments that include/griddy_cuda.h declares (pointers '*, int64_t int64, int/int32_t int, double double)."""
    import re
    hdr = open(os.path.join(ROOT, "include", "griddy_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    protos = {}
    for ret, name, params in re.findall(r"\b(int|void|double|int64_t|const char\*|void\*)\s+(griddy_[a-z_]+)\s*\(([^;{]*?)\)\s*;", hdr):
        def kind(t):
            t = t.strip()
            if "*" in t:
                return "'*"
            base = t.split()[0] if t.split()[0] != "const" else t.split()[1]
            return {"int64_t": "int64", "int32_t": "int", "int": "int", "double": "double", "void": "void"}[base]
        args = [] if params.strip() in ("", "void") else [kind(p) for p in params.split(",")]
        protos[name] = (kind(ret), args)
    scm = open(os.path.join(ROOT, "guile", "griddy", "cuda.scm")).read()
    bound = re.findall(r'\(c-fn "(griddy_[a-z_]+)"\s+([^\s()]+)((?:\s+[^\s()]+)*)\)', scm)
    assert len(bound) >= 15
    for name, ret, args in bound:
        assert name in protos, f"{name} is not declared in the header"
        want_ret, want_args = protos[name]
        assert ret == want_ret, (name, ret, want_ret)
        assert args.split() == want_args, (name, args.split(), want_args)
