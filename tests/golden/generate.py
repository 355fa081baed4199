#!/usr/bin/env python
"""Generate the golden fixtures in this directory by EXECUTING THE REFERENCE'S OWN SOURCES.

Runs in the build container only (needs /root/reference).  For every scenario it loads the reference's
example script under tests/golden/minischeme.py, lets the reference's `simulate` (simulate.scm:133-179)
build its `load` and `update` closures (the stubbed `run-game` captures them instead of opening an SDL
window), calls `load` once and `update` once per tick, and records after every tick the world the
reference built: for each actor in `get-actors` order (world.scm:24-30) its location, agenda and route.

The fixture also holds the INPUTS the oracle / device need to replay the run: the network flattened by
registration index (with the lane lengths the reference's own `get-length` returned), and the actors
with their agendas and the routes the reference's own `find-route` returned.

For the drawing read-back (`get-pos` of every actor, spatial.scm:125-149) it records, per item, the
geometry the reference's own `get-pos` / `get-vec` / `curve` return (8 doubles, layout of
include/griddy_cuda.h griddy_upload_geometry), and every XY_EVERY ticks `(get-pos actor)` of every actor.

    python tests/golden/generate.py            # all scenarios
    python tests/golden/generate.py simple     # one

Scenarios named `grid-*` are the reference's examples/procedural-grid.scm with three literals replaced
in the source text before evaluation (grid size, actor count, maximum sleep), listed in SCENARIOS.
"""
import gzip
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import minischeme as ms  # noqa: E402
from minischeme import S, plist  # noqa: E402

REF = "/root/reference"

SCENARIOS = {
    # name: (example file, textual substitutions, ticks, rng seed)
    "simple": ("examples/simple.scm", [], 40, 1),
    "triangle": ("examples/triangle.scm", [], 10000, 1),
    "grid-5x5-80": ("examples/procedural-grid.scm", [], 600, 7),            # as shipped (n = 5, 80 actors)
    "grid-3x3-crowded": ("examples/procedural-grid.scm",
                         [("(n 5)", "(n 3)"), ("(list-tabulate 80 make-actor)", "(list-tabulate 150 make-actor)"),
                          ("(random-integer 300)", "(random-integer 12)")], 500, 11),
    "grid-2x2-jam": ("examples/procedural-grid.scm",
                     [("(n 5)", "(n 2)"), ("(list-tabulate 80 make-actor)", "(list-tabulate 60 make-actor)"),
                      ("(random-integer 300)", "(random-integer 4)")], 400, 13),
}


XY_EVERY = 10   # scenarios with more than one actor record (get-pos actor) every 10th tick


class Run:
    def __init__(self, name):
        path, subs, self.ticks, seed = SCENARIOS[name]
        self.name = name
        self.I = ms.Interp([REF], seed=seed)
        text = open(os.path.join(REF, path)).read()
        for old, new in subs:
            assert text.count(old) == 1, (old, text.count(old))
            text = text.replace(old, new)
        tmp = os.path.join("/tmp", f"golden_{name}.scm")
        with open(tmp, "w") as f:
            f.write(text)
        self.I.load_file(tmp)                       # ends in (simulate make-skeleton add-actors! ...)
        self.game = self.I.captured_game
        self.sim = self.I.modules[("griddy", "simulate")]
        self.core = self.I.modules[("griddy", "core")]

    def call(self, fname, *args, module=None):
        m = module or self.sim
        return self.I.apply(self.I.lookup(S(fname), m.env), list(args))

    def world(self):
        return self.sim.env.vars[S("world")]

    def cls(self, name):
        return self.I.lookup(S(name), self.sim.env)

    # ---- flattening (mirrors guile/griddy/cuda.scm and griddy_b200/flatten.py) ------------------
    def index(self, world):
        items = plist(world.slots[S("static-items")])
        n = len(items)
        return {id(it): n - 1 - i for i, it in enumerate(items)}, items

    def flatten_network(self, world):
        idx, items = self.index(world)
        n = len(items)
        net = {"kind": [4] * n, "lane_len": [0.0] * n, "lane_dir": [0] * n, "lane_segment": [-1] * n,
               "lane_out": [-1] * n, "lane_junction": [-1] * n, "seg_outer_forw": [-1] * n,
               "seg_outer_back": [-1] * n}
        LocOff = self.cls("<location/off-road>")
        for it in items:
            r = idx[id(it)]
            cname = it.cls.name
            if cname == "<road-junction>":
                net["kind"][r] = 0
            elif cname == "<road-segment>":
                net["kind"][r] = 1
                for side, key in (("forw", "seg_outer_forw"), ("back", "seg_outer_back")):
                    loc = self.call("make", LocOff, ms.Kw("road-segment"), it, ms.Kw("road-side-direction"), S(side),
                                    ms.Kw("pos-param"), 0.5)
                    try:
                        on = self.call("off-road->on-road", loc)
                        net[key][r] = idx[id(on.slots[S("road-lane")])]
                    except ms.SchemeThrow:
                        pass
            elif cname == "<road-lane/segment>":
                net["kind"][r] = 2
                net["lane_len"][r] = float(self.call("get-length", it))
                net["lane_dir"][r] = 0 if it.slots[S("direction")] is S("forw") else 1
                net["lane_segment"][r] = idx[id(it.slots[S("segment")])]
            elif cname == "<road-lane/junction>":
                net["kind"][r] = 3
                net["lane_len"][r] = float(self.call("get-length", it))
                net["lane_out"][r] = idx[id(self.call("get-segment-lane", it, module=self.core))]
                net["lane_junction"][r] = idx[id(it.slots[S("junction")])]
        return net, idx

    def flatten_geometry(self, world, idx):
        """Per item, 8 doubles (include/griddy_cuda.h griddy_upload_geometry), all from the reference's code:
        segment lane   beg.x beg.y vec.x vec.y 0 0 0 0        (get-pos lane 'beg), (get-vec lane)
        junction lane  p0.x p0.y p1.x p1.y p2.x p2.y p3.x p3.y   (ref lane 'curve)
        segment        of.x of.y vec.x vec.y ob.x ob.y 0 0    (get-pos <off-road forw|back at pos-param 0>), (get-vec segment)"""
        _, items = self.index(world)
        geom = [[0.0] * 8 for _ in items]
        LocOff = self.cls("<location/off-road>")
        xy = lambda v: [float(v.x), float(v.y)]
        for it in items:
            r = idx[id(it)]
            cname = it.cls.name
            if cname == "<road-lane/segment>":
                geom[r] = xy(self.call("get-pos", it, S("beg"))) + xy(self.call("get-vec", it)) + [0.0] * 4
            elif cname == "<road-lane/junction>":
                c = it.slots[S("curve")]
                geom[r] = [v for k in range(4) for v in xy(c.p[k])]
            elif cname == "<road-segment>":
                o = []
                for side in ("forw", "back"):
                    loc = self.call("make", LocOff, ms.Kw("road-segment"), it, ms.Kw("road-side-direction"), S(side),
                                    ms.Kw("pos-param"), 0)
                    o.append(xy(self.call("get-pos", loc)))
                geom[r] = o[0] + xy(self.call("get-vec", it)) + o[1] + [0.0] * 2
        return [v for g in geom for v in g]

    def flatten_actors(self, world, idx):
        order = plist(self.call("get-actors", world))
        actors = order[::-1]            # id k-th of get-actors -> M-1-k: linking in id order rebuilds the lists
        A = {k: [] for k in ("loc_kind", "container", "pos", "side", "speed", "speed_dt", "agenda_off", "item_kind",
                             "item_sleep", "item_seg", "item_side", "item_pos", "route_off", "route_lane")}
        A["agenda_off"].append(0)
        A["route_off"].append(0)
        ts = self.I.lookup(S("*simulate/time-step*"), self.sim.env)
        for a in actors:
            loc = a.slots[S("location")]
            on = loc.cls.name == "<location/on-road>"
            A["loc_kind"].append(1 if on else 0)
            A["container"].append(idx[id(loc.slots[S("road-lane" if on else "road-segment")])])
            A["pos"].append(float(loc.slots[S("pos-param")]))
            A["side"].append(0 if on or loc.slots[S("road-side-direction")] is S("forw") else 1)
            speed = a.slots[S("max-speed")]
            A["speed"].append(float(speed))
            A["speed_dt"].append(float(ms.num_mul(speed, ts)))      # (exact->inexact (* max-speed 1/15))
            cur = None if on else loc
            assert a.slots[S("route")] is S("none")
            for item in plist(a.slots[S("agenda")]):
                kind = item.car.name
                if kind == "sleep-for":
                    A["item_kind"].append(0); A["item_sleep"].append(int(item.cdr.car))
                    A["item_seg"].append(-1); A["item_side"].append(0); A["item_pos"].append(0.0)
                else:
                    dest = item.cdr.car
                    A["item_kind"].append(1); A["item_sleep"].append(0)
                    A["item_seg"].append(idx[id(dest.slots[S("road-segment")])])
                    A["item_side"].append(0 if dest.slots[S("road-side-direction")] is S("forw") else 1)
                    A["item_pos"].append(float(dest.slots[S("pos-param")]))
                    dest_on = self.call("off-road->on-road", dest)
                    if cur is not None:
                        steps = plist(self.call("find-route", self.call("off-road->on-road", cur), dest_on))
                        A["route_lane"].extend(idx[id(s.cdr.car)] for s in steps if s.car is S("turn-onto"))
                    cur = self.call("on-road->off-road", dest_on)
                A["route_off"].append(len(A["route_lane"]))
            A["agenda_off"].append(len(A["item_kind"]))
        return A

    # ---- state after a tick -----------------------------------------------------------------------
    def state(self, world, with_xy=False):
        idx, _ = self.index(world)
        st = {k: [] for k in ("loc_kind", "container", "side", "pos", "agenda_left", "sleep_left", "route_status",
                              "route_head", "speed")}
        for a in plist(self.call("get-actors", world)):
            loc = a.slots[S("location")]
            on = loc.cls.name == "<location/on-road>"
            st["loc_kind"].append(1 if on else 0)
            st["container"].append(idx[id(loc.slots[S("road-lane" if on else "road-segment")])])
            st["side"].append(0 if on or loc.slots[S("road-side-direction")] is S("forw") else 1)
            st["pos"].append(float(loc.slots[S("pos-param")]).hex())
            agenda = plist(a.slots[S("agenda")])
            st["agenda_left"].append(len(agenda))
            st["sleep_left"].append(int(agenda[0].cdr.car) if agenda and agenda[0].car is S("sleep-for") else 0)
            route = a.slots[S("route")]
            if route is S("none"):
                st["route_status"].append(0); st["route_head"].append(-2)
            elif route is ms.NIL:
                st["route_status"].append(2); st["route_head"].append(-2)
            else:
                st["route_status"].append(1)
                head = route.car
                st["route_head"].append(idx[id(head.cdr.car)] if head.car is S("turn-onto") else -1)
            st["speed"].append(float(a.slots[S("max-speed")]))
            if with_xy:
                v = self.call("get-pos", a)
                st.setdefault("xy", []).extend([float(v.x), float(v.y)])
        return st

    def run(self):
        t0 = time.time()
        self.I.apply(self.game["load"], [])
        world = self.world()
        net, idx = self.flatten_network(world)
        actors = self.flatten_actors(world, idx)
        net["geom"] = self.flatten_geometry(world, idx)
        every = 1 if len(actors["loc_kind"]) == 1 else XY_EVERY
        states = [self.state(world, True)]           # tick 0: the world `load` built
        for t in range(self.ticks):
            self.I.apply(self.game["update"], [0])
            states.append(self.state(self.world(), (t + 1) % every == 0))
            if (t + 1) % 100 == 0:
                print(f"  {self.name}: tick {t + 1}/{self.ticks}, {len(states[-1]['pos'])} actors, {time.time() - t0:.0f}s",
                      flush=True)
        # run-length encode unchanged states (triangle idles for 9,970 of its 10,000 ticks)
        packed, prev = [], None
        for s in states:
            if s == prev:
                packed[-1]["repeat"] += 1
            else:
                packed.append({"repeat": 1, **s})
                prev = s
        fixture = {"scenario": self.name, "source": SCENARIOS[self.name][0], "substitutions": SCENARIOS[self.name][1],
                   "seed": SCENARIOS[self.name][3], "ticks": self.ticks, "dt": float(self.I.lookup(S("*simulate/time-step*"), self.sim.env)),
                   "two_size": float(2 * self.I.lookup(S("*actor/size*"), self.sim.env)),
                   "network": net, "actors": actors, "states": packed}
        out = os.path.join(HERE, f"{self.name}.json.gz")
        with gzip.GzipFile(out, "wb", mtime=0) as f:
            f.write(json.dumps(fixture, separators=(",", ":")).encode())
        print(f"{self.name}: {self.ticks} ticks, {len(actors['loc_kind'])} actors, {len(packed)} distinct states -> "
              f"{out} ({os.path.getsize(out)} bytes, {time.time() - t0:.0f}s)")


if __name__ == "__main__":
    for name in (sys.argv[1:] or list(SCENARIOS)):
        Run(name).run()
