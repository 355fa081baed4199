#!/usr/bin/env python
"""Compare a dump written by guile/tools/headless-reference.scm (a run of the REAL reference under Guile)
with a fixture of this directory, tick by tick, bit for bit.

    python tests/golden/check_guile_dump.py triangle.dump tests/golden/triangle.json.gz

This is the check that turns "parity with the reference's source code under an independent evaluator"
(README.md) into parity with Guile itself; it needs a host with Guile 3.0.7 + Chickadee 0.8.0 + pfds to
produce the dump, which this repository's build image is not.  Lane lengths and geometry are not compared
(they are inputs); positions (x, y) are compared only with --xy, since Chickadee's vec2 precision decides them.
"""
import gzip
import json
import sys


def parse_dump(path):
    ticks, cur = [], None
    with open(path) as f:
        for line in f:
            w = line.split()
            if not w:
                continue
            if w[0] == "tick":
                cur = []
                ticks.append(cur)
                continue
            route = {"none": (0, -2), "done": (2, -2), "arrive": (1, -1)}.get(w[6])
            if route is None:
                route = (1, int(w[6]))
            cur.append({"loc_kind": int(w[0]), "container": int(w[1]), "side": int(w[2]), "pos": float(w[3]),
                        "agenda_left": int(w[4]), "sleep_left": int(w[5]), "route_status": route[0],
                        "route_head": route[1], "speed": float(w[7]), "xy": (float(w[8]), float(w[9]))})
    return ticks


def fixture_states(path):
    with gzip.open(path, "rt") as f:
        fx = json.load(f)
    for s in fx["states"]:
        for _ in range(s["repeat"]):
            yield s


def main(argv):
    with_xy = "--xy" in argv
    argv = [a for a in argv if a != "--xy"]
    if len(argv) != 3:
        print(__doc__)
        return 2
    dump = parse_dump(argv[1])
    bad = 0
    for tick, (got, want) in enumerate(zip(dump, fixture_states(argv[2]))):
        if len(got) != len(want["pos"]):
            print(f"tick {tick}: {len(got)} actors in the dump, {len(want['pos'])} in the fixture")
            bad += 1
            continue
        for k, a in enumerate(got):
            for key in ("loc_kind", "container", "side", "agenda_left", "sleep_left", "route_status", "route_head", "speed"):
                if a[key] != want[key][k]:
                    print(f"tick {tick} actor {k}: {key} {a[key]} != {want[key][k]}")
                    bad += 1
            if a["pos"] != float.fromhex(want["pos"][k]):
                print(f"tick {tick} actor {k}: pos-param {a['pos'].hex()} != {want['pos'][k]}")
                bad += 1
            if with_xy and "xy" in want and tuple(want["xy"][2 * k: 2 * k + 2]) != a["xy"]:
                print(f"tick {tick} actor {k}: xy {a['xy']} != {want['xy'][2 * k: 2 * k + 2]}")
                bad += 1
        if bad > 50:
            break
    print(f"{min(len(dump), tick + 1)} ticks compared, {bad} differences")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
