"""Known-answer test pinning the oracle on examples/simple.scm.

The reference ships no golden vectors, so this answer is derived by hand from the
reference's arithmetic, evaluated here with Python's exact rationals standing in for Guile's numeric
tower (constants.scm:34-38, route.scm:53-90), not by running the oracle:
  lane length 282 (junction radius 9, spatial.scm:53-66), delta = fl(80/15) * fl(1/282),
  tick 1 begin-route$ at 0.25, ticks 2..27 advance by delta, tick 28 arrive at 0.75,
  tick 29 end-route$ to off-road (segment, forw, 0.75), then idle.
"""
from fractions import Fraction

import numpy as np

from helpers import simple_flat, triangle_flat
from oracle.oracle import Oracle


def expected_simple():
    length = 282.0
    delta = float(Fraction(80) * Fraction(1, 15)) * (1.0 / length)
    pos, traj = 0.25, []
    traj.append(("on", 0.25, 1, -1))            # tick 1: begin-route$
    while True:
        nxt = pos + delta
        if nxt >= 0.75:
            traj.append(("on", 0.75, 2, -2))    # arrive: route '()
            break
        pos = nxt
        traj.append(("on", pos, 1, -1))
    traj.append(("off", 0.75, 0, -2))           # end-route$
    return delta, traj


def test_simple_constants():
    delta, traj = expected_simple()
    assert delta.hex() == "0x1.35dce5f9f2af8p-6"
    assert (10.0 / 282.0).hex() == "0x1.227f179a53849p-5"
    assert traj[1][1] == 0.26891252955082745 and traj[2][1] == 0.2878250591016549
    assert traj[26][1] == 0.7417257683215137
    assert traj[27][:2] == ("on", 0.75) and traj[28][0] == "off" and len(traj) == 29


def test_simple_trajectory():
    net, actors = simple_flat()
    assert net.lane_len[3] == 282.0 and actors.n == 1
    assert actors.route_off.tolist() == [0, 0]             # same lane: route is ((arrive-at 0.75))
    _, traj = expected_simple()
    o = Oracle(net, actors)
    for t, (where, pos, rs, head) in enumerate(traj, start=1):
        o.step(1)
        s = o.state()
        assert s.loc_kind[0] == (1 if where == "on" else 0), t
        assert s.container[0] == (3 if where == "on" else 2), t
        assert s.pos[0] == pos, (t, s.pos[0], pos)
        assert s.route_status[0] == rs and s.route_head[0] == head, t
        assert s.agenda_cur[0] == (0 if t < 29 else 1), t
    o.step(11)
    s = o.state()
    assert (s.loc_kind[0], s.container[0], s.side[0], s.pos[0], s.agenda_cur[0]) == (0, 2, 0, 0.75, 1)
    assert s.order.tolist() == [0] and o.vanished == 0


def test_triangle_runs_to_idle():
    # triangle.scm:44 takes (car (get-static-items world <road-segment>)) = the LAST registered
    # segment (id 5, lane id 8); the junction lanes are never used.
    net, actors = triangle_flat()
    assert actors.container[0] == 5 and net.seg_outer_forw[5] == 8
    o = Oracle(net, actors)
    o.step(1)
    s = o.state()
    assert (s.loc_kind[0], s.container[0], s.pos[0]) == (1, 8, 0.25)
    delta = float(Fraction(80, 15)) * (1.0 / net.lane_len[8])
    o.step(1)
    assert o.state().pos[0] == 0.25 + delta
    o.step(9998)
    s = o.state()
    assert (s.loc_kind[0], s.container[0], s.pos[0], s.agenda_cur[0], s.route_status[0]) == (0, 5, 0.75, 1, 0)
