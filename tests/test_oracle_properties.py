"""Invariants that follow from the reference semantics, checked on the oracle."""
import numpy as np

from helpers import crowded_grid, grid_flat
from oracle.oracle import Oracle


def test_actor_count_is_conserved_up_to_replacements():
    net, actors = crowded_grid(4, 600, n_dest=3)
    o = Oracle(net, actors)
    for _ in range(30):
        o.step(20)
        s = o.state()
        assert int(s.alive.sum()) == actors.n - o.vanished == len(s.order)
        lc, jc = o.occupancy()
        assert int(lc.sum()) == len(s.order)
    assert o.vanished > 0, "the crowded workload is meant to exercise equal-key replacement"


def test_no_overtaking_among_residents():
    """Followers never pass their leader while both stay on a lane (route.scm:68-74)."""
    net, actors = crowded_grid(4, 400, n_dest=5, seed=11)
    o = Oracle(net, actors)
    prev = o.state()
    for _ in range(300):
        o.step(1)
        cur = o.state()
        both = (prev.alive & cur.alive & prev.loc_kind & cur.loc_kind).astype(bool)
        both &= (prev.container == cur.container) & ~((cur.route_status == 2) & (cur.pos < prev.pos))
        ids = np.nonzero(both)[0]
        lane, p0, p1 = prev.container[ids], prev.pos[ids], cur.pos[ids]
        k = np.lexsort((p0, lane))
        same = lane[k][1:] == lane[k][:-1]
        assert np.all(p1[k][1:][same] > p1[k][:-1][same])
        prev = cur


def test_sleep_for_n_lasts_n_plus_one_ticks():
    net, actors = grid_flat(3, 50)
    o = Oracle(net, actors)
    first = actors.item_sleep[actors.agenda_off[:-1]]
    for t in range(1, 40):
        o.step(1)
        s = o.state()
        assert np.array_equal(s.agenda_cur >= 1, first < t)   # simulate.scm:74-78


def test_lazy_grid_routes_equal_the_unrolled_route_lists():
    """With the grid rule the oracle derives the route head on demand instead of unrolling the whole
    (turn-onto lane) ... list at begin-route$ (87 GB at 1024 x 1024 / 16M actors): both modes give the same world,
    tick by tick, state, order and occupancy."""
    for n, n_actors, seed in ((5, 700, 3), (9, 3000, 12)):
        net, actors = crowded_grid(n, n_actors, n_dest=max(4, n), seed=seed, agenda_len=4)
        lazy, eager = Oracle(net, actors, lazy_routes=True), Oracle(net, actors, lazy_routes=False)
        for t in range(250):
            lazy.step(1); eager.step(1)
            a, b = lazy.state(), eager.state()
            assert not a.diff(b) and np.array_equal(a.pos, b.pos), f"{n}x{n} tick {t + 1}"
            assert all(np.array_equal(x, y) for x, y in zip(lazy.occupancy(), eager.occupancy()))
        assert lazy.vanished == eager.vanished > 0
