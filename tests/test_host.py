"""Host-side logic: the API mirror, the flattener and the generator agree; the C ABI is complete."""
import ctypes
import os
import re

import numpy as np
import pytest

import examples
from griddy_b200 import core, device, flatten, synth
from helpers import grid_flat, world_to_flat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n", [2, 3, 5, 7])
def test_generator_matches_api_mirror(n):
    """csrc/synth.cpp emits, bit for bit, what procedural-grid.scm builds through the API."""
    fn, _ = flatten.flatten_network(examples.grid_make_skeleton(n))
    sn = synth.grid_network(n)
    for f in ("kind", "lane_len", "lane_dir", "lane_segment", "lane_out", "lane_junction",
              "seg_outer_forw", "seg_outer_back", "lane_in"):
        assert np.array_equal(getattr(fn, f), getattr(sn, f)), f


def test_grid_sizes_match_survey():
    s = synth.grid_sizes(256)
    assert (s["n_junctions"], s["n_segments"], s["n_seg_lanes"]) == (65536, 130560, 326400)
    assert s["n_items"] == s["n_junctions"] + s["n_segments"] + s["n_seg_lanes"] + s["n_jlanes"]
    assert synth.grid_sizes(5)["n_items"] == 463


def test_registration_order_is_reverse_of_static_items():
    w = examples.simple_make_skeleton()
    idx = flatten.registration_index(w)
    assert [idx[i] for i in w.static_items] == [3, 2, 1, 0]        # world.scm:21-22 conses
    assert isinstance(w.static_items[0], core.RoadLaneSegment)


def test_outer_lane_table():
    """location.scm:16-24 including the two cross-over cases."""
    seg = core.RoadSegment()
    f0, f1 = core.RoadLaneSegment("forw"), core.RoadLaneSegment("forw")
    core.link(f0, seg); core.link(f1, seg)
    assert core.get_outer_lane(seg, "forw") is f1
    assert core.get_outer_lane(seg, "back") is f0        # ('back 0 _) -> first forw lane
    seg2 = core.RoadSegment()
    b0, b1 = core.RoadLaneSegment("back"), core.RoadLaneSegment("back")
    core.link(b0, seg2); core.link(b1, seg2)
    assert core.get_outer_lane(seg2, "forw") is b1       # ('forw _ 0) -> last back lane
    with pytest.raises(core.GriddyError):
        core.get_outer_lane(core.RoadSegment(), "forw")


def test_add_validates_segments():
    w = core.World()
    with pytest.raises(core.GriddyError) as e:
        core.add(w, core.RoadSegment())
    assert e.value.key == "unlinked-road-segment"        # road.scm:113-115


def test_actor_ids_rebuild_segment_lists():
    """Linking actors in id order must reproduce every consed list (flatten.py docstring)."""
    w = examples.grid_make_skeleton(3)
    segs = core.get_static_items(w, core.RoadSegment)
    made = []
    for k in range(7):
        a = core.Actor(max_speed=30 + k)
        core.link(a, core.LocationOffRoad(segs[k % 3], "forw", 0.3 + 0.01 * k))
        made.append(a)
    idx = flatten.registration_index(w)
    fa, objs = flatten.flatten_actors(w, idx)
    order = core.get_actors(w)
    assert objs == order[::-1]
    # rebuild: cons ids in ascending order
    lists = {}
    for a in range(fa.n):
        lists.setdefault(int(fa.container[a]), []).insert(0, a)
    for seg in segs:
        assert [objs[a] for a in lists.get(idx[seg], [])] == seg.actors


def test_speed_dt_follows_the_numeric_tower():
    assert flatten.speed_dt_of(80) == 80 / 15                      # exact rational, rounded once
    assert flatten.speed_dt_of(31.0) == 31.0 * (1 / 15)            # flonum * fl(1/15)
    assert flatten.speed_dt_of(31.0) != 31.0 / 15 == flatten.speed_dt_of(31)


def test_routes_for_triangle_and_grid():
    net, actors = world_to_flat(examples.triangle_make_skeleton, examples.triangle_add_actors)
    assert actors.route_off.tolist() == [0, 0]

    def add_actors(world):
        segs = core.get_static_items(world, core.RoadSegment)
        a = core.Actor(max_speed=40)
        core.link(a, core.LocationOffRoad(segs[-1], "forw", 0.5))
        core.agenda_push(a, ("travel-to", core.LocationOffRoad(segs[0], "back", 0.5)))

    net, actors = world_to_flat(lambda: examples.grid_make_skeleton(3), add_actors)
    lanes = actors.route_lane.tolist()
    assert len(lanes) >= 2 and len(lanes) % 2 == 0
    kinds = net.kind[lanes].tolist()
    assert kinds[0::2] == [3] * (len(lanes) // 2) and kinds[1::2] == [2] * (len(lanes) // 2)
    # consecutive: junction lane -> its output segment lane
    assert all(net.lane_out[lanes[k]] == lanes[k + 1] for k in range(0, len(lanes), 2))
    assert lanes[-1] == net.seg_outer_back[actors.item_seg[0]]


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "griddy_cuda.h")).read()
    declared = set(re.findall(r"\b(griddy_[a-z_]+)\s*\(", hdr))
    assert declared == set(device.SYMBOLS), declared ^ set(device.SYMBOLS)
    L = ctypes.CDLL(device.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert L.griddy_abi_version() == 6


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    net, actors = grid_flat(2, 4)
    with pytest.raises(device.GriddyCudaError):
        device.Simulation(net, actors)


def test_product_code_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "griddy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "libgriddy_oracle" not in src, f


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/griddy_cuda.h compiles as C99 with no C++ or CUDA in sight."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    src = tmp_path / "h.c"
    src.write_text('#include "griddy_cuda.h"\nint main(void) { return griddy_abi_version() == GRIDDY_ABI_VERSION ? 0 : 1; }\n')
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                        "-I", os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_subset_of_a_routed_world_names_its_routes_in_the_whole_table():
    """A rank's share of a partitioned world with explicit routes (griddy_mg_set_routes): the whole route table travels
    to every rank, and each of the rank's agenda items names its own route in it."""
    import random
    rng = random.Random(3)

    def add_actors(world):
        segs = core.get_static_items(world, core.RoadSegment)
        loc = lambda: core.LocationOffRoad(rng.choice(segs), rng.choice(["forw", "back"]), 0.3 + 0.4 * rng.random())
        for _ in range(25):
            a = core.Actor(max_speed=35)
            core.link(a, loc())
            core.agenda_push(a, ("sleep-for", 2))
            core.agenda_push(a, ("travel-to", loc()))
            core.agenda_push(a, ("travel-to", loc()))
    net, actors = world_to_flat(lambda: examples.grid_make_skeleton(3), add_actors)
    mine = np.arange(1, actors.n, 3)
    part = device.subset_actors(actors, mine)
    assert part.route_off is None and part.route_table_off is actors.route_off and part.route_table_lane is actors.route_lane
    assert len(part.item_route) == part.n_agenda_items == int((actors.agenda_off[mine + 1] - actors.agenda_off[mine]).sum())
    k = 0
    for a in mine:
        for item in range(actors.agenda_off[a], actors.agenda_off[a + 1]):
            assert part.item_route[k] == item and part.item_kind[k] == actors.item_kind[item]
            lo, hi = actors.route_off[item], actors.route_off[item + 1]
            assert (hi > lo) == (actors.item_kind[item] == 1 and hi > lo)        # only travel-to items have steps
            k += 1
