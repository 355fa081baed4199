"""The reference's three examples restated against the Python mirror of its API."""
from griddy_b200.core import (Actor, LocationOffRoad, RoadJunction, RoadLaneSegment, RoadSegment, World,
                              add, agenda_push, connect_by_rank, get_static_items, link)


# examples/simple.scm:8-37
def simple_make_skeleton():
    world = World()
    junction_1 = RoadJunction(x=100, y=250)
    junction_2 = RoadJunction(x=400, y=250)
    segment = RoadSegment()
    lane = RoadLaneSegment(direction="forw")
    link(lane, segment)
    link(junction_1, segment, junction_2)
    for item in (junction_1, junction_2, segment, lane):
        add(world, item)
    return world


def simple_add_actors(world):
    segment = get_static_items(world, RoadSegment)[0]
    actor = Actor()
    link(actor, LocationOffRoad(segment, "forw", 0.25))
    agenda_push(actor, ("travel-to", LocationOffRoad(segment, "forw", 0.75)))


# examples/triangle.scm:8-57
def triangle_make_skeleton():
    world = World()
    j1, j2, j3 = RoadJunction(x=100, y=100), RoadJunction(x=400, y=100), RoadJunction(x=250, y=400)
    s1, s2, s3 = RoadSegment(), RoadSegment(), RoadSegment()
    l1, l2, l3 = (RoadLaneSegment(direction="forw") for _ in range(3))
    link(l1, s1); link(l2, s2); link(l3, s3)
    link(j1, s1, j2); link(j2, s2, j3); link(j3, s3, j1)
    for item in (j1, j2, j3, s1, s2, s3, l1, l2, l3):
        add(world, item)
    for j in (j1, j2, j3):
        connect_by_rank(j, world)
    return world


triangle_add_actors = simple_add_actors   # triangle.scm:43-57 is the same procedure


# examples/procedural-grid.scm:13-67 with n as a parameter (n = 5 there)
def grid_make_skeleton(n=5):
    world = World()
    t = lambda x: 50 + x * 150
    junctions = [[None] * n for _ in range(n)]
    for i in range(n):
        for j in range(n):
            junctions[i][j] = RoadJunction(x=t(i), y=t(j))
            add(world, junctions[i][j])

    def add_segments(dim, i, j):
        if (dim == "i" and j == n - 1) or (dim == "j" and i == n - 1):
            return
        junction_1 = junctions[i][j]
        junction_2 = junctions[i][j + 1] if dim == "i" else junctions[i + 1][j]
        segment = RoadSegment()
        lane_1, lane_2 = RoadLaneSegment("forw"), RoadLaneSegment("back")
        lane_3, lane_4 = RoadLaneSegment("forw"), RoadLaneSegment("back")
        link(junction_1, segment, junction_2)
        link(lane_1, segment)
        link(lane_2, segment)
        add(world, segment); add(world, lane_1); add(world, lane_2)
        dim_val = i if dim == "i" else j
        lane = lane_3 if dim_val % 4 == 0 else lane_4
        if dim_val % 2 == 0:
            link(lane, segment)
            add(world, lane)

    for dim in ("i", "j"):
        for i in range(n):
            for j in range(n):
                add_segments(dim, i, j)
    for i in range(n):
        for j in range(n):
            connect_by_rank(junctions[i][j], world)
    return world
