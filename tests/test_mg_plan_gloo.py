"""Host side of the multi-GPU path, two processes over gloo (no GPU): each rank derives, from the
replicated network and the ownership map alone, what it exports to and imports from its neighbour;
the two views must agree without any list ever being exchanged, every cross-rank lane transition must
be covered, and the ranks' shares of the actors must partition the world."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def plan(L, net, owner, owner_rank, reader_rank):
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lanes, junctions = np.empty(net.n_items, np.int32), np.empty(net.n_items, np.int32)
    nl, nj = C.c_int64(0), C.c_int64(0)
    rc = L.griddy_mg_plan(C.c_int64(net.n_items), p(net.kind), p(net.lane_in), p(net.lane_out), p(net.lane_junction),
                          p(owner), C.c_int32(owner_rank), C.c_int32(reader_rank), p(lanes), C.byref(nl), p(junctions),
                          C.byref(nj))
    assert rc == 0
    return lanes[: nl.value].tolist(), junctions[: nj.value].tolist()


def worker(rank, world, port, n_grid, n_actors):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as g
    g.build()
    from griddy_b200 import device, synth
    L = device.lib()
    net = synth.grid_network(n_grid)
    owner = synth.grid_owner(net, world)
    views = {}
    for other in range(world):
        if other != rank:
            views[other] = {"export": plan(L, net, owner, rank, other), "import": plan(L, net, owner, other, rank)}
    ids = synth.actor_ids_of_rank(net, n_actors, owner, rank)
    everyone = [None] * world
    dist.all_gather_object(everyone, {"views": views, "ids": ids})
    # my export to r is r's import from me, item for item
    for other, v in views.items():
        assert v["export"] == everyone[other]["views"][rank]["import"], (rank, other)
        assert v["import"] == everyone[other]["views"][rank]["export"], (rank, other)
    # completeness: every lane -> lane transition that leaves my part is in my import list
    jl = np.flatnonzero(net.kind == 3)
    enters = np.concatenate([np.stack([net.lane_in[jl], jl], 1), np.stack([jl, net.lane_out[jl]], 1)])
    cross = enters[(owner[enters[:, 0]] == rank) & (owner[enters[:, 1]] != rank)]
    for other, v in views.items():
        want = sorted(set(cross[owner[cross[:, 1]] == other][:, 1].tolist()))
        assert v["import"][0] == want, (rank, other)
        gates = sorted(set(net.lane_junction[[x for x in want if net.kind[x] == 3]].tolist()))
        assert v["import"][1] == gates
    if world == 2:
        assert len(cross) > 0
    # the ranks' actors partition the world, each in ascending id order
    allids = np.concatenate([e["ids"] for e in everyone])
    assert len(allids) == n_actors and np.array_equal(np.sort(allids), np.arange(n_actors))
    assert np.all(np.diff(ids) > 0)
    a = synth.grid_actors(net, n_actors, ids=ids, agenda_len=4)
    assert np.all(owner[a.container] == rank)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_grid", [(2, 6), (3, 7)])
def test_partition_plans_agree_across_ranks(world, n_grid):
    port = 29600 + world
    mp.spawn(worker, args=(world, port, n_grid, 5000), nprocs=world, join=True)
