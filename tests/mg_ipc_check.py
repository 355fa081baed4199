"""One process per rank (torchrun): a partitioned world stepped over the window transport — peer stores and
flags between the ranks' kernels, windows shared through cudaIpc — must equal the oracle's single world at every
checkpoint.  Rank 0 merges the ranks' read-backs and compares; every rank's device digest must add up too.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
        tests/mg_ipc_check.py [--grid 12] [--actors 12000] [--ticks 600] [--every 25] [--strips 2] [--tiled]

With fewer GPUs than ranks the ranks share GPUs (several processes per device, time-sliced): the same kernels, the
same flags, the same IPC mapping — slower, but it runs on a one-GPU box.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=12)
    ap.add_argument("--actors", type=int, default=12000)
    ap.add_argument("--ticks", type=int, default=600)
    ap.add_argument("--every", type=int, default=25)
    ap.add_argument("--strips", type=int, default=1)
    ap.add_argument("--tiled", action="store_true")
    ap.add_argument("--async-steps", action="store_true", help="tick by tick with griddy_step_async, one blocking call per checkpoint")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = local % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo")   # control plane only: the 128-byte blobs and the read-backs
    from griddy_b200 import device, digest, synth
    from helpers import crowded_grid

    net, actors = crowded_grid(args.grid, args.actors, n_dest=12, seed=77, agenda_len=4)
    owner = synth.grid_owner(net, world, args.strips)
    mine = np.flatnonzero(owner[actors.container] == rank)
    part, part_owner, part_actors = net, owner, actors
    if args.tiled:   # this rank uploads only its rows and the ring next to them
        rb, re = synth.tile_rows(args.grid, world, args.strips, rank)
        gb, ge = synth.tile_ranges(args.grid, rb, re)
        part = net.tile_of(gb, ge, owner)
        part_owner = part.owner
        part_actors = device.with_destinations(net, actors)
    sim = device.Simulation(part, device.subset_actors(part_actors, mine), device=dev, reorder_period=16,
                            partition=(rank, world, part_owner, 4096, args.actors), gids=mine, max_agenda_items=4)
    device.connect_ipc(sim)
    ref = None
    if rank == 0:
        from oracle.oracle import Oracle
        ref = Oracle(net, actors)
    crossed = 0
    for t in range(0, args.ticks, args.every):
        if args.async_steps:   # a frame loop's way of stepping: nothing waits for the device between ticks
            for _ in range(args.every):
                sim.step_async(1)
            sim.step(0)
        else:
            sim.step(args.every)
        parts, digs = [None] * world, [None] * world
        dist.all_gather_object(parts, sim.local_state())
        dist.all_gather_object(digs, sim.digest())
        if rank == 0:
            ref.step(args.every)
            ds, os_ = device.merge_local_states(parts, actors.n), ref.state()
            bad = ds.diff(os_)
            assert not bad, f"tick {t + args.every}: the partitioned device differs from the oracle in {bad}"
            m = ds.alive.astype(bool)
            assert np.array_equal(ds.pos[m], os_.pos[m])
            assert digest.combine(digs) == digest.commutative(os_), f"tick {t + args.every}: device digests do not add up"
            crossed += int((owner[ds.container[m]] != owner[actors.container[m]]).sum())
    if rank == 0:
        assert crossed > 0 and ref.vanished > 0
        print(f"mg_ipc_check ok: {world} ranks on {torch.cuda.device_count()} GPU(s), {'tiled, ' if args.tiled else ''}"
              f"{args.strips} strip(s) per rank, {args.actors} actors, {args.ticks} ticks, {ref.vanished} vanished, "
              f"boundary crossings observed, last step {sim.last_step_ms:.2f} ms", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
