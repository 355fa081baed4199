"""Run under torchrun with one rank per GPU: a partitioned world stepped over the NCCL transport must
equal the oracle's single world at every checkpoint.  Rank 0 merges the ranks' read-backs and compares.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/mg_nccl_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from griddy_b200 import device, synth
    from helpers import crowded_grid

    n, n_actors, ticks, every = 12, 12000, 600, 25
    net, actors = crowded_grid(n, n_actors, n_dest=12, seed=77, agenda_len=4)
    owner = synth.grid_owner(net, world)
    mine = np.flatnonzero(owner[actors.container] == rank)
    sim = device.Simulation(net, device.subset_actors(actors, mine), device=local, reorder_period=16,
                            partition=(rank, world, owner, 4096, n_actors), gids=mine)
    device.connect_nccl(sim)
    ref = None
    if rank == 0:
        from oracle.oracle import Oracle
        ref = Oracle(net, actors)
    crossed = 0
    for t in range(0, ticks, every):
        sim.step(every)
        parts = [None] * world
        dist.all_gather_object(parts, sim.local_state())
        if rank == 0:
            ref.step(every)
            ds, os_ = device.merge_local_states(parts, actors.n), ref.state()
            bad = ds.diff(os_)
            assert not bad, f"tick {t + every}: NCCL-partitioned device differs from the oracle in {bad}"
            m = ds.alive.astype(bool)
            assert np.array_equal(ds.pos[m], os_.pos[m])
            crossed += int((owner[ds.container[m]] != owner[actors.container[m]]).sum())
    if rank == 0:
        assert crossed > 0 and ref.vanished > 0
        print(f"mg_nccl_check ok: {world} ranks, {n_actors} actors, {ticks} ticks, {ref.vanished} vanished, "
              f"boundary crossings observed, last step {sim.last_step_ms:.2f} ms", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
