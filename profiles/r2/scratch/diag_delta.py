"""Why does a delta frame carry slots whose float x,y did not change?  Category counts per frame."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "..", "tests"))
from griddy_b200 import device, synth
from helpers import crowded_grid

net, actors = crowded_grid(8, 6000, agenda_len=4, seed=11, max_sleep=10)
net.geom = synth.grid_geometry(8)
sim = device.Simulation(net, actors, reorder_period=16)
bufs = [device.DeltaFrame(actors.n), device.DeltaFrame(actors.n)]
full = np.full((actors.n, 2), np.nan, np.float32)
sim.step(30)
prev = None
for k in range(8):
    sim.delta_begin(bufs[k & 1])
    ls = sim.local_state(with_order=False)
    xy64, gid = sim.positions()
    fr = sim.delta_end()
    slots, vals = fr.decode()
    before = full.copy()
    fr.apply(full)
    n = fr.n_slots
    cur = {f: ls[f][:n].copy() for f in ("gid", "alive", "loc_kind", "container", "side", "pos")}
    cur["xy64"] = xy64[:n].copy()
    if prev is not None and not fr.key and len(prev["pos"]) == n:
        sent = np.zeros(n, bool); sent[slots] = True
        posbits = prev["pos"].view(np.uint64) != cur["pos"].view(np.uint64)
        cont = (prev["container"] != cur["container"]) | (prev["side"] != cur["side"]) | (prev["loc_kind"] != cur["loc_kind"])
        vis = np.isnan(prev["xy64"][:, 0]) != np.isnan(cur["xy64"][:, 0])
        xy32 = np.any((before[:n] != full[:n]) & ~(np.isnan(before[:n]) & np.isnan(full[:n])), axis=1)
        xy64c = np.any((prev["xy64"] != cur["xy64"]) & ~(np.isnan(prev["xy64"]) & np.isnan(cur["xy64"])), axis=1)
        print(f"frame {k} tick {sim.tick}: sent {sent.sum()}  pos-bits changed {posbits.sum()}  container/side changed {cont.sum()}  "
              f"visibility changed {vis.sum()}  xy32 changed {xy32.sum()}  xy64 changed {xy64c.sum()}  "
              f"sent&~(pos|cont|vis) {(sent & ~(posbits | cont | vis)).sum()}  (pos|cont|vis)&~sent {((posbits | cont | vis) & ~sent).sum()}")
        odd = np.flatnonzero(sent & ~xy32)[:5]
        for s in odd:
            print("   slot", s, "gid", cur["gid"][s], "alive", prev["alive"][s], cur["alive"][s], "kind", prev["loc_kind"][s], cur["loc_kind"][s],
                  "container", prev["container"][s], cur["container"][s], "pos", prev["pos"][s].hex(), cur["pos"][s].hex(), "xy", prev["xy64"][s], cur["xy64"][s])
    else:
        print(f"frame {k} tick {sim.tick}: key={fr.key} sent {len(slots)} of {n}")
    prev = cur
    sim.step_async(1)
sim.step(0)
