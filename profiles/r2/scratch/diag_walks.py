"""How long are the list walks of a lane change in the bench world?  From two consecutive read-backs."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", ".."))
from griddy_b200 import device, synth

grid, n = int(os.environ.get("G", 1024)), int(os.environ.get("A", 16_000_000))
net = synth.grid_network(grid)
actors = synth.grid_actors(net, n, agenda_len=4)
sim = device.Simulation(net, actors, device=0)
sim.step(int(os.environ.get("T", 320)))
s0 = sim.state(with_order=False)
sim.step(1)
s1 = sim.state(with_order=False)
a0, a1 = s0.alive.astype(bool), s1.alive.astype(bool)
on0, on1 = a0 & (s0.loc_kind == 1), a1 & (s1.loc_kind == 1)
print("alive", a0.sum(), a1.sum(), "on road", on0.sum(), on1.sum(), "pos<=0 on road", (on0 & (s0.pos <= 0)).sum())
ent = on1 & ((s0.container != s1.container) | ~on0)            # entered a lane this tick (incl. begin-route)
print("entrants", ent.sum(), " of which from a lane", (ent & on0).sum(), " leavers to off-road", (on0 & a1 & ~on1).sum(),
      " died", (a0 & ~a1).sum())
# rank of every on-road actor inside its lane at t+1 (0 = rearmost)
idx = np.flatnonzero(on1)
order = idx[np.lexsort((s1.pos[idx], s1.container[idx]))]
c = s1.container[order]
first = np.r_[True, c[1:] != c[:-1]]
start = np.maximum.accumulate(np.where(first, np.arange(len(c)), 0))
rank = np.empty(len(s1.pos), np.int64); rank[order] = np.arange(len(c)) - start
r = rank[ent]
print("k_link walk (residents behind the entrant at t+1): mean %.2f  p50 %d  p90 %d  p99 %d  max %d  zero %.2f" %
      (r.mean(), np.percentile(r, 50), np.percentile(r, 90), np.percentile(r, 99), r.max(), (r == 0).mean()))
# residents at pos <= 0 per lane at t (what turn_onto's search for the rearmost positive key skips)
neg = on0 & (s0.pos <= 0)
cnt = np.bincount(s0.container[neg], minlength=net.n_items)
tgt = s1.container[ent & on0]
w = cnt[tgt]
print("turn_onto walk (residents at pos <= 0 on the target lane at t): mean %.2f  p50 %d  p90 %d  p99 %d  max %d  zero %.2f" %
      (w.mean(), np.percentile(w, 50), np.percentile(w, 90), np.percentile(w, 99), w.max(), (w == 0).mean()))
lanes_entered = len(np.unique(s1.container[ent]))
print("lanes entered", lanes_entered, " entrants per entered lane %.3f" % (ent.sum() / lanes_entered))
per_lane = np.bincount(s1.container[on1], minlength=net.n_items)
print("residents per entered lane at t+1: mean %.2f  p90 %d" % (per_lane[np.unique(s1.container[ent])].mean(), np.percentile(per_lane[np.unique(s1.container[ent])], 90)))
