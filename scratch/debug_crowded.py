import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import crowded_grid
from griddy_b200 import device
from oracle.oracle import Oracle
net, actors = crowded_grid(4, 800, n_dest=3, seed=7)
sim, ref = device.Simulation(net, actors), Oracle(net, actors)
prev_d = sim.state(); prev_o = ref.state()
for t in range(1, 200):
    sim.step(1); ref.step(1)
    d, o = sim.state(), ref.state()
    fields = ["alive","loc_kind","container","side","agenda_cur","sleep_left","route_status","route_head","pos"]
    bad = [f for f in fields if not np.array_equal(getattr(d,f)[o.alive.astype(bool) & d.alive.astype(bool)], getattr(o,f)[o.alive.astype(bool) & d.alive.astype(bool)])]
    if not np.array_equal(d.alive, o.alive) or bad:
        print("tick", t, "bad fields", bad)
        ids = np.nonzero(d.alive != o.alive)[0]
        print("alive differs for", ids, "dev", d.alive[ids], "orc", o.alive[ids])
        for i in ids:
            print(" actor", i, "prev: kind", prev_o.loc_kind[i], "cont", prev_o.container[i], "pos", prev_o.pos[i], "rs", prev_o.route_status[i], "head", prev_o.route_head[i])
            print("   dev now: cont", d.container[i], "pos", repr(d.pos[i]), " orc now: cont", o.container[i], "pos", repr(o.pos[i]))
            # who shares (container,pos) in oracle / device new state
            for nm, s in (("dev", d), ("orc", o)):
                same = np.nonzero((s.container == s.container[i]) & (s.pos == s.pos[i]) & (s.loc_kind == 1))[0]
                print("   ", nm, "same key:", same, "alive", s.alive[same], "prev cont", prev_o.container[same], "prev pos", prev_o.pos[same], "prev kind", prev_o.loc_kind[same])
            lane = o.container[i]
            mem = np.nonzero((prev_o.container == lane) & (prev_o.loc_kind == 1) & (prev_o.alive == 1))[0]
            print("    prev members of lane", lane, "kind", net.kind[lane], ":", mem, prev_o.pos[mem])
            memn = np.nonzero((o.container == lane) & (o.loc_kind == 1) & (o.alive == 1))[0]
            print("    new members (orc):", memn, o.pos[memn])
            memd = np.nonzero((d.container == lane) & (d.loc_kind == 1) & (d.alive == 1))[0]
            print("    new members (dev):", memd, d.pos[memd])
        break
    prev_d, prev_o = d, o
else:
    print("no mismatch")
