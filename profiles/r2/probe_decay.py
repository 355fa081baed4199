"""Round-2 probe (run on a B200 via gpurun): how the tick's kernels slow down with ticks since the last reorder,
what a reorder costs, and which fraction of the living actors changes its drawn position per tick.
Output: gpurun_out/r2_probe_decay.json"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
entry.build()
from griddy_b200 import device, synth

grid, n_actors = int(os.environ.get("PROBE_GRID", 1024)), int(os.environ.get("PROBE_ACTORS", 16_000_000))
net = synth.grid_network(grid, geometry=True)
actors = synth.grid_actors(net, n_actors, agenda_len=4)
out = {"grid": grid, "actors": n_actors}
sim = device.Simulation(net, actors, device=0, reorder_period=0)   # no periodic reorder: we trigger it by hand
L = device.lib()
import ctypes as C
sim.step(300)
# force a reorder now: period = 1 for one tick
def reorder_now():
    L.griddy_set_reorder_period(sim.h, C.c_int32(1))
    k = sim.profile_step(1)
    L.griddy_set_reorder_period(sim.h, C.c_int32(0))
    return k
k = reorder_now()
out["reorder_ms_at_300"] = k["reorder"]
curve = []
for t in range(0, 400, 4):
    k = sim.profile_step(4)
    curve.append({"since_reorder": t + 1, "slots": sim.slots_in_use, **{a: b / 4 for a, b in k.items()}})
out["curve_after_300"] = curve
out["alive_700"] = sim.count_alive()
k = reorder_now()
out["reorder_ms_at_700"] = k["reorder"]
curve = []
for t in range(0, 200, 4):
    k = sim.profile_step(4)
    curve.append({"since_reorder": t + 1, "slots": sim.slots_in_use, **{a: b / 4 for a, b in k.items()}})
out["curve_after_700"] = curve
# moving fraction
mv = []
for rep in range(3):
    xy0, gid0 = sim.positions(f32=True)
    sim.step(1)
    xy1, gid1 = sim.positions(f32=True)
    live = (gid0 >= 0) & (gid1 >= 0) & (gid0 == gid1)
    moved = live & np.any(xy0 != xy1, axis=1)
    mv.append({"tick": sim.tick, "live": int(live.sum()), "moved": int(moved.sum())})
    sim.step(50)
out["moving"] = mv
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "r2_probe_decay.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps({k: v for k, v in out.items() if not k.startswith("curve")}))
for name in ("curve_after_300", "curve_after_700"):
    for c in out[name][::5]:
        print(name, c)
