"""ncu target: the bench world (1024x1024, 16M actors, 4-item agendas) stepped past the spin-up, so that a
`ncu -k regex:... -s N -c M` capture lands on the kernels of a tick inside the driver's timed window."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
entry.build()
from griddy_b200 import device, synth
grid, n_actors = int(os.environ.get("NCU_GRID", 1024)), int(os.environ.get("NCU_ACTORS", 16_000_000))
frames = int(os.environ.get("NCU_FRAMES", 0))   # > 0: also stream that many delta frames, one tick apart (for -k regex:k_frame)
net = synth.grid_network(grid, geometry=frames > 0)
actors = synth.grid_actors(net, n_actors, agenda_len=4)
sim = device.Simulation(net, actors, device=0)
sim.step(300)
sim.step(5)      # (the dead-fraction trigger reorders at tick 300, as in bench.py)
sim.step(int(os.environ.get("NCU_TICKS", 20)))
if frames:
    bufs = [device.DeltaFrame(n_actors), device.DeltaFrame(n_actors)]
    for k in range(frames):
        sim.delta_begin(bufs[k & 1])
        fr = sim.delta_end()
        print("frame", k, "key", fr.key, "changed", fr.n_changed, "of", fr.n_slots)
        sim.step(1)
print("alive", sim.count_alive(), "slots", sim.slots_in_use, "tick", sim.tick)
