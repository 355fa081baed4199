"""Oracle digests of BASELINE.json's full-size configurations, for the -m gpu tests and bench.py.

    python tests/golden/make_scale_digests.py c4        # 1024x1024 grid, 16M actors, 4-item agendas (~15 CPU-minutes)
    python tests/golden/make_scale_digests.py c3        # 256x256 grid, 1M actors

Steps the CPU oracle (oracle/griddy_oracle.cpp, pinned by the fixtures in this directory) through the same
seeded world that bench.py and tests/test_gpu_scale.py build, and at every checkpoint records
  * one SHA-256 per field of the world's state in actor-id order (living actors; pos-params by their bits), of the
    get-actors order and of the per-lane / per-junction occupancy counts   (griddy_b200/digest.py: sha_fields)
  * the partition-independent commutative digest that the device computes itself (griddy_state_digest)
into tests/golden/scale_digests.json.  The device is then compared against these without the oracle having to
run at that size inside the test suite.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (grid, actors, agenda_len, checkpoints) — the worlds of bench.py (defaults) and test_gpu_scale.py
    "c3": (256, 1_000_000, 2, [100, 200, 300]),
    "c4": (1024, 16_000_000, 4, [100, 200, 300, 325]),
}
OUT = os.path.join(ROOT, "tests", "golden", "scale_digests.json")


def main():
    import __graft_entry__ as entry
    entry.build()
    from griddy_b200 import digest, synth
    from oracle.oracle import Oracle
    name = sys.argv[1]
    grid, n_actors, agenda_len, checkpoints = CONFIGS[name]
    t0 = time.time()
    net = synth.grid_network(grid)
    actors = synth.grid_actors(net, n_actors, agenda_len=agenda_len)
    ref = Oracle(net, actors)
    print(f"[{name}] world built in {time.time() - t0:.0f}s", flush=True)
    entry_ = {"grid": grid, "actors": n_actors, "agenda_len": agenda_len, "seed": f"{synth.DEFAULT_SEED:#x}", "checkpoints": {}}
    tick = 0
    for cp in checkpoints:
        ref.step(cp - tick)
        tick = cp
        st = ref.state()
        lane_count, junction_count = ref.occupancy()
        entry_["checkpoints"][str(cp)] = {"sha": digest.sha_fields(st, lane_count, junction_count),
                                          "commutative": digest.commutative(st), "vanished": ref.vanished}
        print(f"[{name}] tick {cp}: {entry_['checkpoints'][str(cp)]['commutative']} after {time.time() - t0:.0f}s", flush=True)
        data = {}   # written after every checkpoint: a long run that is cut short keeps what it has
        if os.path.exists(OUT):
            with open(OUT) as f:
                data = json.load(f)
        data[name] = entry_
        with open(OUT, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
    print(f"[{name}] written to {OUT}")


if __name__ == "__main__":
    main()
