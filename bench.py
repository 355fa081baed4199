#!/usr/bin/env python
"""bench.py — actor-updates/sec of the griddy per-tick actor update on B200.

A "step" is one tick of the hot path over every actor of the world (simulate.scm:141-153).
Workload at N = 1: BASELINE.json configs[3] — synthetic 1024x1024 procedural grid, 16M actors,
mixed travel-to / sleep-for agendas (4 items) — the largest single-GPU configuration; its per-tick
working set (>1.2 GB) is ten times the 126 MB L2, so every tick streams from HBM.  The world is
first advanced `--spinup` ticks (untimed, part of set-up) so that the initial sleep-for items have
run out and the actors are on the road, which is the expensive phase of the update.

  value     device-timed throughput, state resident in HBM (CUDA events around griddy_step)
  e2e       the same metric through the C ABI with host buffers, as the reference's game loop runs it:
            every step is one `update` and one `draw` read-back — (get-pos actor) of every actor
            (simulate.scm:155-160 -> draw.scm:74-88 -> spatial.scm:125-149) computed on the device and
            copied into pinned host memory, 8 bytes per slot, the copy of frame t in flight while tick
            t+1 runs (griddy_positions_begin / _end)
  roofline  the whole tick against the measured HBM copy bandwidth, 80 algorithmic bytes per actor-update
            (SURVEY 8(d)); per-kernel times from CUDA events (griddy_profile_step); the cost of a reorder of
            the slots is measured separately and given amortised over its period
  state_digest  partition-independent digest of the world right after the timed region (griddy_state_digest),
            with the oracle's digest of the same world at the same tick where tests/golden/scale_digests.json has it
  cpu_baseline / --impl reference   the CPU oracle (C++ restatement of the reference; Guile is not in
            this image) on one host core, on a bounded sample of the same workload: the same actors per
            junction on a 256x256 grid

Launch: python bench.py --gpus N --steps K --warmup W  (N>1 under torchrun, one rank per GPU).
N > 1: ONE world partitioned into horizontal strips of junction rows; every rank generates, uploads and holds only
its own rows of the network plus the ring next to them.  Default: weak scaling (N x --actors actors on a
round(--grid * sqrt(N))^2 grid).  --same-world: the --grid / --actors world itself on N GPUs (strong scaling;
BASELINE configs[3] "1 vs 2 B200").  --world-grid / --world-actors: any one world, e.g. configs[4]:
--gpus 8 --world-grid 4096 --world-actors 100000000.
"""
import argparse
import json
import math
import os
import resource
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_UPDATE = 80   # 56 mutable R+W, 8 speed, 16 compulsory gathers
# k_advance's share of those bytes (DESIGN.md "Kernels"): st+ahead 8, pos read 8 + write 8, delta 8,
# buf 8, turn-onto junction 4; the leader's pos-param is adjacent in slot order (0 extra)
ADVANCE_BYTES_PER_ACTOR = 44


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, repr(e)

    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
             0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def run(self):
        if not self.nv:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.NAMES.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---- CPU legs: the oracle (test infrastructure) as the reference's stand-in ---------------------------------
def cpu_sample_shape(args):
    """Bounded CPU sample of the workload: the same actors per junction on a --cpu-grid^2 grid."""
    n = max(1, int(round(args.actors * (args.cpu_grid * args.cpu_grid) / (args.grid * args.grid))))
    return args.cpu_grid, n


def cpu_sample_world(args, seed_offset=0):
    from griddy_b200 import synth
    grid, n = cpu_sample_shape(args)
    t0 = time.time()
    net = synth.grid_network(grid)
    actors = synth.grid_actors(net, n, seed=synth.DEFAULT_SEED + seed_offset, agenda_len=args.agenda_len)
    log(f"[bench] CPU sample: {grid}x{grid} grid ({net.n_items} items), {n} actors generated in {time.time() - t0:.1f}s")
    return net, actors, n


def time_oracle_world(net, actors, spinup, ticks):
    """Living-actor updates per second of the oracle (vanished actors are not advanced, core.scm:173-178)."""
    from oracle.oracle import Oracle
    o = Oracle(net, actors)
    o.step(spinup)
    a0 = int(o.state().alive.sum())
    t0 = time.perf_counter()
    o.step(ticks)
    dt = time.perf_counter() - t0
    a1 = int(o.state().alive.sum())
    return 0.5 * (a0 + a1) * ticks / dt, dt, a1


def cpu_size_table(args):
    """The oracle at the sizes BASELINE.md names: the reference's two shipped single-actor examples and a 64x64 grid
    (the cpu_baseline itself is 256x256 / 1M).  Single-threaded, like the reference (readme.md:23-26)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from griddy_b200 import synth
    rows = []
    try:
        import helpers   # tests/helpers.py: the reference's examples/simple.scm and triangle.scm through the API mirror
        for name, flat, ticks in (("C1 examples/simple.scm", helpers.simple_flat, 20000),
                                  ("C2 examples/triangle.scm", helpers.triangle_flat, 20000)):
            net, actors = flat()
            v, dt, _ = time_oracle_world(net, actors, 0, ticks)
            rows.append({"config": name, "items": net.n_items, "actors": actors.n, "ticks": ticks, "actor_updates_per_s": v,
                         "ticks_per_s": ticks / dt})
    except Exception as e:  # noqa: BLE001
        log(f"[bench] C1/C2 CPU legs skipped: {e!r}")
    n64 = max(1, int(round(args.actors * 64 * 64 / (args.grid * args.grid))))
    net = synth.grid_network(64)
    actors = synth.grid_actors(net, n64, agenda_len=args.agenda_len)
    v, dt, _ = time_oracle_world(net, actors, args.spinup, 400)
    rows.append({"config": "64x64 grid, same actors per junction", "items": net.n_items, "actors": n64, "ticks": 400,
                 "actor_updates_per_s": v, "ticks_per_s": 400 / dt})
    return rows


def sample_text(args, n, extra=""):
    grid, _ = cpu_sample_shape(args)
    return (f"{grid}x{grid} grid, {n} actors — the {args.grid}x{args.grid}/{args.actors} workload's actors per junction — "
            f"{args.agenda_len}-item agendas, after {args.spinup} spin-up ticks{extra}; C++ restatement of the reference "
            f"(oracle/), not Guile, on 1 of {os.cpu_count()} host cores: the reference is single-threaded (readme.md:23-26)")


def run_reference(args):
    """--impl reference: the reference's CPU path.  Guile + Chickadee are not installed (and cannot be:
    no network), so this is the C++ oracle port, single-threaded like the reference (readme.md:23-26).
    A step is one tick of the bounded sample world; config says what was run."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as entry
    entry.build()
    from oracle.oracle import Oracle
    net, actors, n = cpu_sample_world(args)
    grid, _ = cpu_sample_shape(args)
    o = Oracle(net, actors)
    o.step(args.spinup)
    o.step(args.warmup)
    a0 = int(o.state().alive.sum())
    t0 = time.perf_counter()
    o.step(args.steps)
    dt = time.perf_counter() - t0
    a1 = int(o.state().alive.sum())
    value = 0.5 * (a0 + a1) * args.steps / dt   # living actors only
    sample = sample_text(args, n, f", {args.steps} ticks timed")
    cfg = {"workload": f"synthetic {grid}x{grid} procedural grid, {n} actors, {args.agenda_len}-item sleep-for/travel-to agendas, "
                       f"grid routing table, {args.spinup} spin-up ticks — a bounded sample of the {args.grid}x{args.grid} / "
                       f"{args.actors}-actor workload of the GPU arm (same actors per junction)",
           "grid": grid, "actors": n, "agenda_len": args.agenda_len, "spinup_ticks": args.spinup,
           "sample_of": {"grid": args.grid, "actors": args.actors}, "parallelism": "1 host thread"}
    print(json.dumps({
        "impl": "reference", "metric": "actor-updates/sec", "value": value, "unit": "actor-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": "actor-updates/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "actor-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ---- the world of the GPU arm ------------------------------------------------------------------------------
def world_shape(args, world):
    """(grid n, total actors, scaling) of the ONE world that `world` GPUs share."""
    if args.world_grid:
        return args.world_grid, args.world_actors or args.actors * world, "strong"
    if args.same_world or world == 1:
        return args.grid, args.actors, ("strong" if world > 1 else "weak")
    return int(round(args.grid * math.sqrt(world))), args.actors * world, "weak"


def workload_config(args, world):
    n, total, _ = world_shape(args, world)
    cfg = {"workload": f"synthetic {n}x{n} procedural grid, {total} actors, "
                       f"{args.agenda_len}-item sleep-for/travel-to agendas, grid routing table, "
                       f"{args.spinup} spin-up ticks",
           "grid": n, "actors": total, "actors_per_gpu": total // world, "agenda_len": args.agenda_len,
           "spinup_ticks": args.spinup, "l2": "working set per tick >> 126 MB L2 (no flush needed)",
           "parallelism": (f"{world} ranks, {world * args.strips_per_rank} horizontal strips of junction rows of ONE world dealt to "
                           f"the ranks in turn; every rank holds only its rows of the network and the ring next to them. "
                           f"Data plane: peer memory — per tick the halo (boundary lanes' rearmost pos-param, boundary "
                           f"junctions' occupancy) and the actors that crossed are written by the sender's kernels straight "
                           f"into the neighbour's window over NVLink and flagged; no NCCL call, no message, no host work per "
                           f"tick (torch.distributed/NCCL only carries the cudaIpc handles at set-up and this script's "
                           f"barriers)" if world > 1 else "1 gpu")}
    return cfg


def golden_digest(args, world, tick):
    """The oracle's digest of this very world at this tick, if tests/golden/scale_digests.json holds it."""
    path = os.path.join(ROOT, "tests", "golden", "scale_digests.json")
    if not os.path.exists(path):
        return None
    n, total, _ = world_shape(args, world)
    with open(path) as f:
        for entry in json.load(f).values():
            if (entry["grid"], entry["actors"], entry["agenda_len"]) == (n, total, args.agenda_len):
                cp = entry["checkpoints"].get(str(tick))
                return cp["commutative"] if cp else None
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="griddy_b200", choices=["griddy_b200", "reference"])
    ap.add_argument("--grid", type=int, default=1024)
    ap.add_argument("--actors", type=int, default=16_000_000)
    ap.add_argument("--agenda-len", type=int, default=4)
    ap.add_argument("--spinup", type=int, default=300)
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--profile-ticks", type=int, default=50)
    ap.add_argument("--cpu-grid", type=int, default=256)
    ap.add_argument("--cpu-ticks", type=int, default=60)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--reorder-period", type=int, default=384)
    ap.add_argument("--same-world", action="store_true", help="multi-GPU: the --grid / --actors world itself on all GPUs (strong scaling)")
    ap.add_argument("--world-grid", type=int, default=0, help="multi-GPU: junctions per side of the one world (overrides the scaling rule)")
    ap.add_argument("--world-actors", type=int, default=0, help="multi-GPU: actors of the one world (with --world-grid)")
    ap.add_argument("--strips-per-rank", type=int, default=0,
                    help="multi-GPU: strips of junction rows per rank, dealt in turn (load balance); 0 = 2 beyond two GPUs")
    ap.add_argument("--mig-cap", type=int, default=1 << 24, help="multi-GPU: boundary crossings per neighbour per tick")
    ap.add_argument("--extra-slots", type=int, default=4_000_000, help="multi-GPU: room for actors arriving from other ranks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.strips_per_rank <= 0:   # edge strips drain towards the middle of the grid: give every rank some of both
        args.strips_per_rank = 2 if args.gpus > 2 else 1
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    try:   # run next to this GPU: the pinned frame buffers of the e2e leg are then allocated on its NUMA node
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        log(f"[bench] rank {rank}: bound to the CPUs next to GPU {local_rank} ({len(os.sched_getaffinity(0))} of {os.cpu_count()})")
    except Exception as e:  # noqa: BLE001
        log(f"[bench] rank {rank}: CPU affinity left alone ({e!r})")
    real_stdout = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        real_stdout = os.dup(1)      # NCCL's banner lines go to fd 1: keep it for the one JSON line, send the rest to stderr
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    from griddy_b200 import device, digest, synth

    n_grid, total_actors, scaling = world_shape(args, world)
    t0 = time.time()
    if world == 1:
        net = synth.grid_network(args.grid, geometry=True)
        actors = synth.grid_actors(net, args.actors, agenda_len=args.agenda_len)
        log(f"[bench] generated {args.grid}x{args.grid} grid ({net.n_items} items), {args.actors} actors in {time.time() - t0:.1f}s")
        sim = device.Simulation(net, actors, device=local_rank, reorder_period=args.reorder_period)
        n_items_world, held = net.n_items, net.n_items
    else:
        # ONE world, partitioned into horizontal strips; this rank generates, uploads and holds only its own rows
        # of the network (plus the ring next to them) and its own actors
        rb, re = synth.tile_rows(n_grid, world, args.strips_per_rank, rank)
        net = synth.grid_tile(n_grid, rb, re, world, args.strips_per_rank, geometry=True)
        actors, ids = synth.grid_actors_of_rank(n_grid, total_actors, world, args.strips_per_rank, rank, agenda_len=args.agenda_len)
        n_items_world, held = net.n_global, net.n_items
        log(f"[bench] rank {rank}: {n_grid}x{n_grid} grid, holding {net.n_items} of {net.n_global} items (rows "
            f"{list(zip(rb.tolist(), re.tolist()))}), {actors.n} of {total_actors} actors, generated in {time.time() - t0:.1f}s")
        sim = device.Simulation(net, actors, device=local_rank, reorder_period=args.reorder_period,
                                partition=(rank, world, net.owner, args.mig_cap, args.extra_slots), gids=ids,
                                max_agenda_items=args.agenda_len)
        device.connect_ipc(sim)
    net.geom = None   # on the device now
    log(f"[bench] rank {rank}: uploaded in {time.time() - t0:.1f}s")
    sim.step(args.spinup)
    log(f"[bench] rank {rank}: spin-up {args.spinup} ticks in {sim.last_step_ms:.1f} ms")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(obj):
        if world == 1:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    sim.step(args.warmup)
    alive0 = sim.count_alive()
    launches0 = sim.kernel_launches
    reorders0 = sim.reorders
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    w0 = time.perf_counter()
    sim.step(args.steps)            # K ticks, CUDA events on the launching stream inside
    torch.cuda.synchronize()
    wall = time.perf_counter() - w0
    clocks = sampler.stop()
    barrier()
    dev_ms = sim.last_step_ms
    launches = sim.kernel_launches - launches0
    reorders_timed = sim.reorders - reorders0
    tick_after = sim.tick
    # outside the timed region: the state the timed ticks produced, as a partition-independent digest
    state_digest = digest.combine(gather(sim.digest()))
    want = golden_digest(args, world, tick_after)
    state_digest.update({"tick": tick_after, "oracle": want, "matches_oracle": (None if want is None else
                         all(state_digest[k] == want[k] for k in ("sum", "xor", "alive")))})
    if os.environ.get("GRIDDY_DUMP_TRACE"):   # -DGRIDDY_TRACE builds: the last 64 ticks' kernel start stamps
        stamps = np.zeros((64, 16), np.uint64)
        rc = device.lib().griddy_debug_trace(sim.h, stamps.ctypes.data)
        if rc == 0:
            np.save(os.path.join(os.environ["GRIDDY_DUMP_TRACE"], f"trace_n{world}_rank{rank}.npy"), stamps)
            log(f"[bench] rank {rank}: trace written")
    alive1 = sim.count_alive()
    # an actor replaced by an equal-key insert is gone (core.scm:173-178) and is not advanced any
    # more: count the living ones only (mean of before / after; the decline is slow and steady)
    alive_mean = 0.5 * (alive0 + alive1)
    per_rank = None
    if world > 1:
        t = torch.tensor([dev_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
        t = torch.zeros(world, 2, device="cuda", dtype=torch.float64)
        t[rank, 0], t[rank, 1] = alive0, alive1
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        per_rank = {"alive_before": [int(v) for v in t[:, 0].tolist()], "alive_after": [int(v) for v in t[:, 1].tolist()]}
        alive_total = float(t.mean(dim=1).sum().item())
    else:
        alive_total = alive_mean
    value = alive_total * args.steps / (dev_ms * 1e-3)

    # per-kernel device time (events around every launch; rank-local), then one forced reorder, timed
    kms = sim.profile_step(args.profile_ticks)
    alive_prof = sim.count_alive()
    kpt = {k: v / args.profile_ticks for k, v in kms.items()}
    reorder_events = kpt.pop("reorder")   # event overhead unless a reorder fell into the window
    tick_ms_prof = sum(kpt.values())
    reorder_ms = sim.timed_reorder() if args.reorder_period > 0 else None
    peak, peak_src = measured_peaks()
    gbs = lambda nbytes, ms: nbytes / (ms * 1e-3) / 1e9 if ms > 0 else None
    tick_ms = dev_ms / args.steps
    whole_tick = gbs(alive_total / world * ALGO_BYTES_PER_UPDATE, tick_ms)   # per GPU
    amortised = (gbs(alive_total / world * ALGO_BYTES_PER_UPDATE, tick_ms + reorder_ms / args.reorder_period)
                 if reorder_ms is not None else whole_tick)
    adv = gbs(alive_prof * ADVANCE_BYTES_PER_ACTOR, kpt["k_advance"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("tick_dram_bytes")

    # e2e: what the reference's game loop does per tick — `update`, then `draw` reads (get-pos actor) of
    # every actor.  Frame t is formatted on the device and copied into pinned host memory while tick t+1
    # runs; every frame is complete on the host inside the timed region.
    cap = actors.n if world == 1 else actors.n + args.extra_slots
    frames = [device.PinnedFrame(cap), device.PinnedFrame(cap)]
    dframes = [device.DeltaFrame(cap), device.DeltaFrame(cap)]

    copy_ms = []
    fmt_ms = []   # device time of the delta frames' format kernel (events around it; runs beside the previous frame's copy)

    def e2e_full_loop(steps):
        """Full frames: float x, y of every slot, every tick (griddy_positions_begin / _end)."""
        slots = 0
        sim.step_async(1)
        sim.frame_begin(frames[0])
        for k in range(1, steps):
            sim.step_async(1)
            sim.frame_begin(frames[k & 1])
            slots += len(sim.frame_end())        # frame k-1 is on the host
        last = sim.frame_end()
        sim.step(0)                              # the device's error word, checked once per batch of frames
        return 8 * (slots + len(last)), 0, 0

    def e2e_delta_loop(steps):
        """Delta frames: per tick the slots whose drawn position changed (griddy_delta_begin / _end)."""
        copied = changed = keys = 0
        fmt_ms.clear()
        copy_ms.clear()
        sim.step_async(1)
        sim.delta_begin(dframes[0])
        for k in range(1, steps + 1):
            if k < steps:
                sim.step_async(1)
                sim.delta_begin(dframes[k & 1])
            fr = sim.delta_end()                 # frame k-1 is on the host
            copied += fr.bytes_copied
            changed += fr.n_changed
            keys += int(fr.key)
            if not fr.key:
                fmt_ms.append(sim.last_frame_ms)
                copy_ms.append(sim.last_copy_ms)
        sim.step(0)
        return copied, changed, keys

    def timed_e2e(loop):
        loop(3)
        barrier()
        e0 = time.perf_counter()
        copied, changed, keys = loop(args.e2e_steps)
        e_dt = time.perf_counter() - e0
        alive_now = sim.count_alive()
        t = torch.tensor([e_dt, float(alive_now), copied / args.e2e_steps, changed / args.e2e_steps, float(keys)],
                         device="cuda", dtype=torch.float64)
        if world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t[0] = tmax[0]
        e_dt, alive_all, d2h_all, changed_all, keys_all = (float(v) for v in t.tolist())
        return {"value": alive_all * args.e2e_steps / e_dt, "d2h": int(d2h_all), "changed": changed_all,
                "alive": int(alive_all), "keys": int(keys_all), "ms_per_step": 1e3 * e_dt / args.e2e_steps}

    e2e_full = timed_e2e(e2e_full_loop)
    e2e_delta = timed_e2e(e2e_delta_loop)
    rss = gather(round(resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6, 2))
    held_all = gather(held)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    cpu = sizes = None
    if not args.no_cpu_baseline and world == 1:
        net_c, actors_c, n_cpu = cpu_sample_world(args)
        v, dt, _ = time_oracle_world(net_c, actors_c, args.spinup, args.cpu_ticks)
        cpu = {"value": v, "unit": "actor-updates/s", "cores": 1, "kind": "port",
               "sample": sample_text(args, n_cpu, f", {args.cpu_ticks} ticks timed in {dt:.1f}s")}
        try:
            sizes = cpu_size_table(args)
            sizes.append({"config": f"{args.cpu_grid}x{args.cpu_grid} grid, same actors per junction (the cpu_baseline)",
                          "actors": n_cpu, "ticks": args.cpu_ticks, "actor_updates_per_s": v, "ticks_per_s": args.cpu_ticks / dt})
        except Exception as e:  # noqa: BLE001  (an extra, never worth losing the line for)
            log(f"[bench] CPU size table failed: {e!r}")
    out = {
        "metric": "actor-updates/sec", "value": value, "unit": "actor-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tick_ms,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": e2e_delta["value"], "unit": "actor-updates/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": e2e_delta["d2h"], "steps": args.e2e_steps, "ms_per_step": e2e_delta["ms_per_step"],
                "living_actors": e2e_delta["alive"], "positions_changed_per_step": e2e_delta["changed"],
                "key_frames_in_timed_region": e2e_delta["keys"],
                "frame_kernel_ms": (sorted(fmt_ms)[len(fmt_ms) // 2] if fmt_ms else None),
                "frame_copy_ms": (sorted(copy_ms)[len(copy_ms) // 2] if copy_ms else None),
                "note": "per step: one `update` (griddy_step_async(1): no host synchronisation per tick) + one `draw` "
                        "read-back as a DELTA frame (griddy_delta_begin/_end): (get-pos actor) (spatial.scm:125-149) is "
                        "evaluated on the device for every actor whose drawn position differs from the frame before, and "
                        "those float x,y pairs plus one mask bit per slot are copied into pinned host memory — the copy of "
                        "frame t overlaps tick t+1, every frame is on the host inside the timed region, a reorder of the slots "
                        "inside it costs a key frame (all living actors).  Patching a persistent position array with a frame "
                        "(griddy_delta_apply) is the consumer's work, like uploading a vertex buffer, and is not timed; "
                        "tests/test_positions.py shows the patched array equals the full frame of every tick, bit for bit.  "
                        "e2e_full_frames is the same loop with every slot copied every tick.  No per-step input exists: the "
                        "world is uploaded once through the same ABI before the timed region"},
        "e2e_full_frames": {"value": e2e_full["value"], "unit": "actor-updates/s", "d2h_bytes_per_step": e2e_full["d2h"],
                            "ms_per_step": e2e_full["ms_per_step"], "steps": args.e2e_steps,
                            "note": "griddy_positions_begin/_end: float x,y of every slot, every tick (8 bytes per slot over PCIe)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "state_digest": state_digest,
        # The tick is four kernels of comparable duration, so the roofline is stated for the whole
        # tick: 80 algorithmic bytes per living actor over the device time of one tick
        # of the timed region.  k_advance, the kernel that streams the per-actor state, is given
        # beside it against its own 44 bytes per actor.
        "roofline": {"bound": "hbm", "kernel": "whole tick (k_advance + k_slow + k_unlink + k_link), per GPU",
                     "achieved": whole_tick, "peak": peak, "unit": "GB/s", "frac": whole_tick / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_actor_update": ALGO_BYTES_PER_UPDATE,
                     "living_actors": alive_total,
                     "with_amortised_reorder": {"achieved": amortised, "frac": amortised / peak, "reorder_ms": reorder_ms,
                                                "reorder_period_ticks": args.reorder_period,
                                                "reorders_in_timed_region": reorders_timed,
                                                "note": "one reorder of the slots, forced and timed with CUDA events right after "
                                                        "the profile ticks, spread over its period"},
                     "k_advance": {"algorithmic_bytes_per_actor": ADVANCE_BYTES_PER_ACTOR, "achieved": adv,
                                   "frac": adv / peak if adv else None, "ms": kpt["k_advance"]},
                     "kernel_ms_per_tick": kpt,
                     "kernel_share_of_tick": {k: v / tick_ms_prof for k, v in kpt.items()} if tick_ms_prof else None,
                     "profile_window": {"ticks": args.profile_ticks, "first_tick": tick_after,
                                        "reorder_events_ms_per_tick": reorder_events}},
        "cpu_baseline": cpu,
        "cpu_baseline_sizes": sizes,
        "world": {"alive_before": alive0, "alive_after": alive1, "actors_uploaded": total_actors,
                  "items_in_world": n_items_world, "items_held_per_rank": held_all, "host_rss_gb_per_rank": rss,
                  "wall_s_timed_region": wall,
                  "note": "value counts living actors only" + ("; alive_* are rank 0's, per_rank has every rank's: strips "
                          "at the edge of the grid drain towards the middle, and the tick is as long as the fullest rank's"
                          if world > 1 else ""),
                  "per_rank": per_rank},
    }
    line = json.dumps(out)
    if real_stdout is not None:
        os.write(real_stdout, (line + "\n").encode())
    else:
        print(line, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
