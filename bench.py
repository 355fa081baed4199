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
            every step is one `update` (griddy_step(1)) and one `draw` read-back — (get-pos actor) of every
            actor (simulate.scm:155-160 -> draw.scm:74-88 -> spatial.scm:125-149) computed on the device
            and copied into pinned host memory, 8 bytes per slot, the copy of frame t in flight while tick
            t+1 runs (griddy_positions_begin / _end).  e2e_full_state: the round-1 definition, the whole
            per-actor state (15 bytes per actor) fetched synchronously after every tick
  roofline  k_advance against the measured HBM copy bandwidth, 80 algorithmic bytes per actor-update
            (SURVEY 8(d)), timed with per-kernel CUDA events (griddy_profile_step)
  cpu_baseline / --impl reference   the CPU oracle (C++ restatement of the reference; Guile is not in
            this image) on one host core, on a bounded sample with the same actors per segment

Launch: python bench.py --gpus N --steps K --warmup W  (N>1 under torchrun, one rank per GPU).
"""
import argparse
import json
import math
import os
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


def build_world(grid, n_actors, agenda_len, seed_offset=0, geometry=False):
    from griddy_b200 import synth
    t0 = time.time()
    net = synth.grid_network(grid, geometry=geometry)
    actors = synth.grid_actors(net, n_actors, seed=synth.DEFAULT_SEED + seed_offset, agenda_len=agenda_len)
    log(f"[bench] generated {grid}x{grid} grid ({net.n_items} items), {n_actors} actors in {time.time() - t0:.1f}s")
    return net, actors


def cpu_sample_world(args, seed_offset=0):
    """Bounded CPU sample: a 64x64 grid with the same actors per segment and agenda shape."""
    from griddy_b200 import synth
    sz_full, sz = synth.grid_sizes(args.grid), synth.grid_sizes(args.cpu_grid)
    n = max(1, int(round(args.actors * sz["n_segments"] / sz_full["n_segments"])))
    net, actors = build_world(args.cpu_grid, n, args.agenda_len, seed_offset=seed_offset)
    return net, actors, n


def time_oracle(args, ticks, seed_offset=0):
    """Living-actor updates per second of the oracle (vanished actors are not advanced, core.scm:173-178)."""
    from oracle.oracle import Oracle
    net, actors, n = cpu_sample_world(args, seed_offset)
    o = Oracle(net, actors)
    o.step(args.spinup)
    a0 = int(o.state().alive.sum())
    t0 = time.perf_counter()
    o.step(ticks)
    dt = time.perf_counter() - t0
    a1 = int(o.state().alive.sum())
    return 0.5 * (a0 + a1) * ticks / dt, dt, n, a1


def time_oracle_all_cores(args, ticks):
    """The same sample on every host core at once: independent replicas of the single-threaded loop (different
    actors each), one process per core.  The reference cannot use more than one core for one world
    (readme.md:23-26); this is what the box's CPUs deliver in total, for scale."""
    import subprocess
    cores = len(os.sched_getaffinity(0))
    cmd = [sys.executable, os.path.abspath(__file__), "--grid", str(args.grid), "--actors", str(args.actors),
           "--agenda-len", str(args.agenda_len), "--spinup", str(args.spinup), "--cpu-grid", str(args.cpu_grid),
           "--cpu-ticks", str(ticks)]
    procs = [subprocess.Popen(cmd + ["--oracle-worker", str(k + 1)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
             for k in range(cores)]
    updates, slowest = 0.0, 0.0
    for pr in procs:
        out, _ = pr.communicate(timeout=600)
        r = json.loads(out.strip().splitlines()[-1])
        updates += r["updates"]
        slowest = max(slowest, r["seconds"])
    return updates / slowest, slowest, cores


def run_reference(args):
    """--impl reference: the reference's CPU path.  Guile + Chickadee are not installed (and cannot be:
    no network), so this is the C++ oracle port, single-threaded like the reference (readme.md:23-26)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # size the per-step sample so the whole run stays within a couple of minutes
    ticks_per_step = 10
    for _ in range(args.warmup):
        pass
    from oracle.oracle import Oracle
    net, actors, n = cpu_sample_world(args)
    o = Oracle(net, actors)
    o.step(args.spinup)
    o.step(args.warmup * ticks_per_step)
    steps = min(args.steps, 50)
    a0 = int(o.state().alive.sum())
    t0 = time.perf_counter()
    o.step(steps * ticks_per_step)
    dt = time.perf_counter() - t0
    a1 = int(o.state().alive.sum())
    value = 0.5 * (a0 + a1) * steps * ticks_per_step / dt   # living actors only
    sample = (f"{args.cpu_grid}x{args.cpu_grid} grid, {n} actors (same actors per segment as the "
              f"{args.grid}x{args.grid}/{args.actors} workload), {ticks_per_step} ticks per step after "
              f"{args.spinup} spin-up ticks; C++ restatement of the reference, not Guile")
    print(json.dumps({
        "impl": "reference", "metric": "actor-updates/sec", "value": value, "unit": "actor-updates/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "actor-updates/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "actor-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def world_shape(args, world):
    """Weak scaling: every GPU gets args.grid^2 junctions and args.actors actors of ONE world whose
    junction grid is square, n = round(args.grid * sqrt(world)).  The static network is replicated
    on every rank, so the tile is shrunk if the box's host memory could not hold `world` copies."""
    per_gpu = args.grid
    note = None
    while world > 1:
        n = int(round(per_gpu * math.sqrt(world)))
        need = world * 30 * n * n * 170          # ~30 items per junction, ~100 host bytes per item + 64 of geometry
        try:
            avail = next(int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable"))
        except Exception:  # noqa: BLE001
            avail = None
        if avail is None or need < 0.5 * avail or per_gpu <= 128:
            break
        per_gpu //= 2
        note = f"tile shrunk to {per_gpu}^2 junctions per GPU: host memory ({avail >> 30} GiB) cannot hold {world} network replicas"
    n = int(round(per_gpu * math.sqrt(world))) if world > 1 else args.grid
    actors = args.actors * (per_gpu * per_gpu) // (args.grid * args.grid)
    return n, actors * world, note


def workload_config(args, world=None):
    world = world or args.gpus
    n, total, note = world_shape(args, world)
    cfg = {"workload": f"synthetic {n}x{n} procedural grid, {total} actors, "
                       f"{args.agenda_len}-item sleep-for/travel-to agendas, grid routing table, "
                       f"{args.spinup} spin-up ticks",
           "grid": n, "actors": total, "actors_per_gpu": total // world, "agenda_len": args.agenda_len,
           "spinup_ticks": args.spinup, "l2": "working set per tick >> 126 MB L2 (no flush needed)",
           "parallelism": (f"{world} ranks, {world * args.strips_per_rank} horizontal strips of junction rows of one world dealt "
                           f"to the ranks in turn; per tick the halo "
                           f"(boundary lanes' rearmost pos-param, boundary junctions' occupancy) and the actors that "
                           f"crossed are exchanged with grouped ncclSend/ncclRecv" if world > 1 else "1 gpu")}
    if note:
        cfg["note"] = note
    return cfg


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
    ap.add_argument("--cpu-grid", type=int, default=64)
    ap.add_argument("--cpu-ticks", type=int, default=1500)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--oracle-worker", type=int, default=0, help=argparse.SUPPRESS)   # one replica of time_oracle_all_cores
    ap.add_argument("--reorder-period", type=int, default=384)
    ap.add_argument("--strips-per-rank", type=int, default=0,
                    help="multi-GPU: strips of junction rows per rank, dealt in turn (load balance); 0 = 2 beyond two GPUs")
    ap.add_argument("--mig-cap", type=int, default=1 << 24, help="multi-GPU: boundary crossings per neighbour per tick")
    ap.add_argument("--extra-slots", type=int, default=4_000_000, help="multi-GPU: room for actors arriving from other ranks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.strips_per_rank <= 0:   # edge strips drain towards the middle of the grid: give every rank some of both
        args.strips_per_rank = 2 if args.gpus > 2 else 1
    os.environ["NCCL_DEBUG"] = "WARN"   # NCCL's version banner goes to stdout, which carries the one JSON line

    if args.oracle_worker:
        v, dt, _, _ = time_oracle(args, args.cpu_ticks, seed_offset=args.oracle_worker)
        print(json.dumps({"updates": v * dt, "seconds": dt}), flush=True)
        return
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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    from griddy_b200 import device

    from griddy_b200 import synth
    n_grid, total_actors, _ = world_shape(args, world)
    t0 = time.time()
    if world == 1:
        net, actors = build_world(args.grid, args.actors, args.agenda_len, geometry=True)
        sim = device.Simulation(net, actors, device=local_rank, reorder_period=args.reorder_period)
    else:
        # ONE world, partitioned into horizontal strips; this rank generates and uploads its own actors
        net = synth.grid_network(n_grid, geometry=True)
        owner = synth.grid_owner(net, world, args.strips_per_rank)
        ids = synth.actor_ids_of_rank(net, total_actors, owner, rank)
        actors = synth.grid_actors(net, total_actors, agenda_len=args.agenda_len, ids=ids)
        log(f"[bench] rank {rank}: {n_grid}x{n_grid} grid ({net.n_items} items), {actors.n} of {total_actors} actors "
            f"generated in {time.time() - t0:.1f}s")
        sim = device.Simulation(net, actors, device=local_rank, reorder_period=args.reorder_period,
                                partition=(rank, world, owner, args.mig_cap, args.extra_slots), gids=ids)
        device.connect_ipc(sim)
    net.geom = None   # on the device now
    log(f"[bench] rank {rank}: uploaded in {time.time() - t0:.1f}s")
    sim.step(args.spinup)
    log(f"[bench] rank {rank}: spin-up {args.spinup} ticks in {sim.last_step_ms:.1f} ms")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sim.step(args.warmup)
    alive0 = sim.count_alive()
    launches0 = sim.kernel_launches
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

    # per-kernel device time (events around every launch; rank-local)
    kms = sim.profile_step(args.profile_ticks)
    alive_prof = sim.count_alive()
    kpt = {k: v / args.profile_ticks for k, v in kms.items()}
    tick_ms_prof = sum(kpt.values())
    peak, peak_src = measured_peaks()
    gbs = lambda nbytes, ms: nbytes / (ms * 1e-3) / 1e9 if ms > 0 else None
    whole_tick = gbs(alive_mean * ALGO_BYTES_PER_UPDATE, dev_ms / args.steps)
    adv = gbs(alive_prof * ADVANCE_BYTES_PER_ACTOR, kpt["k_advance"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("tick_dram_bytes")

    # e2e: what the reference's game loop does per tick — `update`, then `draw` reads (get-pos actor) of
    # every actor.  Frame t is formatted on the device and copied into pinned host memory while tick t+1
    # runs; every frame is complete on the host inside the timed region.
    import ctypes as C
    L = device.lib()
    cap = actors.n if world == 1 else actors.n + args.extra_slots
    frames = [device.PinnedFrame(cap), device.PinnedFrame(cap)]

    def e2e_loop(steps):
        slots = 0
        sim.step(1)
        sim.frame_begin(frames[0])
        for k in range(1, steps):
            sim.step(1)
            sim.frame_begin(frames[k & 1])
            slots += len(sim.frame_end())        # frame k-1 is on the host
        last = sim.frame_end()
        return slots + len(last), last

    e2e_loop(3)
    barrier()
    e0 = time.perf_counter()
    slots_read, last_frame = e2e_loop(args.e2e_steps)
    e_dt = time.perf_counter() - e0
    living_in_frame = int((~np.isnan(last_frame[:, 0])).sum())
    d2h = 8.0 * slots_read / args.e2e_steps
    alive_e2e = sim.count_alive()
    t = torch.tensor([e_dt, float(alive_e2e), d2h, float(living_in_frame)], device="cuda", dtype=torch.float64)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        e_dt = float(tmax[0].item())
    alive_e2e_total, d2h_total, living_in_frame = float(t[1].item()), float(t[2].item()), int(t[3].item())
    e2e_value = alive_e2e_total * args.e2e_steps / e_dt

    # the round-1 definition, kept for continuity: the whole per-actor state `draw` could want (alive,
    # location kind, container, road side, pos-param), fetched synchronously after every tick
    n = cap
    pin = lambda dt: torch.empty(n, dtype=dt, pin_memory=True).numpy()
    bufs = [pin(torch.uint8), pin(torch.uint8), pin(torch.int32), pin(torch.uint8), pin(torch.float64)]
    ptrs = [b.ctypes.data_as(C.c_void_p) for b in bufs] + [None] * 6
    gid_buf = pin(torch.int32) if world > 1 else None
    n_local = C.c_int64(0)

    def full_state_step():
        sim.step(1)
        if world == 1:
            rc = L.griddy_download_actors(sim.h, *ptrs)
        else:
            rc = L.griddy_download_local(sim.h, C.c_int64(n), C.byref(n_local), gid_buf.ctypes.data_as(C.c_void_p), *ptrs)
        assert rc == 0, L.griddy_last_error(sim.h)

    full_steps = max(1, min(args.e2e_steps, 5))
    full_state_step()
    barrier()
    f0 = time.perf_counter()
    for _ in range(full_steps):
        full_state_step()
    torch.cuda.synchronize()
    f_dt = time.perf_counter() - f0
    if world > 1:
        full_bytes = int(n_local.value) * (sum(b.itemsize for b in bufs) + 4)
        alive = int(bufs[0][: n_local.value].sum())
        on_road = int((bufs[0][: n_local.value] & bufs[1][: n_local.value]).sum())
        t = torch.tensor([float(alive), float(full_bytes)], device="cuda", dtype=torch.float64)
        tm = torch.tensor([f_dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        full_value, full_bytes = float(t[0].item()) * full_steps / float(tm.item()), int(t[1].item())
    else:
        full_bytes = sum(b.nbytes for b in bufs)
        alive = int(bufs[0].sum())
        on_road = int((bufs[0] & bufs[1]).sum())
        full_value = alive * full_steps / f_dt

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    cpu = cpu_all = None
    if not args.no_cpu_baseline and world == 1:
        v, dt, n_cpu, _ = time_oracle(args, args.cpu_ticks)
        cpu = {"value": v, "unit": "actor-updates/s", "cores": 1, "kind": "port",
               "sample": f"{args.cpu_grid}x{args.cpu_grid} grid, {n_cpu} actors (same actors per segment), "
                         f"{args.cpu_ticks} ticks after {args.spinup} spin-up ticks, {dt:.1f}s; C++ restatement "
                         f"of the reference on 1 of {os.cpu_count()} host cores (the reference is single-threaded; "
                         f"Guile is not installed)"}
        try:
            va, dta, cores = time_oracle_all_cores(args, max(args.cpu_ticks // 2, 1))
            cpu_all = {"value": va, "unit": "actor-updates/s", "cores": cores, "kind": "port",
                       "sample": f"{cores} independent replicas of the same sample (different actors each), one process "
                                 f"per core, {max(args.cpu_ticks // 2, 1)} ticks after {args.spinup} spin-up ticks, "
                                 f"{dta:.1f}s; the reference itself is single-threaded, so this is the box's total, not "
                                 f"one world's speed"}
        except Exception as e:  # noqa: BLE001  (an extra, never worth losing the line for)
            log(f"[bench] all-cores CPU leg failed: {e!r}")
    out = {
        "metric": "actor-updates/sec", "value": value, "unit": "actor-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": e2e_value, "unit": "actor-updates/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": int(d2h_total), "steps": args.e2e_steps,
                "living_actors_in_last_frame": living_in_frame,
                "note": "per step: one `update` (griddy_step(1)) + one `draw` read-back: (get-pos actor) of every "
                        "slot computed on the device (spatial.scm:125-149), float x,y, copied into pinned host memory "
                        "with griddy_positions_begin/_end — the copy of frame t overlaps tick t+1, every frame is on "
                        "the host inside the timed region; no per-step input exists: the world is uploaded once "
                        "through the same ABI before the timed region"},
        "e2e_full_state": {"value": full_value, "unit": "actor-updates/s", "d2h_bytes_per_step": full_bytes,
                           "steps": full_steps,
                           "note": "round-1 definition: griddy_step(1) + synchronous griddy_download_actors of alive, "
                                   "location kind, container, side, pos-param (15 bytes per actor)"},
        "gpu_launches": launches,
        "clocks": clocks,
        # The tick is four kernels of comparable duration, so the roofline is stated for the whole
        # tick: 80 algorithmic bytes per living actor over the device time of one tick
        # of the timed region.  k_advance, the kernel that streams the per-actor state, is given
        # beside it against its own 44 bytes per actor.
        "roofline": {"bound": "hbm", "kernel": "whole tick (k_advance + k_slow + k_unlink + k_link + amortised reorder)",
                     "achieved": whole_tick, "peak": peak, "unit": "GB/s", "frac": whole_tick / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_actor_update": ALGO_BYTES_PER_UPDATE,
                     "living_actors": alive_mean,
                     "k_advance": {"algorithmic_bytes_per_actor": ADVANCE_BYTES_PER_ACTOR, "achieved": adv,
                                   "frac": adv / peak if adv else None, "ms": kpt["k_advance"]},
                     "kernel_ms_per_tick": kpt,
                     "kernel_share_of_tick": {k: v / tick_ms_prof for k, v in kpt.items()} if tick_ms_prof else None,
                     "reorder_period_ticks": args.reorder_period},
        "cpu_baseline": cpu,
        "cpu_baseline_all_cores": cpu_all,
        "world": {"alive_before": alive0, "alive_after": alive1, "alive_at_end": alive, "on_road": on_road,
                  "actors_uploaded": total_actors, "wall_s_timed_region": wall,
                  "note": "value counts living actors only" + ("; alive_* are rank 0's, per_rank has every rank's: strips "
                          "at the edge of the grid drain towards the middle, and the tick is as long as the fullest rank's"
                          if world > 1 else ""),
                  "per_rank": per_rank},
    }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
