#!/usr/bin/env python
"""bench.py -- simulated person-hours / second of the transmission + progression hot path.

Workload (BASELINE.json configs[0], the configuration the metric is quoted on): The Hague baseline scenario --
553 667 persons, ~143 k locations, disease model of thehague.properties (distance model, alpha-distance parameters),
seed 111, 100 initial infections -- on a SYNTHETIC population of the documented schema (the reference's people /
location blobs are not shipped).  A step is one simulated day (one 24 h window) of the hot path: occupancy bucketing,
every infectPeople evaluation of the day, Philox draws, resolve, disease-phase state machine.  The stays of each day
come from the reference's activity schedule rows through the host-side schedule generator (the part of the medlabs
engine the Java host keeps); they are recorded once, untimed, and replayed, so the timed runs are bit-identical.

  value : whole-job person-hours / s with every day's stays already resident in HBM (heros_submit_visits_device)
  e2e   : the same through the C ABI with pinned HOST buffers: H2D of the day's stays, heros_step, D2H of the phase
          counters and the new infection events, all inside the timed region
  --impl reference : the CPU oracle (sequential event-by-event restatement of the reference) on the same city and
          days, on all host cores the way the reference's authors fill a server: one single-threaded replication
          (seed) per core, aggregate throughput; the single-replication figure is reported beside it

N > 1 (launched by torch.distributed.run, one rank per GPU): weak scaling.  Every rank owns one city tile of the
configured size (persons + locations, tiles side by side); --commute of the workers of a tile work in the next tile,
so every window exchanges the stays of those commuters (with their packed agent words) and sends the candidate
infections back -- written by the pack / export kernels straight into the peers' memory over NVLink (--exchange p2p,
the default; csrc/exchange.cu) or as two NCCL all-to-alls per window (--exchange nccl; heros_b200.sharded).  Timing:
CUDA events on the engine stream, max over ranks.

usage: python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HAGUE_PERSONS = 553667
SEED = 111
N_INFECTED = 100
CITY_SEED = 20240807


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--persons", type=int, default=HAGUE_PERSONS, help="persons per city tile (per GPU)")
    ap.add_argument("--preroll", type=int, default=7, help="untimed simulated days before the warm-up steps")
    ap.add_argument("--model", default="distance", choices=["area", "distance"])
    ap.add_argument("--trigger", default="exit", choices=["exit", "both"])
    ap.add_argument("--cpu-days", type=int, default=14, help="days of the bounded cpu_baseline sample (about 12 s of CPU work)")
    ap.add_argument("--ref-cores", type=int, default=0,
                    help="--impl reference: replications run side by side (0 = one per available host core, at most 64)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: kernels write into the peers' memory over NVLink (CUDA IPC), or NCCL all-to-alls")
    ap.add_argument("--commute", type=float, default=0.05,
                    help="N > 1: fraction of the workers whose workplace is drawn from the whole city (the others work at one "
                         "of the 64 workplaces nearest to their home); BASELINE configs[4] names 0.01 / 0.05 / 0.20")
    ap.add_argument("--parity-persons", type=int, default=6000,
                    help="N > 1: persons per GPU of the small partitioned city that is run on all GPUs first and compared, "
                         "event for event, with the CPU oracle on the whole city (0 = skip)")
    ap.add_argument("--parity-days", type=int, default=8)
    ap.add_argument("--config", default="hague", choices=["hague", "replicas", "europe"],
                    help="hague: BASELINE configs[0] / [4] (the default the driver runs); replicas: configs[2], seeds side by side "
                         "on every GPU (--replicas engines per GPU, --persons defaults to 2 x The Hague, --steps 180 for the "
                         "configuration as written); europe: configs[3], one city of its own size per GPU linked by commuters")
    ap.add_argument("--replicas", type=int, default=4, help="--config replicas: engines (seeds) per GPU")
    return ap.parse_args()


def load_props(model):
    name = "params_alpha_area.json" if model == "area" else "params_alpha_distance.json"
    with open(os.path.join(ROOT, "tests", "golden", name)) as f:
        p = json.load(f)["properties"]
    p["generic.diseasePropertiesModel"] = model
    return p


class ClockSampler:
    """SM clock and throttle reasons of one GPU DURING the timed region (B200_PROFILING.md): sampled in-process through
    NVML every 10 ms (a spawned nvidia-smi takes tens of milliseconds -- longer than the timed region -- and its driver
    calls disturb the run it is meant to observe); nvidia-smi is the fallback when NVML cannot be loaded."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.reasons, self.sm_max = [], set(), None
        self._stop = threading.Event()
        self._index = index
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._nvml = self._dev = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml, self._dev = pynvml, pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._dev, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if index < len(ids) and ids[index].isdigit():
                return int(ids[index])
        return index

    def _sample(self):
        try:
            if self._nvml is not None:
                n = self._nvml
                self.samples.append(float(n.nvmlDeviceGetClockInfo(self._dev, n.NVML_CLOCK_SM)))
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._dev)
                for name, bit in (("hw_slowdown", n.nvmlClocksThrottleReasonHwSlowdown),
                                  ("hw_thermal_slowdown", n.nvmlClocksThrottleReasonHwThermalSlowdown),
                                  ("sw_thermal_slowdown", n.nvmlClocksThrottleReasonSwThermalSlowdown),
                                  ("sw_power_cap", n.nvmlClocksThrottleReasonSwPowerCap)):
                    if r & bit:
                        self.reasons.add(name)
                return
            out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                  "-i", str(self._index)], capture_output=True, text=True, timeout=5).stdout
            f = [x.strip() for x in out.strip().split(",")]
            self.samples.append(float(f[0]))
            self.sm_max = float(f[1])
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._sample()
            self._stop.wait(0.01 if self._nvml is not None else 0.05)

    def __enter__(self):
        self._sample()
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=10)
        self._sample()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "source": "NVML, every 10 ms during the timed region" if self._nvml is not None else "nvidia-smi"}


def oracle_objects(args, props, city, seed=SEED):
    from oracle import binding as ob
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import oracle_props
    area, dist, prog = oracle_props(props)
    model = ob.MODEL_AREA if args.model == "area" else ob.MODEL_DIST
    trig = ob.TRIGGER_EXIT_ONLY if args.trigger == "exit" else ob.TRIGGER_ENTER_AND_EXIT
    sim = ob.Sim(model, trig, seed)
    if model == ob.MODEL_AREA:
        sim.set_area(area)
    else:
        sim.set_dist(dist)
    sim.set_prog(prog)
    sim.set_location_types(city.type_infect_sub, city.type_correction)
    sim.set_locations(city.loc_type, city.loc_nsub, city.loc_area)
    sim.set_persons(city.age, city.female)
    return sim


def keyed_seeds(city, seed, n):
    """The persons heros_seed_infections picks (smallest Philox hash), recomputed on the host for the oracle arm."""
    from heros_b200 import scenario
    x, y, _, _ = scenario.philox4x32_10(np.arange(city.n_persons), 0, 0, 0x200, seed, seed >> 32)
    h = (x << np.uint64(32)) | y
    order = np.lexsort((np.arange(city.n_persons), h))
    return np.sort(order[:n]).astype(np.int32)


def _replication_worker(w, args, props, city, total, barrier, times, infections, gen_times=None):
    """One replication (seed SEED + w) of the oracle in its own process; every day's stays are generated untimed, then
    all replications run their day side by side between two barriers."""
    from heros_b200 import scenario
    seed = SEED + w
    sim = oracle_objects(args, props, city, seed)
    seeds = keyed_seeds(city, seed, N_INFECTED)
    sim.expose(seeds, np.zeros(len(seeds)))
    gen = scenario.ScheduleGenerator(city, seed)
    for d in range(total):
        tg = time.perf_counter()
        st = gen.generate_day(sim.agent_state()["phase"])
        if gen_times is not None:
            gen_times[w * total + d] = time.perf_counter() - tg   # the activity engine's share (numpy stand-in for medlabs)
        barrier.wait(timeout=900)                                  # a replication that died breaks the barrier for all
        times[2 * (w * total + d)] = time.perf_counter()          # CLOCK_MONOTONIC: one clock for all processes
        sim.add_visits(st["person"], st["loc"], st["sub"], st["t_in"], st["t_out"])
        sim.run(24.0 * (d + 1))
        times[2 * (w * total + d) + 1] = time.perf_counter()
        barrier.wait(timeout=900)                                  # a replication that died breaks the barrier for all
    infections[w] = int(sim.phase_counts()[1:].sum())


def host_cores():
    """Cores this process may use: the affinity mask, cut by the cgroup CPU quota of a container."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except (OSError, ValueError):
        pass
    return n


def host_memory_bytes():
    """Memory this process may still take: MemAvailable, cut by the cgroup limit of a container."""
    avail = None
    try:
        with open("/proc/meminfo") as f:
            for row in f:
                if row.startswith("MemAvailable:"):
                    avail = int(row.split()[1]) * 1024
    except OSError:
        pass
    try:
        with open("/sys/fs/cgroup/memory.max") as f:
            limit = f.read().strip()
        if limit != "max":
            used = 0
            try:
                with open("/sys/fs/cgroup/memory.current") as f:
                    used = int(f.read())
            except (OSError, ValueError):
                pass
            avail = min(avail, int(limit) - used) if avail is not None else int(limit) - used
    except (OSError, ValueError):
        pass
    return avail


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (the sequential oracle; the JVM reference is
    not runnable here: no JDK, medlabs / DSOL jars not vendored) on all the host threads it can use.  A replication is
    single-threaded by construction in the reference; its authors fill a server by starting one JVM per seed
    (src/main/resources/server/*.sh: seeds 111..120 side by side), so that is what this arm does: one oracle process
    per host core, seeds SEED, SEED+1, ..., the same city and days; the value is the aggregate over the replications
    (persons x 24 h x days x replications / wall time of the days, every day timed from the first start to the last
    finish).  cpu_baseline.one_replication holds the single-core figure of replication 0 beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from heros_b200 import scenario
    from oracle import binding as ob
    ob.lib()                                                      # built and loaded before the fork
    props = load_props(args.model)
    city = scenario.City(args.persons, seed=CITY_SEED, **scenario.HAGUE_CALIBRATED)
    scenario.capacity_diversion(city, SEED)
    workers = args.ref_cores if args.ref_cores > 0 else min(host_cores(), 64)
    mem = host_memory_bytes()
    if args.ref_cores <= 0 and mem is not None:                   # a replication holds ~1 KB per person beside the shared city
        workers = max(1, min(workers, int(0.5 * mem / (2048.0 * args.persons + 2.5e8))))
    total = args.preroll + args.warmup + args.steps
    ctx = mp.get_context("fork")                                  # the city tables are shared copy-on-write
    barrier = ctx.Barrier(workers)
    times = ctx.Array("d", 2 * workers * total, lock=False)       # start / end of every (replication, day)
    infected = ctx.Array("q", workers, lock=False)
    gen_times = ctx.Array("d", workers * total, lock=False)       # schedule generation of every (replication, day)
    procs = [ctx.Process(target=_replication_worker, args=(w, args, props, city, total, barrier, times, infected, gen_times), daemon=True)
             for w in range(workers)]
    for p in procs:
        p.start()
    for p in procs:
        p.join()
    if any(p.exitcode != 0 for p in procs):
        raise SystemExit("a replication of the reference arm failed: exit codes " + str([p.exitcode for p in procs]))
    t = np.asarray(times[:]).reshape(workers, total, 2)
    first = args.preroll + args.warmup
    day_wall = t[:, first:, 1].max(axis=0) - t[:, first:, 0].min(axis=0)      # per timed day: first start .. last finish
    seconds = float(day_wall.sum())
    value = workers * city.n_persons * 24.0 * args.steps / seconds
    own = float((t[0, first:, 1] - t[0, first:, 0]).sum())
    one_replication = city.n_persons * 24.0 * args.steps / own
    infections = [int(x) for x in infected[:]]
    # the same days with every replication's schedule generation counted in (the GPU arm's e2e_device_schedule produces
    # the stays on the device inside its timed region): per day the slowest generation + the day's run
    g = np.asarray(gen_times[:]).reshape(workers, total)[:, first:]
    seconds_with_schedule = float((day_wall + g.max(axis=0)).sum())
    # the same workload description as the GPU arm's at this N (the sample that was timed is one tile of it)
    world = max(int(args.gpus), 1)
    line = {
        "impl": "reference", "metric": "simulated person-hours/sec", "value": value, "unit": "person-hours/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * seconds / args.steps,
        "sec_per_sim_day": seconds / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, city, world),
        "cpu_baseline": {"value": value, "unit": "person-hours/s", "cores": workers, "kind": "port",
                         "host_cores": host_cores(),
                         "with_schedule_generation": {
                             "value": workers * city.n_persons * 24.0 * args.steps / seconds_with_schedule, "unit": "person-hours/s",
                             "what": "the same days with the generation of every day's stays timed too (scenario.ScheduleGenerator, "
                                     "numpy, one per replication) -- what the GPU arm's e2e_device_schedule is to be compared with"},
                         "one_replication": {"value": one_replication, "unit": "person-hours/s", "cores": 1,
                                             "what": "replication 0 alone on its core while the others run beside it"},
                         "sample": f"{args.steps} simulated days (days {first}..{total - 1}) of "
                                   + (f"a city of {args.persons} persons (one GPU's share of the {world} x {args.persons} city)"
                                      if world > 1 else "the city") +
                                   f", {workers} replications (seeds {SEED}..{SEED + workers - 1}) side by side, one "
                                   "oracle/heros_oracle.c process per core as the reference's server scripts start one "
                                   "single-threaded JVM per seed; value = aggregate over the replications; "
                                   "ms_per_step = wall time of one simulated day of all of them"},
        "e2e": {"value": value, "unit": "person-hours/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "infections": infections[0], "infections_per_replication": infections,
    }
    emit(line)


def scenario_calibration():
    from heros_b200 import scenario
    return scenario.HAGUE_CALIBRATED


def workload_config(args, city, world, exchange="p2p"):
    """The workload both arms are quoted on, from the arguments alone (the two arms must print the same object)."""
    step = {"step": "one simulated day (24 h window)", "first_timed_day": args.preroll + args.warmup,
            "l2": "inputs larger than L2: every step reads a fresh day of stays (26 B per stay, 61 MB at Hague size, written "
                  "to HBM before the timed region) and streams 32-byte records of all of them through the bucket buffers "
                  "(L2: 126 MB)"}
    if world == 1:
        n_own = city.n_locations
        return {"workload": f"thehague-synthetic (calibrated: {json.dumps(scenario_calibration())}) N={city.n_persons} locations={n_own} "
                            f"sublocations={int(city.loc_nsub[:n_own].sum())} model={args.model} trigger={args.trigger} "
                            f"seed={SEED} infected0={N_INFECTED}", **step}
    if args.config == "europe":
        sizes = [max(20000, int(args.persons * EUROPE_SIZES[r % len(EUROPE_SIZES)])) for r in range(world)]
        return {"workload": f"europe-synthetic: {world} cities of {sizes} persons (calibrated Hague density), one per GPU, "
                            f"{args.commute:.3f} of every city's workers work in another city chosen in proportion to its size "
                            f"(all pairs exchange), model={args.model} trigger={args.trigger} seed={SEED} infected0={N_INFECTED} per city",
                **step}
    return {"workload": f"one synthetic city of {world} x {args.persons} persons (Hague density and location mix, calibrated: "
                        f"{json.dumps(scenario_calibration())}) cut into "
                        f"{world} geographic shards along the Morton curve of its locations (equal expected load, a person "
                        f"belongs to the shard of its home), model={args.model} trigger={args.trigger} seed={SEED} "
                        f"infected0={N_INFECTED} per shard", **step,
            "cross_shard": f"{args.commute:.3f} of the workers work anywhere in the city, the others at one of the 64 workplaces "
                           "nearest to their home; every stay in another shard's location travels to its owner and successful "
                           "draws travel back, every window, between all pairs of shards (no source mask), "
                           "exchange = " + ("kernels writing into the peers' memory over NVLink (CUDA IPC), flags instead of "
                                            "collectives" if exchange == "p2p" else "2 equal-split NCCL all-to-alls") +
                           ", no host synchronisation inside a window"}


def make_engine(args, props, city, device, flags=0):
    import heros_b200 as hb
    from heros_b200 import properties
    model = hb.MODEL_AREA if args.model == "area" else hb.MODEL_DISTANCE
    trig = hb.TRIGGER_EXIT_ONLY if args.trigger == "exit" else hb.TRIGGER_ENTER_AND_EXIT
    eng = hb.HerosEngine(model, SEED, trigger=trig, device=device, flags=flags)
    city.install(eng)
    if model == hb.MODEL_AREA:
        eng.set_transmission_area(properties.area_params(props))
    else:
        eng.set_transmission_distance(properties.dist_params(props))
    eng.set_progression(properties.prog_params(props))
    return eng


EUROPE_SIZES = (1.5, 0.4, 0.9, 0.25, 1.2, 0.6, 1.0, 0.3)  # persons of the cities relative to --persons


def europe_city(args, rank, world):
    """BASELINE configs[3]: the Europe data set -- several cities of their own, one per GPU, linked only by the commuters of
    role 15 ("worker country to city", data/europe).  Synthetic stand-in: `world` cities of different sizes (EUROPE_SIZES x
    --persons), every one with its own locations and schedule; the share --commute of the workers of a city works in ANOTHER
    city, chosen in proportion to its size (all pairs of cities exchange).  A commuter's workplace is a foreign location of
    its shard's city table (scenario.City.ext_global); after work it goes straight home.  Returns (this rank's city, first
    global person id of every rank + total, first global location id of every rank + total)."""
    from heros_b200 import scenario
    sizes = [max(20000, int(args.persons * EUROPE_SIZES[r % len(EUROPE_SIZES)])) for r in range(world)]
    cities = [scenario.City(sizes[r], seed=CITY_SEED + 100 + r, origin_m=(40000.0 * r, 0.0), **scenario.HAGUE_CALIBRATED)
              for r in range(world)]   # every rank builds all of them: it needs the workplace tables of the others
    person_bases = np.concatenate([[0], np.cumsum([c.n_persons for c in cities])]).astype(np.int64)
    loc_bases = np.concatenate([[0], np.cumsum([c.n_locations for c in cities])]).astype(np.int64)
    city = cities[rank]
    scenario.capacity_diversion(city, SEED)
    rng = np.random.default_rng(5000 + rank)
    workers = np.nonzero(city.role == scenario.ROLE_WORKER)[0]
    pick = workers[rng.random(len(workers)) < args.commute]
    others = [r for r in range(world) if r != rank]
    w = np.array([sizes[r] for r in others], np.float64)
    dest_city = np.array(others)[rng.choice(len(others), len(pick), p=w / w.sum())]
    ext_gid, ext_nsub, ext_type = [], [], []
    work_loc, work_sub = city.work_loc.copy(), city.work_sub.copy()
    for r in others:
        who = pick[dest_city == r]
        if not len(who):
            continue
        other = cities[r]
        wp = np.nonzero(other.loc_type == scenario.TYPE_ID["Workplace"])[0]
        wts = np.maximum(other.loc_nsub[wp], 1).astype(np.float64)
        dest = wp[rng.choice(len(wp), len(who), p=wts / wts.sum())]
        uniq, inv = np.unique(dest, return_inverse=True)
        work_loc[who] = (city.n_locations + len(ext_gid) + inv).astype(np.int32)
        work_sub[who] = (rng.integers(0, 1 << 30, len(who)) % np.maximum(other.loc_nsub[dest], 1)).astype(np.int16)
        ext_gid += (loc_bases[r] + uniq).tolist()
        ext_nsub += other.loc_nsub[uniq].tolist()
        ext_type += other.loc_type[uniq].tolist()
    n_ext = len(ext_gid)
    city.work_loc, city.work_sub = work_loc, work_sub
    city.ext_global = np.array(ext_gid, np.int64)
    # the tables of the schedule generator get a row per foreign workplace; nothing is near it: whoever leaves it goes home
    city.loc_type = np.concatenate([city.loc_type, np.array(ext_type, np.uint8)])
    city.loc_nsub = np.concatenate([city.loc_nsub, np.array(ext_nsub, np.int16)])
    city.loc_area = np.concatenate([city.loc_area, np.zeros(n_ext, np.float32)])
    city.loc_x = np.concatenate([city.loc_x, np.zeros(n_ext, np.float32)])
    city.loc_y = np.concatenate([city.loc_y, np.zeros(n_ext, np.float32)])
    city.nearest = np.concatenate([city.nearest, np.full((n_ext,) + city.nearest.shape[1:], scenario.NONE, np.int32)])
    city.cand = np.concatenate([city.cand, np.full((n_ext,) + city.cand.shape[1:], scenario.NONE, np.int32)])
    city.n_cand = np.concatenate([city.n_cand, np.zeros((n_ext,) + city.n_cand.shape[1:], np.int32)])
    if getattr(city, "divert", None) is not None:
        city.divert = np.concatenate([city.divert, np.zeros(n_ext)])
    city.person_gid = person_bases[rank] + np.arange(city.n_persons, dtype=np.int64)   # draws keyed by global ids
    city.loc_gid = np.concatenate([loc_bases[rank] + np.arange(city.n_locations, dtype=np.int64), city.ext_global])
    return city, person_bases, loc_bases


def open_exchange(sh, guest_capacity, candidate_capacity, exchange, rank):
    """The window exchange of a shard: peer-to-peer stores over NVLink (the shards map each other's receive regions by
    CUDA IPC), or -- if any rank cannot -- the NCCL block exchange on all of them.  Returns (exchange object, kind)."""
    import torch
    import torch.distributed as dist
    from heros_b200 import sharded
    bx = None
    if exchange == "p2p":
        try:
            bx = sharded.PeerExchange(sh, guest_capacity, candidate_capacity=candidate_capacity)
            bx.connect()  # every shard listens to every other shard (no source mask)
            ok = 1
        except Exception as e:  # noqa: BLE001
            print(f"[bench] rank {rank}: peer-to-peer exchange unavailable ({e}); falling back to NCCL", file=sys.stderr)
            bx, ok = None, 0
        flag = torch.tensor([ok], dtype=torch.int32, device=sh.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            exchange, bx = "nccl", None
    if bx is None:
        bx = sharded.BlockExchange(sh, guest_capacity, candidate_capacity=candidate_capacity)
    return bx, exchange


def parity_check(args, rank, world, local, dev):
    """N > 1: a small city cut into `world` geographic shards runs on all GPUs -- one process per GPU, every shard with its
    own schedule generator, the same exchange as the timed run, all shards exchanging with all shards -- and rank 0
    compares the gathered infection events, event for event, with the CPU oracle on the whole city."""
    import torch
    import torch.distributed as dist
    import heros_b200 as hb
    from heros_b200 import properties, scenario, sharded
    seed, days = SEED, args.parity_days
    props = load_props(args.model)
    # more contagious than the scenario, so that a few days of a small city hold a few hundred events
    props["covidT_area.contagiousness" if args.model == "area" else "covidT_dist.alpha"] = "4.0" if args.model == "area" else "1.5"
    whole0 = scenario.City(args.parity_persons * world, seed=CITY_SEED + 1, work_anywhere_fraction=max(args.commute, 0.2),
                           **scenario.HAGUE_CALIBRATED)
    scenario.capacity_diversion(whole0, seed)
    part = sharded.partition_city(whole0, world)
    city = part.shard_city(rank)
    seeds = keyed_seeds(whole0, seed, 40 * world)
    seeds = part.person_new[seeds].astype(np.int32)   # ids of the sharded numbering
    eng = make_engine(args, props, city, local, hb._lib.FLAG_EXACT_STATS)
    sh = sharded.ShardedEngine(eng, rank, world, part.person_bases, part.loc_bases, device=dev)
    mine = np.sort(seeds[(seeds >= part.person_bases[rank]) & (seeds < part.person_bases[rank + 1])])
    eng.expose(mine, np.zeros(len(mine)))
    gen = scenario.ScheduleGenerator(city, seed)
    probe = scenario.ScheduleGenerator(city, seed)
    _, _, off0, _ = sharded.split_stays(city, probe.generate_day(np.zeros(city.n_persons, np.uint8)), rank,
                                        part.person_bases, part.loc_bases)
    cap = torch.tensor([int(2 * max(int(np.diff(off0).max()), 1024))], dtype=torch.int64, device=dev)
    dist.all_reduce(cap, op=dist.ReduceOp.MAX)
    bx, kind = open_exchange(sh, int(cap.item()), 65536, args.exchange, rank)
    open_remote, n_remote = None, 0
    for d in range(days):
        st = gen.generate_day(eng.agent_state()["phase"])
        own, rem, off, open_remote = sharded.split_stays(city, st, rank, part.person_bases, part.loc_bases, open_remote)
        eng.submit_visits(own["person"], own["loc"], own["sub"], own["t_in"], own["t_out"])
        n_remote += len(rem["person"])
        remote = {k: torch.from_numpy(np.ascontiguousarray(rem[k])).to(dev) for k, _ in COLS}
        bx.step_window(24.0 * (d + 1), remote, off)
    inf = eng.infections()
    counts = torch.from_numpy(eng.phase_counts().astype(np.int64)).to(dev)
    dist.all_reduce(counts)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object({k: inf[k] for k in ("person", "loc", "sub", "time", "p")}, gathered, dst=0)
    remote_total = torch.tensor([n_remote], dtype=torch.int64, device=dev)
    dist.all_reduce(remote_total)
    result = None
    if rank == 0:
        whole = part.global_city()
        sim = oracle_objects(args, props, whole, seed)
        sim.expose(np.sort(seeds), np.zeros(len(seeds)))
        g = scenario.ScheduleGenerator(whole, seed)
        for d in range(days):
            st = g.generate_day(sim.agent_state()["phase"])
            sim.add_visits(st["person"], st["loc"], st["sub"], st["t_in"], st["t_out"])
            sim.run(24.0 * (d + 1))
        oi = sim.infections()
        got = {k: np.concatenate([x[k] for x in gathered]) for k in gathered[0]}
        o1 = np.lexsort((oi["person"], oi["time"]))
        o2 = np.lexsort((got["person"], got["time"]))
        same = len(o1) == len(o2) and all((np.asarray(oi[k])[o1] == np.asarray(got[k])[o2]).all() for k in ("person", "loc", "sub", "time"))
        p_ok = same and bool(np.allclose(np.asarray(got["p"])[o2], np.asarray(oi["p"])[o1], rtol=1e-12, atol=4.5e-16))
        result = {"n_gpus": world, "persons": int(whole.n_persons), "days": days, "exchange": kind,
                  "infection_events": int(len(o2)), "oracle_events": int(len(o1)),
                  "cross_shard_stays": int(remote_total.item()),
                  "infection_lists_identical": bool(same), "probabilities_within_1e-12": p_ok,
                  "phase_counts_identical": bool((counts.cpu().numpy() == sim.phase_counts()).all()),
                  "what": "one process per GPU, geographic shards of one small city, all shards exchanging with all shards, "
                          "against oracle/heros_oracle.c on the whole city"}
        if not (same and p_ok):
            print(f"[bench] PARITY FAILURE at N = {world}: {result}", file=sys.stderr)
    torch.cuda.synchronize()
    dist.barrier()
    return result


COLS = (("person", np.int32), ("loc", np.int32), ("sub", np.int16), ("t_in", np.float64), ("t_out", np.float64))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from heros_b200 import scenario, sharded

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the HEROS engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep NCCL's banner / debug lines out of stdout (one JSON line only), whatever NCCL_DEBUG the launcher set
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/heros_bench_nccl_%h_%p.log")
        dist.init_process_group("nccl", device_id=dev)

    props = load_props(args.model)
    n_commuters = 0
    parity = None
    if world > 1:
        # Before anything is timed: a small partitioned city on all GPUs through the same exchange, against the CPU oracle
        # on the whole city (rank 0), event for event.
        if args.parity_persons > 0:
            parity = parity_check(args, rank, world, local, dev)
        if args.config == "europe":
            city, person_bases, loc_bases = europe_city(args, rank, world)
            n_commuters = int((city.work_loc >= city.n_locations).sum())
    if world > 1 and args.config != "europe":
        # weak scaling: ONE city of world x persons (constant density), cut into world geographic shards; every rank
        # builds the same city from the same seed and keeps its shard
        side = float(np.sqrt(world))
        whole = scenario.City(args.persons * world, seed=CITY_SEED, box_m=(12000.0 * side, 9000.0 * side),
                              work_anywhere_fraction=args.commute, **scenario.HAGUE_CALIBRATED)
        scenario.capacity_diversion(whole, SEED)
        part = sharded.partition_city(whole, world)
        city = part.shard_city(rank)
        person_bases, loc_bases = part.person_bases.copy(), part.loc_bases.copy()
        n_commuters = int((city.work_loc >= city.n_locations).sum())
        del whole, part
    elif world == 1:
        city = scenario.City(args.persons, seed=CITY_SEED, **scenario.HAGUE_CALIBRATED)
        scenario.capacity_diversion(city, SEED)
        person_bases = np.array([0, city.n_persons])
        loc_bases = np.array([0, city.n_locations])
    total = args.preroll + args.warmup + args.steps
    first_timed = args.preroll + args.warmup

    guest_capacity = [4096]
    exchange = [args.exchange]
    engines = []  # at N > 1 the engines (and their streams, which NCCL has seen) live until the process group is gone

    def new_engine(flags=0):
        eng = make_engine(args, props, city, local, flags)
        bx = None
        if world > 1:
            sh = sharded.ShardedEngine(eng, rank, world, person_bases, loc_bases, device=dev)
            bx, exchange[0] = open_exchange(sh, guest_capacity[0], 65536, exchange[0], rank)
        eng.seed_infections(N_INFECTED)
        engines.append(eng)
        return eng, bx, torch.cuda.ExternalStream(eng.stream, device=dev)

    def release(eng):
        if world == 1:
            eng.close()

    def submit_day(eng, d, local_cols, device):
        eng.submit_visits_ptr(local_cols["person"].numel(), *(local_cols[k].data_ptr() for k, _ in COLS), device=device)

    def day_step(eng, bx, d, local_cols, remote_cols, offsets, device):
        submit_day(eng, d, local_cols, device)
        if bx is None:
            return eng.step(24.0 * (d + 1))
        # block exchange (peer-to-peer stores over NVLink, or two equal-split NCCL all-to-alls): everything of a window is
        # enqueued on the engine stream (pinned host columns of the commuters' stays are copied on that stream too)
        return bx.step_window(24.0 * (d + 1), remote_cols, offsets)

    def split_day(st, open_remote):
        """global ids; the stays in other shards' locations are the remote ones (kept while they are open)"""
        return sharded.split_stays(city, st, rank, person_bases, loc_bases, open_remote)

    def pin(cols):
        return {k: torch.from_numpy(np.ascontiguousarray(cols[k])).pin_memory() for k, _ in COLS}

    # ---- record pass (untimed): the medlabs activity engine's job, done once on the host ----
    if world > 1:
        # block capacity of the guest exchange (rows per destination shard and window): twice what a day of a healthy
        # population sends to its busiest destination, agreed by all ranks
        probe = scenario.ScheduleGenerator(city, SEED)
        _, _, off0, _ = split_day(probe.generate_day(np.zeros(city.n_persons, np.uint8)), None)
        cap = torch.tensor([int(2 * max(int(np.diff(off0).max()), 1024))], dtype=torch.int64, device=dev)
        dist.all_reduce(cap, op=dist.ReduceOp.MAX)
        guest_capacity[0] = int(cap.item())
        del probe
    eng, bx, _ = new_engine()
    gen = scenario.ScheduleGenerator(city, SEED)
    days, open_remote = [], None
    for d in range(total):
        st = gen.generate_day(eng.agent_state()["phase"])
        loc_cols, rem_cols, offsets, open_remote = split_day(st, open_remote)
        days.append((pin(loc_cols), pin(rem_cols), offsets))
        day_step(eng, bx, d, *days[-1], device=False)
    ref_counts = eng.phase_counts()
    ref_infections = eng.infections()
    release(eng)
    h2d_bytes = [sum(c[k].numel() * c[k].element_size() for c in days[d][:2] for k, _ in COLS) for d in range(total)]
    remote_rows = float(np.mean([days[d][1]["person"].numel() for d in range(first_timed, total)]))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    step_ms = []  # per-step device times of the last timed_days call (this rank)
    host_ms = {"submit_begin": 0.0, "enqueue_window": 0.0, "submit_end_copy_done": 0.0, "wait": 0.0, "read_results": 0.0}

    def timed_days(eng, bx, stream, cols_of_day, device, after_step=None, d0=None, d1=None):
        """K steps bracketed by barrier + synchronize, timed with CUDA events on the engine stream.  A single engine is
        driven the way the asynchronous C ABI is meant to be used: the window of day d is enqueued (heros_step_async), the
        stays of day d + 1 are handed over while it runs (heros_submit_visits copies on its own stream / takes resident
        device buffers in on the side stream), then the host waits for day d (heros_wait) and reads its results."""
        d0 = first_timed if d0 is None else d0
        d1 = total if d1 is None else d1
        stats = []
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = []  # one event per step on the engine stream: the time of every single day

        def mark():
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            marks.append(ev)

        barrier()
        t0 = time.perf_counter()
        start.record(stream)
        if bx is None:
            submit_day(eng, d0, cols_of_day(d0)[0], device)
            host = host_ms
            for k in host:
                host[k] = 0.0
            for d in range(d0, d1):
                ta = time.perf_counter()
                if d + 1 < d1 and not device:
                    # host buffers: the copy of day d + 1 starts before the window of day d is enqueued
                    c = cols_of_day(d + 1)[0]
                    eng.submit_visits_begin_ptr(c["person"].numel(), *(c[k].data_ptr() for k, _ in COLS))
                tb = time.perf_counter()
                eng.step_async(24.0 * (d + 1))
                mark()
                tc = time.perf_counter()
                if d + 1 < d1:
                    if device:
                        submit_day(eng, d + 1, cols_of_day(d + 1)[0], device)
                    else:
                        eng.submit_visits_end()
                td = time.perf_counter()
                if after_step:
                    # the host reads the day's results before it hands over the next day (e2e)
                    stats.append(eng.wait())
                    te = time.perf_counter()
                    after_step()
                    tf = time.perf_counter()
                    host["wait"] += 1e3 * (te - td); host["read_results"] += 1e3 * (tf - te)
                host["submit_begin"] += 1e3 * (tb - ta); host["enqueue_window"] += 1e3 * (tc - tb)
                host["submit_end_copy_done"] += 1e3 * (td - tc)
            if not after_step:
                # everything is resident in HBM: the whole job is enqueued (no window waits for the host: the fill
                # counts of the pools, the work lists and the candidates stay on the device), one wait at the end
                stats.append(eng.wait())
        elif isinstance(bx, sharded.PeerExchange) and not (after_step and os.environ.get("HEROS_BENCH_SERIAL")):
            # the same pattern with the peer-to-peer exchange: the window of day d (guests out, candidates back -- flags in
            # the peers' memory, no host in the loop) is enqueued, the shard's own stays of day d + 1 follow beside it
            submit_day(eng, d0, cols_of_day(d0)[0], device)
            host = host_ms
            for k in host:
                host[k] = 0.0
            diag = [] if (os.environ.get("HEROS_BENCH_DIAG") and after_step) else None
            for d in range(d0, d1):
                ta = time.perf_counter()
                remote = cols_of_day(d)[1]
                if not device:
                    # The (few) stays of this window in other shards' locations cross PCIe FIRST, on the engine stream: behind
                    # the 50 MB of tomorrow's own stays they would wait a millisecond in the copy engine's queue, the window
                    # with them -- and, through the flags, every peer's window too.
                    with bx.shard.on_stream():
                        remote = {k: v.to(dev, non_blocking=True) for k, v in remote.items()}
                # Measured at N = 2: with tomorrow's copy enqueued BEFORE the window (as in the single-engine loop) the window
                # only started when the copy had finished (heros_wait 0.58 ms instead of 0.16 ms); enqueued behind the window
                # it overlaps.  HEROS_BENCH_EARLY_COPY=1 restores the other order.
                early = os.environ.get("HEROS_BENCH_EARLY_COPY") is not None
                if d + 1 < d1 and not device and early:
                    c = cols_of_day(d + 1)[0]
                    eng.submit_visits_begin_ptr(c["person"].numel(), *(c[k].data_ptr() for k, _ in COLS))
                tb = time.perf_counter()
                bx.step_window(24.0 * (d + 1), remote, cols_of_day(d)[2], wait=False)
                if d + 1 < d1 and not device and not early:
                    c = cols_of_day(d + 1)[0]
                    eng.submit_visits_begin_ptr(c["person"].numel(), *(c[k].data_ptr() for k, _ in COLS))
                mark()
                tc = time.perf_counter()
                host["submit_begin"] += 1e3 * (tb - ta); host["enqueue_window"] += 1e3 * (tc - tb)
                if d + 1 < d1:
                    if device:
                        submit_day(eng, d + 1, cols_of_day(d + 1)[0], device)
                    else:
                        eng.submit_visits_end()
                td = time.perf_counter()
                host["submit_end_copy_done"] += 1e3 * (td - tc)
                if after_step:
                    stats.append(eng.wait())
                    te = time.perf_counter()
                    after_step()
                    host["wait"] += 1e3 * (te - td); host["read_results"] += 1e3 * (time.perf_counter() - te)
                    if diag is not None:  # CLOCK_MONOTONIC: one clock for all processes of the machine
                        diag.append([d, ta, tb, tc, td, te, time.perf_counter(), stats[-1]["bucket_scan_ms"], stats[-1]["resolve_ms"],
                                     stats[-1]["gpu_ms"], stats[-1]["progress_ms"]])
            if not after_step:
                stats.append(eng.wait())
            if diag:
                with open(f"{os.environ['HEROS_BENCH_DIAG']}_rank{rank}.json", "w") as f:
                    json.dump({"columns": ["day", "start", "remote+begin", "window enqueued", "copy done", "wait done", "read done",
                                           "bucket_scan_ms", "resolve_ms", "gpu_ms", "progress_ms"], "rows": diag}, f)
        else:
            for d in range(d0, d1):
                stats.append(day_step(eng, bx, d, *cols_of_day(d), device=device))
                if after_step:
                    after_step()
        end.record(stream)
        end.synchronize()
        barrier()
        wall = time.perf_counter() - t0
        step_ms.clear()
        prev = start
        for ev in marks:
            step_ms.append(prev.elapsed_time(ev))
            prev = ev
        return max_over_ranks(start.elapsed_time(end) * 1e-3), max_over_ranks(wall), stats

    # ---- value: stays resident in HBM ----
    import heros_b200 as hb
    eng, bx, stream = new_engine(hb._lib.FLAG_RESIDENT_INPUTS)  # the device columns below are complete before they are submitted
    dev_days = [({k: days[d][0][k].to(dev) for k, _ in COLS}, {k: days[d][1][k].to(dev) for k, _ in COLS}, days[d][2])
                for d in range(total)]
    torch.cuda.synchronize()
    for d in range(args.preroll):
        day_step(eng, bx, d, *dev_days[d], device=True)
    timed_days(eng, bx, stream, lambda d: dev_days[d], True, d0=args.preroll, d1=first_timed)  # warm-up, driven like the timed steps
    counts_before = eng.phase_counts()
    with ClockSampler(local) as clocks:
        seconds, wall_seconds, stats = timed_days(eng, bx, stream, lambda d: dev_days[d], True)
    value_step_ms = list(step_ms)
    counts_after = eng.phase_counts()
    assert (eng.phase_counts() == ref_counts).all(), "replay diverged from the recorded run"
    release(eng)
    del dev_days

    # ---- e2e: through the C ABI with pinned host buffers, copies inside the timed region ----
    eng, bx, stream = new_engine()
    for d in range(args.preroll):
        day_step(eng, bx, d, *days[d], device=False)
    box = {"n_inf": len(eng.infections(ordered=False)["person"]), "d2h": 0, "counts": None}

    def read_results():
        box["counts"] = eng.phase_counts()          # D2H: the DiseaseMonitor counters
        new = eng.infections(first=box["n_inf"], ordered=False)    # D2H: the day's infection events (as logged)
        box["n_inf"] += len(new["person"])
        box["d2h"] += 64 + len(new["person"]) * 26

    timed_days(eng, bx, stream, lambda d: days[d], False, read_results, d0=args.preroll, d1=first_timed)  # warm-up
    box["d2h"] = 0
    e2e_seconds, e2e_wall, e2e_stats = timed_days(eng, bx, stream, lambda d: days[d], False, read_results)
    e2e_host_ms = {k: v / args.steps for k, v in host_ms.items()}
    e2e_host_ms["device_stage_ms"] = {k: float(np.sum([s_[k] for s_ in e2e_stats])) / args.steps
                                      for k in ("progress_ms", "bucket_ms", "bucket_scan_ms", "transmit_small_ms",
                                                "transmit_big_ms", "resolve_ms", "gpu_ms")}
    assert (box["counts"] == ref_counts).all()
    e2e_d2h = int(box["d2h"] / args.steps)
    release(eng)

    # ---- the same days with the activity schedule advanced on the device (heros_advance_schedule): no stays cross
    #      PCIe, the host only reads the counters and the infection events of each day (single engine only) ----
    dev_sched = None
    if world == 1:
        eng, _, stream = new_engine()
        eng.set_schedule(scenario.compile_schedule(city), SEED)
        for d in range(first_timed):
            eng.advance_schedule(d)
            eng.step(24.0 * (d + 1))
        box = {"n_inf": len(eng.infections()["person"]), "d2h": 0, "counts": None}
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record(stream)
        for d in range(first_timed, total):
            eng.advance_schedule(d)
            eng.step(24.0 * (d + 1))
            read_results()
        end.record(stream)
        end.synchronize()
        assert (box["counts"] == ref_counts).all(), "the device-side schedule diverged from the host-side generator"
        ds_seconds = start.elapsed_time(end) * 1e-3
        dev_sched = {"value": city.n_persons * 24.0 * args.steps / ds_seconds, "unit": "person-hours/s",
                     "ms_per_step": 1e3 * ds_seconds / args.steps, "h2d_bytes_per_step": 0,
                     "d2h_bytes_per_step": int(box["d2h"] / args.steps),
                     "what": "heros_advance_schedule + heros_step + D2H of the phase counters and the day's infection "
                             "events; same stays and results as the host-generated days, bit for bit"}
        release(eng)

    n_total_persons = int(person_bases[-1])
    value = n_total_persons * 24.0 * args.steps / seconds
    e2e_value = n_total_persons * 24.0 * args.steps / e2e_seconds

    def shutdown():
        if world > 1:
            # leave in an orderly way: everything the engine streams have touched is released before the streams go, and
            # the interpreter is left without running destructors against a CUDA context that is being torn down
            torch.cuda.synchronize()
            dist.barrier()
            dist.destroy_process_group()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        shutdown()
        return

    # ---- roofline (device times from CUDA events on the engine's streams, summed over the timed steps) ----
    keys = ("progress_ms", "bucket_ms", "bucket_scan_ms", "bucket_scatter_ms", "transmit_small_ms",
            "transmit_big_ms", "resolve_ms", "sub32_done_ms", "gpu_ms")
    stage = {k: float(np.sum([s[k] for s in stats])) for k in keys}
    stage["hot_done_ms"] = float(np.sum([s["group_done_ms"][0] for s in stats]))        # fork -> end of k_hot_eval
    stage["hot_chain_small_done_ms"] = float(np.sum([s["group_done_ms"][1] for s in stats]))
    visits = float(np.sum([s["visits"] for s in stats]))
    big_stays = float(np.sum([s["big_stays"] for s in stats]))
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            peak, peak_source = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_source = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
    # Algorithmic bytes: transmission 32 B per stay (24 B stay + 8 B agent word): every
    # record is read once by k_transmit_thread or by a group kernel; bucketing 16 B per stay (key + index in, out) for
    # each of the two passes over all stays (tickets: k_submit / k_window_open, which also runs the state machine, 16 B
    # per agent; move: k_bucket_scatter).  The transmission kernels run side by side (group kernels on auxiliary
    # streams next to k_transmit_thread and the sub-warp kernels, then k_draw), so the stage is timed as a whole.
    agents = float(city.n_persons * args.steps)
    transmit_ms = stage["transmit_small_ms"] + stage["transmit_big_ms"]
    kernels = {"transmission (k_transmit_stream | k_hot_filter, k_hot_chain<0..2>, k_hot_eval | k_transmit_sub<8,32>)": (transmit_ms, 32.0 * visits),
               "k_bucket_scatter": (stage["bucket_scatter_ms"], 16.0 * visits),
               "k_window_open": (stage["progress_ms"], 16.0 * agents),
               "k_transmit_stream (the hot path running beside it)": (stage["transmit_small_ms"], 32.0 * (visits - big_stays))}
    dominant = max(kernels, key=lambda k: kernels[k][0])
    achieved = kernels[dominant][1] / (kernels[dominant][0] * 1e-3) / 1e9
    traffic, traffic_source = None, None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_full.json")), reverse=True):
        with open(path) as f:
            rows = json.load(f)
        names = ("k_transmit", "k_hot") if dominant.startswith("transmission") else (dominant.split(" ")[0],)
        per_kernel = {}
        for r in rows:
            if any(n in r["kernel"] for n in names):
                per_kernel.setdefault(r["kernel"], []).append(r["traffic_B"])
        if per_kernel:
            traffic = float(sum(np.mean(v) for v in per_kernel.values()))
            traffic_source = (f"profiles/{os.path.basename(path)}: ncu --set full, dram bytes read + written per launch, "
                              f"summed over {sorted(per_kernel)}")
            break
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_source,
                "algorithmic_bytes_per_launch": kernels[dominant][1] / args.steps, "peak_source": peak_source,
                "peak_nominal": 8000.0, "frac_nominal": achieved / 8000.0,   # the north star's 8 TB/s, beside the measured peak
                "note": "a Hague-sized window is 2.1 M stays (68 MB of records): the stage is bound by the latency of the "
                        "largest sub-locations (group kernels), not by HBM; profiles/ holds the same bench at 5 M persons",
                "per_kernel_gbs": {k: (v[1] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else None) for k, v in kernels.items()},
                "stage_ms_per_step": {k: v / args.steps for k, v in stage.items()}}

    line = {
        "metric": "simulated person-hours/sec", "value": value, "unit": "person-hours/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * seconds / args.steps,
        "sec_per_sim_day": seconds / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, city, world, exchange[0]),
        "clocks": clocks.summary(),
        "timing": "CUDA events on the engine stream around the K steps, max over ranks; wall clock of the same region: "
                  f"{1e3 * wall_seconds / args.steps:.4f} ms/step",
        "e2e": {"value": e2e_value, "unit": "person-hours/s",
                "h2d_bytes_per_step": int(np.mean(h2d_bytes[first_timed:])), "d2h_bytes_per_step": e2e_d2h,
                "ms_per_step": 1e3 * e2e_seconds / args.steps, "wall_ms_per_step": 1e3 * e2e_wall / args.steps,
                "host_ms_per_step": e2e_host_ms},
        "gpu_launches": int(np.sum([s["gpu_launches"] for s in stats])) + args.steps,
        "roofline": roofline,
        "ms_per_step_min_median_max": ([float(np.min(value_step_ms)), float(np.median(value_step_ms)), float(np.max(value_step_ms))]
                                       if value_step_ms else None),
        "epidemic": {"what": "share of this rank's persons infected (E, I(A), I(S), H, ICU) / ever infected at the start and "
                             "at the end of the timed days (no policy in force)",
                     "infected_share": [float(counts_before[1:6].sum() / city.n_persons), float(counts_after[1:6].sum() / city.n_persons)],
                     "ever_infected_share": [float(counts_before[1:].sum() / city.n_persons), float(counts_after[1:].sum() / city.n_persons)]},
        "infections_total": int(ref_counts[1:].sum()), "visits_per_step": visits / args.steps,
        "device_ms_per_step": stage["gpu_ms"] / args.steps,
    }
    if dev_sched is not None:
        line["e2e_device_schedule"] = dev_sched
    if world > 1:
        line["cross_shard_stays_per_step_rank0"] = remote_rows
        line["rank0_shard"] = {"persons": city.n_persons, "locations": city.n_locations,
                               "persons_working_in_another_shard": n_commuters}
        line["cross_shard_fraction_rank0"] = remote_rows / max(1.0, visits / args.steps)
        if parity is not None:
            line["parity"] = parity
    elif args.cpu_days > 0:
        # ---- cpu_baseline: the oracle on a bounded sample (first cpu_days days), checked against the GPU run ----
        cpu_days = min(args.cpu_days, total)
        sim = oracle_objects(args, props, city)
        seeds = keyed_seeds(city, SEED, N_INFECTED)
        sim.expose(seeds, np.zeros(len(seeds)))
        t_cpu = 0.0
        for d in range(cpu_days):
            t0 = time.perf_counter()
            sim.add_visits(*(days[d][0][k].numpy() for k, _ in COLS))
            sim.run(24.0 * (d + 1))
            t_cpu += time.perf_counter() - t0
        oi = sim.infections()
        keep = ref_infections["time"] < 24.0 * cpu_days
        order = np.lexsort((oi["person"], oi["time"]))
        parity = bool(len(order) == int(keep.sum()) and (oi["person"][order] == ref_infections["person"][keep]).all()
                      and (oi["time"][order] == ref_infections["time"][keep]).all())
        line["cpu_baseline"] = {"value": city.n_persons * 24.0 * cpu_days / t_cpu, "unit": "person-hours/s", "cores": 1,
                                "kind": "port",
                                "sample": f"first {cpu_days} simulated days of the same city and stays, "
                                          f"oracle/heros_oracle.c single-threaded ({t_cpu:.1f} s)",
                                "infection_lists_identical": parity}
    emit(line)
    shutdown()


def run_replicas(args):
    """BASELINE configs[2]: Haaglanden (about twice The Hague), 32-seed batched replications -- the reference's own practice
    of one JVM per seed side by side (src/main/resources/server/flip-normal-area-cap.sh) becomes `--replicas` engines per GPU,
    every engine with its own seed, its own stream and its own host thread (a handle is single-threaded like a DSOL
    replication; the C ABI releases the GIL), schedule advanced on the device.  The path does not shard: N GPUs = N x
    replicas independent engines, no data-path collective."""
    import threading
    import torch
    import torch.distributed as dist
    import heros_b200 as hb
    from heros_b200 import properties, scenario
    rank, world, local = (int(os.environ.get(k, "0")) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    world = max(world, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/heros_bench_nccl_%h_%p.log")
        dist.init_process_group("nccl", device_id=dev)
    persons = args.persons if args.persons != HAGUE_PERSONS else 2 * HAGUE_PERSONS
    props = load_props(args.model)
    city = scenario.City(persons, seed=CITY_SEED, box_m=(17000.0, 12700.0), **scenario.HAGUE_CALIBRATED)
    scenario.capacity_diversion(city, SEED)
    tables = scenario.compile_schedule(city)
    R = args.replicas
    total = args.preroll + args.warmup + args.steps
    first_timed = args.preroll + args.warmup

    engines = []
    for r in range(R):
        seed = SEED + rank * R + r
        model = hb.MODEL_AREA if args.model == "area" else hb.MODEL_DISTANCE
        eng = hb.HerosEngine(model, seed, device=local)
        city.install(eng)
        (eng.set_transmission_area(properties.area_params(props)) if model == hb.MODEL_AREA
         else eng.set_transmission_distance(properties.dist_params(props)))
        eng.set_progression(properties.prog_params(props))
        eng.seed_infections(N_INFECTED)
        eng.set_schedule(tables, seed)
        engines.append(eng)
    streams = [torch.cuda.ExternalStream(e.stream, device=dev) for e in engines]
    launches = [0] * R
    go = threading.Barrier(R)
    start = [torch.cuda.Event(enable_timing=True) for _ in engines]
    end = [torch.cuda.Event(enable_timing=True) for _ in engines]
    origin = torch.cuda.Event(enable_timing=True)

    def run(i):
        e = engines[i]
        for d in range(first_timed):
            e.advance_schedule(d)
            e.step(24.0 * (d + 1))
        go.wait()
        if i == 0:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()
            origin.record()
        go.wait()
        start[i].record(streams[i])
        for d in range(first_timed, total):
            e.advance_schedule(d)
            st = e.step(24.0 * (d + 1))          # the replication's host reads the day's counters (DiseaseMonitor)
            launches[i] += st["gpu_launches"] + 1
            e.phase_counts()
        end[i].record(streams[i])
        end[i].synchronize()

    with ClockSampler(local) as clocks:
        threads = [threading.Thread(target=run, args=(i,)) for i in range(R)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        torch.cuda.synchronize()
    seconds = (max(origin.elapsed_time(end[i]) for i in range(R)) - min(origin.elapsed_time(start[i]) for i in range(R))) * 1e-3
    if world > 1:
        t = torch.tensor([seconds], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        seconds = float(t.item())
    infected = [int(e.phase_counts()[1:].sum()) for e in engines]
    value = world * R * city.n_persons * 24.0 * args.steps / seconds
    line = {
        "metric": "simulated person-hours/sec", "value": value, "unit": "person-hours/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * seconds / args.steps,
        "sec_per_sim_day": seconds / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"haaglanden-synthetic (calibrated Hague density, {json.dumps(scenario.HAGUE_CALIBRATED)}) N={city.n_persons} "
                               f"locations={city.n_locations}, {R} replications per GPU (seeds {SEED}..{SEED + world * R - 1}), "
                               f"model={args.model}, activity schedule on the device, replicas only (no exchange)",
                   "step": "one simulated day of every replication", "first_timed_day": first_timed,
                   "l2": "every replication streams its own 4.7 M stays per day (150 MB of records) through HBM; "
                         f"{R} replications side by side exceed the 126 MB L2"},
        "clocks": clocks.summary(),
        "timing": "CUDA events on every engine stream, from the first start to the last end of the timed days, max over ranks",
        "e2e": {"value": value, "unit": "person-hours/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 64 * R,
                "ms_per_step": 1e3 * seconds / args.steps,
                "what": "the schedule is advanced on the device: per day and replication the host only reads the counters"},
        "gpu_launches": int(sum(launches)),
        "replicas_per_gpu": R, "ms_per_replica_day": 1e3 * seconds / args.steps / R,
        "infected_per_seed_rank0": infected,
        "roofline": None, "cpu_baseline": None,
    }
    for e in engines:
        e.close()
    if rank == 0:
        emit(line)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    # stdout carries ONE JSON line: while the bench runs, file descriptor 1 points to stderr, so that whatever a library
    # prints from native code (NCCL's version banner, for one) cannot end up next to it; print() below restores it
    sys.stdout.flush()
    global _STDOUT_FD
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "replicas":
        run_replicas(args)
    else:
        run_ours(args)


_STDOUT_FD = None


def emit(line: dict):
    """The one JSON line, on the real stdout."""
    sys.stdout.flush()
    if _STDOUT_FD is not None:
        os.dup2(_STDOUT_FD, 1)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
