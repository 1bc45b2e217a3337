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
  --impl reference : the CPU oracle (sequential event-by-event restatement of the reference, one core) on the same
          city and days

usage: python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
       (N > 1: launched by torch.distributed.run, one rank per GPU; every rank simulates its own city tile)
"""
from __future__ import annotations

import argparse
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--persons", type=int, default=HAGUE_PERSONS)
    ap.add_argument("--preroll", type=int, default=7, help="untimed simulated days before the warm-up steps")
    ap.add_argument("--model", default="distance", choices=["area", "distance"])
    ap.add_argument("--trigger", default="exit", choices=["exit", "both"])
    ap.add_argument("--cpu-days", type=int, default=6, help="days of the bounded cpu_baseline sample")
    return ap.parse_args()


def load_props(model):
    name = "params_alpha_area.json" if model == "area" else "params_alpha_distance.json"
    with open(os.path.join(ROOT, "tests", "golden", name)) as f:
        p = json.load(f)["properties"]
    p["generic.diseasePropertiesModel"] = model
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.reasons, self.sm_max = [], set(), None
        self._stop = threading.Event()
        self._index = index
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _sample(self):
        try:
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
            self._stop.wait(0.1)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=10)
        self._sample()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def oracle_objects(args, props, city):
    from oracle import binding as ob
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import oracle_props
    area, dist, prog = oracle_props(props)
    model = ob.MODEL_AREA if args.model == "area" else ob.MODEL_DIST
    trig = ob.TRIGGER_EXIT_ONLY if args.trigger == "exit" else ob.TRIGGER_ENTER_AND_EXIT
    sim = ob.Sim(model, trig, SEED)
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


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (the sequential oracle; the JVM reference is
    not runnable here: no JDK, medlabs / DSOL jars not vendored).  One core: a replication is single-threaded by
    construction in the reference (server/*.sh start one JVM per seed)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from heros_b200 import scenario
    props = load_props(args.model)
    city = scenario.City(args.persons)
    sim = oracle_objects(args, props, city)
    seeds = keyed_seeds(city, SEED, N_INFECTED)
    sim.expose(seeds, np.zeros(len(seeds)))
    gen = scenario.ScheduleGenerator(city, SEED)
    total = args.preroll + args.warmup + args.steps
    secs = []
    for d in range(total):
        phase = sim.agent_state()["phase"]
        st = gen.generate_day(phase)
        t = time.perf_counter()
        sim.add_visits(st["person"], st["loc"], st["sub"], st["t_in"], st["t_out"])
        sim.run(24.0 * (d + 1))
        secs.append(time.perf_counter() - t)
    timed = secs[args.preroll + args.warmup:]
    seconds = float(np.sum(timed))
    value = city.n_persons * 24.0 * args.steps / seconds
    line = {
        "impl": "reference", "metric": "simulated person-hours/sec", "value": value, "unit": "person-hours/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * seconds / args.steps,
        "sec_per_sim_day": seconds / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, city),
        "cpu_baseline": {"value": value, "unit": "person-hours/s", "cores": 1, "kind": "port",
                         "sample": f"{args.steps} simulated days (days {args.preroll + args.warmup}.."
                                   f"{total - 1}) of the same city, oracle/heros_oracle.c single-threaded"},
        "e2e": {"value": value, "unit": "person-hours/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "infections": int(sim.phase_counts()[1:].sum()),
    }
    print(json.dumps(line))


def workload_config(args, city):
    return {"workload": f"thehague-synthetic N={city.n_persons} locations={city.n_locations} "
                        f"sublocations={int(city.loc_nsub.sum())} model={args.model} trigger={args.trigger} seed={SEED} "
                        f"infected0={N_INFECTED}",
            "step": "one simulated day (24 h window)", "first_timed_day": args.preroll + args.warmup,
            "l2": "inputs larger than L2: each step reads that day's fresh stay buffers and rewrites the bucketed copies "
                  "(~200 MB per step)"}


def make_engine(args, props, city, device):
    import heros_b200 as hb
    from heros_b200 import properties
    model = hb.MODEL_AREA if args.model == "area" else hb.MODEL_DISTANCE
    trig = hb.TRIGGER_EXIT_ONLY if args.trigger == "exit" else hb.TRIGGER_ENTER_AND_EXIT
    eng = hb.HerosEngine(model, SEED, trigger=trig, device=device)
    city.install(eng)
    if model == hb.MODEL_AREA:
        eng.set_transmission_area(properties.area_params(props))
    else:
        eng.set_transmission_distance(properties.dist_params(props))
    eng.set_progression(properties.prog_params(props))
    eng.seed_infections(N_INFECTED)
    return eng


COLS = (("person", np.int32), ("loc", np.int32), ("sub", np.int16), ("t_in", np.float64), ("t_out", np.float64))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import heros_b200 as hb
    from heros_b200 import scenario

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the HEROS engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    props = load_props(args.model)
    # weak scaling: every rank simulates its own city tile of the configured size
    city = scenario.City(args.persons, seed=20240807 + rank, origin_m=(12000.0 * rank, 0.0))
    total = args.preroll + args.warmup + args.steps
    first_timed = args.preroll + args.warmup

    # ---- record pass (untimed): the medlabs activity engine's job, done once on the host ----
    eng = make_engine(args, props, city, local)
    gen = scenario.ScheduleGenerator(city, SEED)
    days = []
    for d in range(total):
        st = gen.generate_day(eng.agent_state()["phase"])
        pinned = {k: torch.from_numpy(st[k]).pin_memory() for k, _ in COLS}
        days.append(pinned)
        eng.submit_visits_ptr(len(st["person"]), *(pinned[k].data_ptr() for k, _ in COLS), device=False)
        eng.step(24.0 * (d + 1))
    ref_counts = eng.phase_counts()
    ref_infections = eng.infections()
    eng.close()
    h2d_bytes = [sum(days[d][k].numel() * days[d][k].element_size() for k, _ in COLS) for d in range(total)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: stays resident in HBM ----
    eng = make_engine(args, props, city, local)
    dev_days = [{k: days[d][k].cuda() for k, _ in COLS} for d in range(total)]
    stats = []
    for d in range(first_timed):
        eng.submit_visits_ptr(dev_days[d]["person"].numel(), *(dev_days[d][k].data_ptr() for k, _ in COLS), device=True)
        eng.step(24.0 * (d + 1))
    barrier()
    with ClockSampler(local) as clocks:
        t0 = time.perf_counter()
        for d in range(first_timed, total):
            eng.submit_visits_ptr(dev_days[d]["person"].numel(), *(dev_days[d][k].data_ptr() for k, _ in COLS),
                                  device=True)
            stats.append(eng.step(24.0 * (d + 1)))
        barrier()
        seconds = time.perf_counter() - t0
    seconds = max_over_ranks(seconds)
    counts_value = eng.phase_counts()
    assert (counts_value == ref_counts).all(), "replay diverged from the recorded run"
    eng.close()
    del dev_days

    # ---- e2e: through the C ABI with pinned host buffers, copies inside the timed region ----
    eng = make_engine(args, props, city, local)
    for d in range(first_timed):
        eng.submit_visits_ptr(days[d]["person"].numel(), *(days[d][k].data_ptr() for k, _ in COLS), device=False)
        eng.step(24.0 * (d + 1))
    n_inf = len(eng.infections()["person"])
    d2h = 0
    barrier()
    t0 = time.perf_counter()
    for d in range(first_timed, total):
        eng.submit_visits_ptr(days[d]["person"].numel(), *(days[d][k].data_ptr() for k, _ in COLS), device=False)
        eng.step(24.0 * (d + 1))
        counts = eng.phase_counts()             # D2H: the DiseaseMonitor counters
        new = eng.infections(first=n_inf)       # D2H: the day's infection events
        n_inf += len(new["person"])
        d2h += 64 + len(new["person"]) * 26
    barrier()
    e2e_seconds = max_over_ranks(time.perf_counter() - t0)
    assert (counts == ref_counts).all()
    eng.close()

    n_total_persons = city.n_persons * world
    value = n_total_persons * 24.0 * args.steps / seconds
    e2e_value = n_total_persons * 24.0 * args.steps / e2e_seconds

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (device time from CUDA events on the engine stream) ----
    keys = ("progress_ms", "bucket_ms", "bucket_count_ms", "bucket_scan_ms", "bucket_scatter_ms", "transmit_small_ms",
            "transmit_big_ms", "resolve_ms", "retire_ms", "gpu_ms")
    stage = {k: float(np.sum([s[k] for s in stats])) for k in keys}
    visits = float(np.sum([s["visits"] for s in stats]))
    big_stays = float(np.sum([s["big_stays"] for s in stats]))
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            peak, peak_source = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_source = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
    # single kernels with their algorithmic bytes: k_transmit_thread streams every stay record once
    # (32 B / stay); k_bucket_scatter is the bucketing pass proper (16 B / stay: key + index in, out); the group kernels
    # handle the stays of the large sub-locations (32 B / stay)
    kernels = {"k_transmit_thread": (stage["transmit_small_ms"], 32.0 * visits),
               "k_bucket_scatter": (stage["bucket_scatter_ms"], 16.0 * visits),
               "k_bucket_count": (stage["bucket_count_ms"], 16.0 * visits),
               "k_transmit_group+k_transmit_warp (concurrent)": (stage["transmit_big_ms"], 32.0 * big_stays)}
    dominant = max(kernels, key=lambda k: kernels[k][0])
    achieved = kernels[dominant][1] / (kernels[dominant][0] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_source,
                "per_kernel_gbs": {k: (v[1] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else None) for k, v in kernels.items()},
                "stage_ms_per_step": {k: v / args.steps for k, v in stage.items()}}

    # ---- cpu_baseline: the oracle on a bounded sample (first cpu_days days), checked against the GPU run ----
    cpu_days = min(args.cpu_days, total)
    sim = oracle_objects(args, props, city)
    seeds = keyed_seeds(city, SEED, N_INFECTED)
    sim.expose(seeds, np.zeros(len(seeds)))
    t_cpu = 0.0
    for d in range(cpu_days):
        t0 = time.perf_counter()
        sim.add_visits(*(days[d][k].numpy() for k, _ in COLS))
        sim.run(24.0 * (d + 1))
        t_cpu += time.perf_counter() - t0
    oi = sim.infections()
    keep = ref_infections["time"] < 24.0 * cpu_days
    order = np.lexsort((oi["person"], oi["time"]))
    parity = bool(len(order) == int(keep.sum()) and (oi["person"][order] == ref_infections["person"][keep]).all()
                  and (oi["time"][order] == ref_infections["time"][keep]).all())
    cpu_value = city.n_persons * 24.0 * cpu_days / t_cpu

    line = {
        "metric": "simulated person-hours/sec", "value": value, "unit": "person-hours/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * seconds / args.steps,
        "sec_per_sim_day": seconds / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, city),
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": "person-hours/s",
                "h2d_bytes_per_step": int(np.mean(h2d_bytes[first_timed:])), "d2h_bytes_per_step": int(d2h / args.steps),
                "ms_per_step": 1e3 * e2e_seconds / args.steps},
        "gpu_launches": int(np.sum([s["gpu_launches"] for s in stats])) + 2 * args.steps,
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_value, "unit": "person-hours/s", "cores": 1, "kind": "port",
                         "sample": f"first {cpu_days} simulated days of the same city and stays, oracle/heros_oracle.c "
                                   f"single-threaded ({t_cpu:.1f} s)", "infection_lists_identical": parity},
        "infections_total": int(ref_counts[1:].sum()), "visits_per_step": visits / args.steps,
        "device_ms_per_step": stage["gpu_ms"] / args.steps,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
