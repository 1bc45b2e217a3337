#!/usr/bin/env python
"""When every kernel of a window runs (heros_debug_timeline): the Hague-sized synthetic city, device-resident stays.

  python tools/timeline.py [--persons N] [--model distance] [--preroll 10] [--days 6] [--mode resident|sched] [--sync]
Prints, per kernel, the start of its first thread block and the end of its last one in microseconds after the start of
the window's first kernel, averaged over the timed windows.  --sync waits for every window (the host-driven pattern);
without it the windows are enqueued back to back.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import heros_b200 as hb  # noqa: E402
from heros_b200 import properties, scenario  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--persons", type=int, default=553667)
ap.add_argument("--preroll", type=int, default=10)
ap.add_argument("--days", type=int, default=6)
ap.add_argument("--model", default="distance")
ap.add_argument("--mode", default="resident")
ap.add_argument("--sync", action="store_true")
ap.add_argument("--alpha", type=float, default=None, help="covidT_dist.alpha / covidT_area.contagiousness override")
args = ap.parse_args()

with open(os.path.join(ROOT, "tests", "golden", f"params_alpha_{'area' if args.model == 'area' else 'distance'}.json")) as f:
    props = json.load(f)["properties"]
if args.alpha is not None:
    props["covidT_area.contagiousness" if args.model == "area" else "covidT_dist.alpha"] = str(args.alpha)
city = scenario.City(args.persons, **scenario.HAGUE_CALIBRATED)
scenario.capacity_diversion(city, 111)
model = hb.MODEL_AREA if args.model == "area" else hb.MODEL_DISTANCE
eng = hb.HerosEngine(model, 111, flags=hb._lib.FLAG_RESIDENT_INPUTS if args.mode == "resident" else 0)
city.install(eng)
if model == hb.MODEL_AREA:
    eng.set_transmission_area(properties.area_params(props))
else:
    eng.set_transmission_distance(properties.dist_params(props))
eng.set_progression(properties.prog_params(props))
eng.seed_infections(100)
eng.set_schedule(scenario.compile_schedule(city), 111)
total = args.preroll + args.days
COLS = ("person", "loc", "sub", "t_in", "t_out")
dev_days = []
if args.mode == "resident":
    # the days to be timed as device-resident columns: generated on the host from the phases of a scheduled run
    ref = hb.HerosEngine(model, 111)
    city.install(ref)
    (ref.set_transmission_area(properties.area_params(props)) if model == hb.MODEL_AREA
     else ref.set_transmission_distance(properties.dist_params(props)))
    ref.set_progression(properties.prog_params(props))
    ref.seed_infections(100)
    gen = scenario.ScheduleGenerator(city, 111)
    for d in range(total):
        st = gen.generate_day(ref.agent_state()["phase"])
        ref.submit_visits(*(st[k] for k in COLS))
        ref.step(24.0 * (d + 1))
        dev_days.append({k: torch.from_numpy(np.ascontiguousarray(st[k])).cuda() for k in COLS})
    ref.close()
    torch.cuda.synchronize()


def hand_over(d):
    if args.mode == "resident":
        c = dev_days[d]
        eng.submit_visits_ptr(c["person"].numel(), *(c[k].data_ptr() for k in COLS), device=True)
    else:
        eng.advance_schedule(d)


for d in range(args.preroll):
    hand_over(d)
    eng.step(24.0 * (d + 1))
eng.timeline(True)
t0 = time.perf_counter()
hand_over(args.preroll)
gpu_ms = 0.0
for d in range(args.preroll, total):
    eng.step_async(24.0 * (d + 1))
    if d + 1 < total:
        hand_over(d + 1)
    if args.sync:
        gpu_ms += eng.wait()["gpu_ms"]
if not args.sync:
    gpu_ms = eng.wait()["gpu_ms"]
wall = time.perf_counter() - t0
rows, n = eng.timeline(False)
print(f"{args.persons} persons, {args.model} model, days {args.preroll}..{total - 1}, mode {args.mode}, "
      f"{'one wait per window' if args.sync else 'windows enqueued back to back'}: "
      f"{1e3 * wall / args.days:.3f} ms wall per day, {gpu_ms / args.days:.3f} ms on the engine stream; {n} windows")
print(f"{'kernel':24s} {'start us':>9s} {'end us':>9s} {'length':>8s}")
for name, (a, b) in sorted(rows.items(), key=lambda kv: kv[1][0]):
    print(f"{name:24s} {a:9.1f} {b:9.1f} {b - a:8.1f}")
print("phase counts", eng.phase_counts())
