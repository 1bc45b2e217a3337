#!/usr/bin/env python
"""BASELINE.json configs[0] end to end: The Hague baseline scenario (synthetic population of the documented schema),
one seed, 60 simulated days, entirely on the GPU -- heros_advance_schedule produces each day's stays, heros_step runs
transmission + progression -- and the outputs a user of the reference looks at: the DiseaseMonitor compartment series
(0.5 h samples, golden-spreadsheet column layout) as CSV and the infection event list.

usage: python examples/run_thehague.py [--days 60] [--seed 111] [--model distance|area] [--persons 553667] [--out DIR]
                                       [--location-policy FILE.csv|NAME] [--disease-policy FILE.csv] [--helsinki]
                                       [--summary FILE.json]
(policy files as in data/thehague/policies/ of the reference)
"""
from __future__ import annotations

import argparse
import csv
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--days", type=int, default=60)
    ap.add_argument("--seed", type=int, default=111)
    ap.add_argument("--model", default="distance", choices=["area", "distance"])
    ap.add_argument("--persons", type=int, default=553667)
    ap.add_argument("--infected", type=int, default=100)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "thehague"))
    ap.add_argument("--location-policy", default="",
                    help="policies.LocationPolicyFile (ConstructHerosModel.java:1145): a .csv file, or the name of one of the "
                         "reference's files kept in tests/golden/location_policies.json (e.g. HardLockdown_10-40)")
    ap.add_argument("--helsinki", action="store_true",
                    help="BASELINE.json configs[1]: the real Helsinki location table (tests/golden/helsinki_locations.npz = "
                         "data/helsinki/locations.csv.gz) under a synthetic population instead of the synthetic Hague")
    ap.add_argument("--summary", default="", help="write a JSON summary (daily compartment counts, timings) here")
    ap.add_argument("--disease-policy", default="", help="policies.DiseasePolicyFile (ConstructHerosModel.java:1179)")
    args = ap.parse_args()

    import heros_b200 as hb
    from heros_b200 import scenario

    with open(os.path.join(ROOT, "tests", "golden", f"params_alpha_{args.model}.json")) as f:
        props = json.load(f)["properties"]
    props["generic.diseasePropertiesModel"] = args.model
    t = time.time()
    if args.helsinki:
        import numpy as np
        table = dict(np.load(os.path.join(ROOT, "tests", "golden", "helsinki_locations.npz")))
        city = scenario.City(args.persons, location_table=table)
    else:
        city = scenario.City(args.persons, **scenario.HAGUE_CALIBRATED)
        scenario.capacity_diversion(city, 111)
    model = hb.HerosModel(props, seed=args.seed)                       # parameter map + engine
    city.install(model.engine)
    progression, transmission = hb.construct_disease_model(model)       # ConstructHerosModel.java:136-144
    monitor = hb.DiseaseMonitor(model, progression, 0.5)                # ConstructHerosModel.java:145
    model.engine.seed_infections(args.infected)                          # ConstructHerosModel.java:1109-1128
    model.engine.set_schedule(scenario.compile_schedule(city), args.seed)
    if args.location_policy.endswith(".csv"):
        hb.schedule_location_policies(model, args.location_policy)
    elif args.location_policy:
        with open(os.path.join(ROOT, "tests", "golden", "location_policies.json")) as f:
            hb.schedule_location_policy_rows(model, json.load(f)["files"][args.location_policy]["rows"])
    if args.disease_policy:
        hb.schedule_disease_policies(model, args.disease_policy)
    print(f"city of {city.n_persons} persons, {city.n_locations} locations built in {time.time() - t:.1f} s")

    t = time.time()
    gpu_ms = 0.0
    daily = [model.engine.phase_counts().tolist()]
    for day in range(args.days):
        hb.apply_location_policies(model, 24.0 * day)
        model.engine.advance_schedule(day)
        stats = model.engine.step(24.0 * (day + 1))
        gpu_ms += stats["gpu_ms"]
        daily.append(model.engine.phase_counts().tolist())
        if day % 10 == 9 or day == args.days - 1:
            c = model.engine.phase_counts()
            print(f"day {day + 1:3d}: S {c[0]}  E {c[1]}  I(A) {c[2]}  I(S) {c[3]}  H {c[4]}  ICU {c[5]}  D {c[6]}  R {c[7]}")
    wall = time.time() - t
    print(f"{args.days} simulated days in {wall:.2f} s wall ({gpu_ms:.1f} ms of window kernels): "
          f"{city.n_persons * 24.0 * args.days / wall:.3g} person-hours/s")

    if args.summary:
        os.makedirs(os.path.dirname(os.path.abspath(args.summary)), exist_ok=True)
        with open(args.summary, "w") as f:
            json.dump(dict(city="helsinki locations + synthetic persons" if args.helsinki else "synthetic Hague",
                           persons=city.n_persons, locations=city.n_locations, model=args.model, seed=args.seed,
                           days=args.days, location_policy=args.location_policy, disease_policy=args.disease_policy,
                           wall_s=wall, window_kernel_ms=gpu_ms,
                           person_hours_per_s=city.n_persons * 24.0 * args.days / wall,
                           header=["S", "E", "Ia", "Is", "H", "ICU", "D", "R"], daily_counts=daily), f)
    os.makedirs(args.out, exist_ok=True)
    monitor.write_csv(os.path.join(args.out, f"compartments_seed{args.seed}.csv"))
    inf = model.engine.infections()
    with open(os.path.join(args.out, f"infections_seed{args.seed}.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["time_h", "person", "location", "sublocation", "p"])
        for row in zip(inf["time"], inf["person"], inf["loc"], inf["sub"], inf["p"]):
            w.writerow([repr(float(row[0])), int(row[1]), int(row[2]), int(row[3]), repr(float(row[4]))])
    print(f"wrote {args.out}/compartments_seed{args.seed}.csv ({2 * 24 * args.days + 1} rows) and "
          f"infections_seed{args.seed}.csv ({len(inf['time'])} events)")
    model.close()


if __name__ == "__main__":
    main()
