#!/usr/bin/env python
"""Calibration of the synthetic Hague against the reference's published curves (seed comparison.xlsx, seeds 111-114:
area model, lockdown from day 20): runs a few seeds of every candidate population entirely on the GPU and prints the
numbers the reference curves fix -- infected (E + I + H + ICU) on days 5 / 10 / 15 / 20 / 30 / 40 / 60, the symptomatic
peak and its day, recovered on day 120 -- next to the reference's ranges.

  python tools/calibrate_city.py --grid "household_size=2.15,2.4;workplace_scale=1,0.5,0.25;workplace_sub_mean=1.75,0.5"
  python tools/calibrate_city.py --set household_size=2.3 workplace_scale=0.3 --seeds 8 --days 120 --out profiles/x.json
"""
from __future__ import annotations

import argparse
import itertools
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DAYS_SHOWN = (5, 10, 15, 20, 30, 40, 60)


def run_city(hb, scenario, props, policy_rows, kwargs, seeds, days, persons):
    type_scale = {k[5:]: v for k, v in kwargs.items() if k.startswith("type_")}
    kw = {k: v for k, v in kwargs.items() if not k.startswith("type_")}
    city = scenario.City(persons, type_scale=type_scale, **kw)
    scenario.capacity_diversion(city, 111)
    tables = scenario.compile_schedule(city)
    out = []
    for seed in seeds:
        model = hb.HerosModel(props, seed=seed)
        city.install(model.engine)
        hb.construct_disease_model(model)
        model.engine.seed_infections(100)
        model.engine.set_schedule(tables, seed)
        hb.schedule_location_policy_rows(model, policy_rows)
        daily = [model.engine.phase_counts().tolist()]
        for day in range(days):
            hb.apply_location_policies(model, 24.0 * day)
            model.engine.advance_schedule(day)
            model.engine.step(24.0 * (day + 1))
            daily.append(model.engine.phase_counts().tolist())
        model.close()
        out.append(daily)
    return np.array(out, np.int64)


def describe(curves):
    inf = curves[:, :, 1:6].sum(axis=2)
    d = {f"d{k}": [int(inf[:, k].min()), int(inf[:, k].max())] for k in DAYS_SHOWN if k < inf.shape[1]}
    d["peak_Is"] = [int(curves[:, :, 3].max(axis=1).min()), int(curves[:, :, 3].max(axis=1).max())]
    d["peak_day"] = [int(curves[:, :, 3].argmax(axis=1).min()), int(curves[:, :, 3].argmax(axis=1).max())]
    d["R_end"] = [int(curves[:, -1, 7].min()), int(curves[:, -1, 7].max())]
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="")
    ap.add_argument("--set", nargs="*", default=[])
    ap.add_argument("--seeds", type=int, default=3)
    ap.add_argument("--days", type=int, default=60)
    ap.add_argument("--persons", type=int, default=553667)
    ap.add_argument("--model", default="area")
    ap.add_argument("--policy", default="HardLockdown_20")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import heros_b200 as hb
    from heros_b200 import scenario
    with open(os.path.join(ROOT, "tests", "golden", f"params_alpha_{args.model}.json")) as f:
        props = json.load(f)["properties"]
    props["generic.diseasePropertiesModel"] = args.model
    with open(os.path.join(ROOT, "tests", "golden", "location_policies.json")) as f:
        policy_rows = json.load(f)["files"][args.policy]["rows"] if args.policy != "none" else []
    with open(os.path.join(ROOT, "tests", "golden", "seed_comparison.json")) as f:
        golden = json.load(f)["seeds"]
    ref = np.array([g["daily_counts"] for g in golden.values()], np.int64)[:, : args.days + 1]
    print("reference", json.dumps(describe(ref)))
    combos = []
    if args.grid:
        axes = []
        for part in args.grid.split(";"):
            k, vals = part.split("=")
            axes.append([(k.strip(), float(v)) for v in vals.split(",")])
        combos = [dict(c) for c in itertools.product(*axes)]
    if args.set or not combos:
        combos.append({kv.split("=")[0]: float(kv.split("=")[1]) for kv in args.set})
    results = []
    for kw in combos:
        t = time.time()
        curves = run_city(hb, scenario, props, policy_rows, kw, range(111, 111 + args.seeds), args.days, args.persons)
        d = describe(curves)
        # distance to the reference: root mean square of the log ratio of the mean infected curves on the days shown plus
        # every tenth day, and of the recovered at the end
        days = sorted(set(DAYS_SHOWN) | set(range(10, args.days + 1, 10)))
        days = [k for k in days if k <= args.days]
        ours, theirs = curves[:, :, 1:6].sum(axis=2).mean(axis=0), ref[:, :, 1:6].sum(axis=2).mean(axis=0)
        terms = [np.log(max(ours[k], 1.0) / max(theirs[k], 1.0)) for k in days]
        terms.append(np.log(max(curves[:, -1, 7].mean(), 1.0) / max(ref[:, -1, 7].mean(), 1.0)))
        d["rms_log_ratio"] = float(np.sqrt(np.mean(np.square(terms))))
        results.append(dict(city=kw, summary=d, curves=curves.tolist() if args.out else None))
        print(json.dumps(kw), json.dumps(d), f"{time.time() - t:.0f}s", flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(dict(reference=describe(ref), model=args.model, policy=args.policy, persons=args.persons, runs=results), f)


if __name__ == "__main__":
    main()
