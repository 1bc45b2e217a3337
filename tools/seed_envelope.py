#!/usr/bin/env python
"""Epidemic curves of many seeds next to the only numbers the reference publishes for this path: the compartment
curves of ``seed comparison.xlsx`` (seeds 111-114, The Hague, lockdown from day 20, 120 days;
tests/golden/seed_comparison.json) and the day-60 / day-120 rows of the no-policy and lockdown runs quoted in SURVEY.md
section 6.  Every seed runs entirely on the GPU (device-side schedule, LocationPolicy file through
heros_set_closure_policy), like src/main/resources/flip-lockdown-day20-{area,dist}.properties: generic.RunLength = 120,
policies.NumberInfected = 100, policies.LocationPolicyFile = policies/HardLockdown_20.csv.

The population is the synthetic Hague of scenario.City (the real people / location blobs are not in the reference
tree), so this is a comparison of shapes and orders of magnitude, not a parity test; it is what a calibrated
population would be judged with (32-seed envelope against the reference's seeds).

usage: python tools/seed_envelope.py [--seeds 32] [--days 120] [--model area|distance] [--policy FILE.csv|NAME|none]
                                     [--persons 553667] [--out gpurun_out/seed_envelope.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def summary(daily):
    """daily: [days + 1, 8] compartment counts at midnight (S, E, Ia, Is, H, ICU, D, R)"""
    daily = np.asarray(daily, np.int64)
    return dict(peak_symptomatic=int(daily[:, 3].max()), peak_symptomatic_day=int(daily[:, 3].argmax()),
                final_recovered=int(daily[-1, 7]), final_dead=int(daily[-1, 6]), final_susceptible=int(daily[-1, 0]))


def run(args):
    """args: seeds, first_seed, days, model, policy, persons.  Returns the result object main() writes."""
    return _run(args)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=32)
    ap.add_argument("--first-seed", type=int, default=111)
    ap.add_argument("--days", type=int, default=120)
    ap.add_argument("--model", default="area", choices=["area", "distance"])
    ap.add_argument("--policy", default="HardLockdown_20",
                    help="a .csv file (policies.LocationPolicyFile), a name in tests/golden/location_policies.json, or none")
    ap.add_argument("--persons", type=int, default=553667)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "seed_envelope.json"))
    args = ap.parse_args()
    out = _run(args)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f)
    print(json.dumps({k: out[k] for k in ("model", "policy", "wall_s", "person_hours_per_s", "ours_peak_symptomatic",
                                          "reference_peak_symptomatic", "ours_final_recovered",
                                          "reference_final_recovered", "reference_days_inside_envelope",
                                          "reference_days_inside_envelope_day_15_to_60")}))


def _run(args):
    import heros_b200 as hb
    from heros_b200 import scenario

    with open(os.path.join(ROOT, "tests", "golden", f"params_alpha_{args.model}.json")) as f:
        props = json.load(f)["properties"]
    props["generic.diseasePropertiesModel"] = args.model
    with open(os.path.join(ROOT, "tests", "golden", "seed_comparison.json")) as f:
        golden = json.load(f)["seeds"]
    policy_rows = []
    if args.policy and args.policy != "none" and not args.policy.endswith(".csv"):
        with open(os.path.join(ROOT, "tests", "golden", "location_policies.json")) as f:
            policy_rows = json.load(f)["files"][args.policy]["rows"]
    t = time.time()
    city = scenario.City(args.persons, **scenario.HAGUE_CALIBRATED)
    scenario.capacity_diversion(city, 111)
    tables = scenario.compile_schedule(city)
    print(f"city of {city.n_persons} persons built in {time.time() - t:.1f} s", file=sys.stderr)

    runs, gpu_ms, wall = {}, 0.0, time.time()
    for seed in range(args.first_seed, args.first_seed + args.seeds):
        model = hb.HerosModel(props, seed=seed)
        city.install(model.engine)
        hb.construct_disease_model(model)
        model.engine.seed_infections(100)
        model.engine.set_schedule(tables, seed)
        if args.policy.endswith(".csv"):
            hb.schedule_location_policies(model, args.policy)
        elif args.policy and args.policy != "none":
            hb.schedule_location_policy_rows(model, policy_rows)
        daily = [model.engine.phase_counts().tolist()]
        for day in range(args.days):
            hb.apply_location_policies(model, 24.0 * day)
            model.engine.advance_schedule(day)
            gpu_ms += model.engine.step(24.0 * (day + 1))["gpu_ms"]
            daily.append(model.engine.phase_counts().tolist())
        runs[seed] = daily
        model.close()
        print(f"seed {seed}: {summary(daily)}", file=sys.stderr)
    wall = time.time() - wall

    curves = np.array([runs[s] for s in sorted(runs)], np.int64)          # [seed, day, phase]
    infected = curves[:, :, 1:6].sum(axis=2)
    env = dict(day=list(range(args.days + 1)),
               infected_min=infected.min(axis=0).tolist(), infected_median=np.median(infected, axis=0).tolist(),
               infected_max=infected.max(axis=0).tolist(),
               recovered_min=curves[:, :, 7].min(axis=0).tolist(), recovered_max=curves[:, :, 7].max(axis=0).tolist())
    ref = {}
    for s, g in golden.items():
        d = np.array(g["daily_counts"], np.int64)[: args.days + 1]
        ref[s] = dict(summary(d), infected=d[:, 1:6].sum(axis=1).tolist())
    ours = {str(s): summary(runs[s]) for s in sorted(runs)}
    days = min(args.days, 120)
    inside = [int(env["infected_min"][k] <= ref[s]["infected"][k] <= env["infected_max"][k]) for s in ref for k in range(days + 1)]
    bulk = [int(env["infected_min"][k] <= ref[s]["infected"][k] <= env["infected_max"][k]) for s in ref
            for k in range(15, min(days, 60) + 1)]
    out = dict(what="epidemic curves of %d seeds on the synthetic Hague next to the reference's seeds 111-114 "
                    "(seed comparison.xlsx: lockdown from day 20)" % args.seeds,
               model=args.model, policy=os.path.basename(args.policy), persons=city.n_persons, days=args.days,
               seeds=[args.first_seed, args.first_seed + args.seeds - 1],
               wall_s=wall, window_kernel_ms=gpu_ms,
               person_hours_per_s=args.seeds * city.n_persons * 24.0 * args.days / wall,
               ours=ours, envelope=env, reference=ref,
               ours_peak_symptomatic=[min(v["peak_symptomatic"] for v in ours.values()),
                                      max(v["peak_symptomatic"] for v in ours.values())],
               reference_peak_symptomatic=[min(v["peak_symptomatic"] for v in ref.values()),
                                           max(v["peak_symptomatic"] for v in ref.values())],
               ours_final_recovered=[min(v["final_recovered"] for v in ours.values()),
                                     max(v["final_recovered"] for v in ours.values())],
               reference_final_recovered=[min(v["final_recovered"] for v in ref.values()),
                                          max(v["final_recovered"] for v in ref.values())],
               reference_days_inside_envelope=float(np.mean(inside)),
               reference_days_inside_envelope_day_15_to_60=float(np.mean(bulk)) if bulk else None,
               city_calibration=dict(scenario.HAGUE_CALIBRATED),
               caveat="synthetic population (the reference's people / location files are missing from its tree) with three "
                      "free parameters fitted to these very curves: a statistical comparison, not a parity claim")
    return out


if __name__ == "__main__":
    main()
