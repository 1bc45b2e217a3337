"""Compile the reference's activity schedule workbook into the compact JSON the schedule generators read.

Run in the build container (reads /root/reference):  python tools/make_schedule_fixture.py
Source: data/thehague/activities/activityschedules_cap.xlsx, sheet "activityschedules", policy 0 rows only (policies
1 and 2 are never referenced: ConstructHerosModel.java:400).  A day pattern is cut after its first
``UntilFixedTimeActivity 24`` (10 of the 15 roles store the Susceptible block twice).
Row format: ConstructHerosModel.java:763-904 (columns A-Q).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from xlsx_reader import read_xlsx  # noqa: E402

# role id 1..15 -> week-pattern role name (ConstructHerosModel.java:565-743, HerosModel.java:313-330)
ROLES = ["infant", "kindergarten student", "primary school student", "secondary school student", "college student",
         "university student", "worker", "pensioner", "unemployed job-seeker", "weekend worker", "essential worker",
         "worker satellite to city", "worker city to satellite", "worker satellite to satellite",
         "worker country to city"]
PHASES = ["Susceptible", "Exposed", "Infected-Asymptomatic", "Infected-Symptomatic", "Hospitalized", "ICU", "Dead",
          "Recovered"]
KIND = {"StochasticDurationActivity": 0, "FixedDurationActivity": 0, "TravelActivityDistanceBased": 1,
        "TravelActivity": 1, "UntilFixedTimeActivity": 2}
LOCATOR = {"HomeLocator": 0, "WorkLocator": 1, "SchoolLocator": 1, "CurrentLocator": 2, "NearestLocator": 3,
           "NearestLocatorCap": 3, "NearestLocatorChoice": 3, "RandomLocator": 4, "RandomLocatorCap": 4,
           "DistanceBasedTravelLocator": 5}


def num(x, default=0.0):
    return default if x in (None, "") else float(x)


def main():
    src = "/root/reference/data/thehague/activities/activityschedules_cap.xlsx"
    rows = read_xlsx(src)["activityschedules"][1:]
    choices, choice_index = [], {}

    def choice_id(text):
        if not text:
            return -1
        items = []
        for part in text.split(";"):
            part = part.strip()
            if not part:
                continue
            if ":" in part:
                name, p = part.split(":")
                items.append([name.strip().replace("LocationType.", ""), float(p)])
            else:
                items.append([part.replace("LocationType.", ""), 1.0])
        key = json.dumps(items)
        if key not in choice_index:
            choice_index[key] = len(choices)
            choices.append(items)
        return choice_index[key]

    patterns, closed = {}, set()
    for r in rows:
        r = list(r) + [None] * (17 - len(r))
        if r[0] != "0":
            continue
        key = f"{r[1]}|{r[3]}|{r[2]}"
        if key in closed:
            continue
        kind = KIND[r[5]]
        dist = {"triangular": 2, "uniform": 1}.get((r[7] or "").strip(), 0)
        mode, mn, mx = num(r[8]), num(r[9]), num(r[10])
        if kind == 1:
            locator, max_distance, cid = LOCATOR[r[14]], num(r[15]), choice_id(r[16])
            if dist == 0:
                dist = 1  # travel time ~ U(min, max) hours (the reference derives it from the distance)
        else:
            locator, max_distance, cid = LOCATOR[r[12]], num(r[15]), choice_id(r[16])
        if dist == 0 and kind == 0:
            mn = mx = mode  # FixedDurationActivity
        act = {"name": (r[4] or "").strip(), "kind": kind, "locator": locator, "dist": dist, "min": mn, "mode": mode,
               "max": mx, "until": num(r[6]), "choice": cid, "max_distance": max_distance}
        patterns.setdefault(key, []).append(act)
        if kind == 2 and act["until"] >= 24:
            closed.add(key)
    out = {"source": "data/thehague/activities/activityschedules_cap.xlsx (policy 0)", "roles": ROLES, "phases": PHASES,
           "days": ["Monday", "Friday", "Saturday", "Sunday"], "choices": choices, "patterns": patterns}
    path = os.path.join(ROOT, "medlabs-heros_b200", "data", "activity_patterns_thehague.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("patterns", len(patterns), "choices", len(choices), "bytes", os.path.getsize(path))
    n_act = {k: len(v) for k, v in patterns.items()}
    print("max activities per day pattern", max(n_act.values()))
    missing = [f"{p}|{r}" for p in PHASES if p != "Dead" for r in ROLES if f"{p}|{r}|Monday" not in patterns]
    print("missing Monday patterns:", missing)


if __name__ == "__main__":
    main()
