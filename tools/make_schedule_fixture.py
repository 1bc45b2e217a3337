"""Compile the reference's activity schedule workbook into the compact JSON the schedule generators read.

Run in the build container (reads /root/reference):  python tools/make_schedule_fixture.py
Source: data/thehague/activities/activityschedules_cap.xlsx, sheet "activityschedules", policy 0 rows only; the reading
rules are loaders.read_activity_schedules (ConstructHerosModel.java:763-904)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from heros_b200 import loaders  # noqa: E402


def main():
    out = loaders.read_activity_schedules("/root/reference/data/thehague/activities/activityschedules_cap.xlsx")
    out["source"] = "data/thehague/activities/activityschedules_cap.xlsx (policy 0)"
    path = os.path.join(ROOT, "medlabs-heros_b200", "data", "activity_patterns_thehague.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    patterns = out["patterns"]
    print("patterns", len(patterns), "choices", len(out["choices"]), "bytes", os.path.getsize(path))
    print("max activities per day pattern", max(len(v) for v in patterns.values()))
    missing = [f"{p}|{r}" for p in loaders.PHASES if p != "Dead" for r in loaders.ROLES if f"{p}|{r}|Monday" not in patterns]
    print("missing Monday patterns:", missing)


if __name__ == "__main__":
    main()
