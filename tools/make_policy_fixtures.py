"""Collects the rows of the reference's policy files the tests and tools use (data, not code) into
tests/golden/location_policies.json so that they travel to the GPU box: data/thehague/policies/
{HardLockdown_Realistic_10-40, HardLockdown_10-40, HardLockdown_20}.csv (location policies; the disease policy
MaskingDistance_10-40.csv is tests/golden/policy_masking_distance_10_40.json, tools/make_golden_params.py).
Run in the container that has /root/reference."""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/data/thehague/policies"

if __name__ == "__main__":
    out = {"source": "data/thehague/policies/<name>.csv of the reference", "files": {}}
    for name in ("HardLockdown_Realistic_10-40", "HardLockdown_10-40", "HardLockdown_20"):
        with open(os.path.join(SRC, name + ".csv"), newline="") as f:
            rows = [r for r in csv.reader(f) if r]
        out["files"][name] = {"header": rows[0], "rows": rows[1:]}
        print(name, len(rows) - 1, "policy rows")
    with open(os.path.join(ROOT, "tests", "golden", "location_policies.json"), "w") as f:
        json.dump(out, f, indent=1)
