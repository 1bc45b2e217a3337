"""Extract the disease parameter sets and location-type tables of the reference into small JSON fixtures.

Run in the build container (reads /root/reference, which does not exist on the GPU box):
    python tools/make_golden_params.py
Outputs tests/golden/params_alpha_area.json, params_alpha_distance.json, locationtypes_thehague.json,
policy_masking_distance_10_40.json.  These are parameter VALUES (data), not code.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from heros_b200 import properties  # noqa: E402


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for name in ("alpha-area", "alpha-distance"):
        p = properties.load_properties(f"{REF}/src/main/resources/{name}.properties")
        with open(os.path.join(out, f"params_{name.replace('-', '_')}.json"), "w") as f:
            json.dump({"source": f"src/main/resources/{name}.properties", "properties": p}, f, indent=1, sort_keys=True)
    with open(f"{REF}/data/thehague/locations/locationtypes.csv", newline="") as f:
        rows = list(csv.reader(f))
    with open(os.path.join(out, "locationtypes_thehague.json"), "w") as f:
        json.dump({"source": "data/thehague/locations/locationtypes.csv", "header": rows[0], "rows": rows[1:]}, f, indent=1)
    with open(f"{REF}/data/thehague/policies/MaskingDistance_10-40.csv", newline="") as f:
        rows = list(csv.reader(f))
    with open(os.path.join(out, "policy_masking_distance_10_40.json"), "w") as f:
        json.dump({"source": "data/thehague/policies/MaskingDistance_10-40.csv", "header": rows[0], "rows": rows[1:]}, f, indent=1)
    print("wrote fixtures to", out)


if __name__ == "__main__":
    main()
