"""Copies the rows of the reference's policy files the tests use (data, not code) into tests/golden/policies/ so that
they travel to the GPU box: data/thehague/policies/{HardLockdown_Realistic_10-40,HardLockdown_10-40,
HardLockdown_20,MaskingDistance_10-40}.csv.  Run in the container that has /root/reference."""
import csv
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/data/thehague/policies"
DST = os.path.join(ROOT, "tests", "golden", "policies")

if __name__ == "__main__":
    os.makedirs(DST, exist_ok=True)
    for name in ("HardLockdown_Realistic_10-40.csv", "HardLockdown_10-40.csv", "HardLockdown_20.csv",
                 "MaskingDistance_10-40.csv"):
        with open(os.path.join(SRC, name), newline="") as f:
            rows = [r for r in csv.reader(f) if r]
        with open(os.path.join(DST, name), "w", newline="") as f:
            csv.writer(f, lineterminator="\n").writerows(rows)
        print(name, len(rows) - 1, "policy rows")
