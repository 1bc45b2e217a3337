"""Extracts the reference-produced epidemic curve of seed 111 from /root/reference/"seed comparison.xlsx" (the only
numbers the reference publishes for this path) into tests/golden/seed_comparison_111.json: the daily compartment
counts, the hour-of-day histogram of new exposures (decrease of Susceptible between two 0.5 h samples, binned by the
time of day of the later sample), the first exposure / first infectious times and the half-hourly I(A) / I(S) counts
of hours 48..96 (the incubation of the 100 initial cases).  Run in the build container
(the reference tree does not travel to the GPU box); the JSON is committed."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from heros_b200.xlsx import read_xlsx  # noqa: E402

book = read_xlsx("/root/reference/seed comparison.xlsx")
out = {"source": "seed comparison.xlsx, sheets 111-114 (The Hague, N = 553 667, 120 days, 0.5 h samples)",
       "header": book["111"][0][:9], "seeds": {}}
for sheet in ("111", "112", "113", "114"):
    a = np.array([[float(x) for x in r[:9]] for r in book[sheet][1:] if r and r[0] not in (None, "")])
    t, s = a[:, 0], a[:, 1]
    new = -np.diff(s)
    hist = np.zeros(48)
    for h, n in zip(t[1:] % 24.0, new):
        hist[int(round(h * 2)) % 48] += n
    infectious = a[:, 3] + a[:, 4]
    out["seeds"][sheet] = {
        "daily_counts": a[::48, 1:].astype(int).tolist(),
        "new_exposures_by_half_hour_of_day": hist.astype(int).tolist(),
        "first_exposure_h": float(t[1:][new > 0][0]), "first_infectious_h": float(t[infectious > 0][0]),
        "all_seeds_infectious_by_h": float(t[a[:, 2] + 0 * t < 1e9][np.argmax(infectious >= 100)]),
        "peak_symptomatic": [float(t[np.argmax(a[:, 4])]), int(a[:, 4].max())], "final_recovered": int(a[-1, 8]),
        # the 100 persons exposed at t = 0 turn infectious between 60 h and 91.2 h; nobody exposed later can before
        # 54.5 + 60 h, so I(A) + I(S) over 48..96 h is 100 x the empirical CDF of the incubation period
        # first sample with somebody in Hospitalized / ICU / Dead / Recovered: lower ends of the sojourn-time supports
        "first_time_positive_h": {n: float(t[np.argmax(a[:, c] > 0)]) for n, c in
                                  (("Hospitalized", 5), ("ICU", 6), ("Dead", 7), ("Recovered", 8))},
        "infectious_48_96h": {"t0": 48.0, "dt": 0.5,
                              "asymptomatic": a[(t >= 48) & (t <= 96), 3].astype(int).tolist(),
                              "symptomatic": a[(t >= 48) & (t <= 96), 4].astype(int).tolist()}}
with open(os.path.join(ROOT, "tests", "golden", "seed_comparison.json"), "w") as f:
    json.dump(out, f, indent=1)
print({k: (v["first_exposure_h"], v["first_infectious_h"], v["peak_symptomatic"], v["final_recovered"])
       for k, v in out["seeds"].items()})


# ---- out-flip-lockdown-no-10-20-comparison.xlsx: the same seed (111) without policy and with the lockdown starting on
#      day 10 / day 20.  What it pins: runs of one seed coincide sample for sample until the policy takes effect, and
#      the policy of midnight of day d shows from the first morning movements of day d on. ----
book = read_xlsx("/root/reference/out-flip-lockdown-no-10-20-comparison.xlsx")
runs = {k: np.array([[float(x) for x in r[:9]] for r in v[1:] if r and r[0] not in (None, "")]) for k, v in book.items()}


def first_difference(a, b):
    n = min(len(a), len(b))
    rows = np.nonzero((a[:n, 1:] != b[:n, 1:]).any(axis=1))[0]
    return float(a[rows[0], 0]) if len(rows) else None


pol = {"source": "out-flip-lockdown-no-10-20-comparison.xlsx, sheets Nothing / Day 10 / Day 20 (The Hague, seed 111, 0.5 h samples)",
       "samples": {k: int(len(v)) for k, v in runs.items()},
       "first_differing_sample_h": {"Day 10 vs Nothing": first_difference(runs["Day 10"], runs["Nothing"]),
                                    "Day 20 vs Nothing": first_difference(runs["Day 20"], runs["Nothing"]),
                                    "Day 10 vs Day 20": first_difference(runs["Day 10"], runs["Day 20"])},
       "nothing_equals_seed_111_of_seed_comparison_until_h": first_difference(
           runs["Nothing"], np.array([[float(x) for x in r[:9]] for r in read_xlsx("/root/reference/seed comparison.xlsx")["111"][1:]
                                      if r and r[0] not in (None, "")])),
       "susceptible_at_day": {k: {str(d): int(v[d * 48, 1]) for d in (10, 20, 30, 60, 120) if d * 48 < len(v)} for k, v in runs.items()},
       "susceptible_daily": {k: v[::48, 1].astype(int).tolist() for k, v in runs.items()}}
with open(os.path.join(ROOT, "tests", "golden", "policy_comparison.json"), "w") as f:
    json.dump(pol, f, indent=1)
print(pol["first_differing_sample_h"], pol["nothing_equals_seed_111_of_seed_comparison_until_h"])
