"""Random campaign of the brute-force statement of the driver contract (tests/test_oracle.py::brute_force_run) against
the C oracle: 400 random small cities (ties, sub-threshold gaps, parks with several zones) x both models x both triggers.
CPU only, about two minutes.  Last result (2026-10-20): 1600 runs, 174 954 infections compared, 0 mismatching runs."""
import sys, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_oracle import brute_force_run
from common import Scenario
rng = np.random.default_rng(20261020)
total = mismatches = runs = 0
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 400):
    kw = dict(seed=int(rng.integers(1, 10**6)), n_persons=int(rng.integers(80, 700)), n_households=int(rng.integers(30, 250)),
              n_workplaces=int(rng.integers(3, 25)), n_shops=int(rng.integers(1, 5)), days=4,
              tie_fraction=float(rng.uniform(0, 0.4)), short_gap_fraction=float(rng.uniform(0, 0.3)),
              n_parks=int(rng.integers(0, 3)), park_nsub=int(rng.integers(2, 5)))
    scn = Scenario(**kw)
    for model, trigger in (("area", "exit"), ("distance", "exit"), ("area", "both"), ("distance", "both")):
        events, got = brute_force_run(model, trigger, scn, seed=int(rng.integers(1, 2**31)))
        ok = (len(events) == len(got["person"]) and [e[1] for e in events] == got["person"].tolist()
              and [e[0] for e in events] == got["time"].tolist() and [e[2] for e in events] == got["loc"].tolist()
              and [e[3] for e in events] == got["sub"].tolist()
              and np.allclose([e[4] for e in events], got["p"], rtol=1e-12, atol=4.5e-16))
        runs += 1; total += len(events); mismatches += (not ok)
        if not ok: print("MISMATCH", kw, model, trigger, len(events), len(got["person"]))
print(f"{runs} runs, {total} infections compared, {mismatches} mismatching runs")
