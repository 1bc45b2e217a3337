#!/usr/bin/env python
"""One-screen summary of a bench.py JSON line: ms per step (value / e2e / device schedule), stage times, roofline."""
import json
import sys

for path in sys.argv[1:]:
    with open(path) as f:
        d = json.loads(f.read().strip().splitlines()[-1])
    e = d.get("e2e", {})
    ds = d.get("e2e_device_schedule", {})
    r = d.get("roofline", {})
    print(f"{path}: n_gpus {d.get('n_gpus')} value {d['value']:.3e} ms/step {d['ms_per_step']:.4f}  e2e {e.get('ms_per_step', 0):.4f}  "
          f"dev-sched {ds.get('ms_per_step', 0):.4f}  roofline frac {r.get('frac', 0):.4f} ({r.get('achieved', 0):.0f} GB/s)")
    st = r.get("stage_ms_per_step", {})
    print("   " + "  ".join(f"{k.replace('_ms', '')} {v:.4f}" for k, v in st.items()))
    cb = d.get("cpu_baseline")
    if cb:
        print(f"   cpu_baseline {cb['value']:.3e} ({cb['cores']} cores, {cb['kind']}) identical={cb.get('infection_lists_identical')}")
