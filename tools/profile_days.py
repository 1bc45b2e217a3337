"""Run the Hague-sized synthetic city for a few days and print per-stage device times (development aid)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import heros_b200 as hb  # noqa: E402
from heros_b200 import properties, scenario  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--persons", type=int, default=553667)
ap.add_argument("--days", type=int, default=6)
ap.add_argument("--model", default="area")
ap.add_argument("--trigger", type=int, default=0)
ap.add_argument("--seeds", type=int, default=100)
ap.add_argument("--contagiousness", type=float, default=None)
ap.add_argument("--clocks", action="store_true")
ap.add_argument("--flags", type=int, default=0)
args = ap.parse_args()

with open(os.path.join(ROOT, "tests", "golden", f"params_alpha_{'area' if args.model == 'area' else 'distance'}.json")) as f:
    props = json.load(f)["properties"]
if args.contagiousness is not None:
    props["covidT_area.contagiousness"] = str(args.contagiousness)
t = time.time()
city = scenario.City(args.persons, **scenario.HAGUE_CALIBRATED)
scenario.capacity_diversion(city, 111)
print(f"city: {city.n_persons} persons, {city.n_locations} locations, {int(city.loc_nsub.sum())} sub-locations, {time.time()-t:.1f}s")
model = hb.MODEL_AREA if args.model == "area" else hb.MODEL_DISTANCE
eng = hb.HerosEngine(model, 111, trigger=args.trigger, flags=(2 if args.clocks else 0) | args.flags)
city.install(eng)
if model == hb.MODEL_AREA:
    eng.set_transmission_area(properties.area_params(props))
else:
    eng.set_transmission_distance(properties.dist_params(props))
eng.set_progression(properties.prog_params(props))
eng.seed_infections(args.seeds)
gen = scenario.ScheduleGenerator(city, 111)
for d in range(args.days):
    t = time.time()
    phase = eng.agent_state()["phase"]
    st = gen.generate_day(phase)
    t_gen = time.time() - t
    t = time.time()
    eng.submit_visits(st["person"], st["loc"], st["sub"], st["t_in"], st["t_out"])
    t_sub = time.time() - t
    t = time.time()
    s = eng.step(24.0 * (d + 1))
    t_step = time.time() - t
    print(f"day {d}: stays {len(st['person'])} gen {t_gen:.2f}s submit {t_sub*1e3:.1f}ms step wall {t_step*1e3:.2f}ms gpu {s['gpu_ms']:.3f}ms | "
          f"progress {s['progress_ms']:.3f} bucket {s['bucket_ms']:.3f} small {s['transmit_small_ms']:.3f} big {s['transmit_big_ms']:.3f} "
          f"resolve {s['resolve_ms']:.3f} retire {s['retire_ms']:.3f} | visits {s['visits']} segs {s['sublocations']} big {s['big_sublocations']} "
          f"evals {s['evaluations']} draws {s['draws']} inf {s['infections']} launches {s['gpu_launches']}")
    if args.clocks and d >= args.days - 3:
        np.set_printoptions(linewidth=200, suppress=True)
        pc = eng.phase_clocks()
        print("  phase kcycles (rows = class 0..3):"); print((pc[:, :10] / 1e3).round(1))
        print("  slowest evaluation phase: n, n_ill, R, m_ev, key | sub-locations of the class"); print(pc[:, 10:16])
print("counts", eng.phase_counts())
