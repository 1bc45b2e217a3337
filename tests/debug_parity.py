"""Development aid (kept under tests/: it runs the oracle, which only tests may do): generator-driven city on GPU vs
oracle, report the first differing infection events.  usage: python tests/debug_parity.py [--persons N] [--days D]"""
import argparse
import os
import sys

import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from heros_b200 import scenario  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--persons", type=int, default=60000)
ap.add_argument("--days", type=int, default=6)
ap.add_argument("--model", default="distance")
ap.add_argument("--trigger", default="exit")
ap.add_argument("--boost", type=float, default=1.0)
a = ap.parse_args()
a.preroll = 0; a.warmup = 0; a.steps = a.days
props = bench.load_props(a.model)
if a.boost != 1.0:
    k = "covidT_area.contagiousness" if a.model == "area" else "covidT_dist.alpha"
    props[k] = str(float(props[k]) * a.boost)
city = scenario.City(a.persons)
eng = bench.make_engine(a, props, city, 0)
gen = scenario.ScheduleGenerator(city, 111)
sim = bench.oracle_objects(a, props, city)
seeds = bench.keyed_seeds(city, 111, 100)
sim.expose(seeds, np.zeros(len(seeds)))
for d in range(a.days):
    ph = eng.agent_state()["phase"]
    st = gen.generate_day(ph)
    eng.submit_visits(st["person"], st["loc"], st["sub"], st["t_in"], st["t_out"])
    s = eng.step(24.0 * (d + 1))
    sim.add_visits(st["person"], st["loc"], st["sub"], st["t_in"], st["t_out"])
    sim.run(24.0 * (d + 1))
    gi = eng.infections(); oi = sim.infections()
    o = np.lexsort((oi["person"], oi["time"]))
    oi = {k: v[o] for k, v in oi.items()}
    same = len(gi["person"]) == len(oi["person"]) and (gi["person"] == oi["person"]).all() and (gi["time"] == oi["time"]).all()
    print(f"day {d}: gpu inf {len(gi['person'])} oracle inf {len(oi['person'])} same={same} evals gpu {s['evaluations']} oracle(total) {sim.n_evaluations()} phasecounts equal {(eng.phase_counts()==sim.phase_counts()).all()}")
    if not same:
        gs = set(zip(gi["person"].tolist(), gi["time"].tolist())); os_ = set(zip(oi["person"].tolist(), oi["time"].tolist()))
        only_g = sorted(gs - os_, key=lambda x: x[1])[:8]; only_o = sorted(os_ - gs, key=lambda x: x[1])[:8]
        print(" only gpu:", only_g); print(" only oracle:", only_o)
        for (p, t) in (only_g[:3] + only_o[:3]):
            src = gi if (p, t) in gs else oi
            i = np.nonzero((src["person"] == p) & (src["time"] == t))[0][0]
            loc, sub = int(src["loc"][i]), int(src["sub"][i])
            m = (st["loc"] == loc) & (st["sub"] == sub)
            print(f"  person {p} t {t!r} loc {loc} sub {sub} type {scenario.TYPE_NAMES[city.loc_type[loc]]} nsub {city.loc_nsub[loc]} p {src['p'][i]!r} stays today in seg {m.sum()} lastcalc gpu {eng.last_calculation(loc, sub)!r} oracle {sim.last_calc(loc, sub)!r}")
            mine = st["person"] == p
            print("   person stays:", list(zip(st["loc"][mine].tolist(), st["sub"][mine].tolist(), st["t_in"][mine].tolist(), st["t_out"][mine].tolist())))
        break
