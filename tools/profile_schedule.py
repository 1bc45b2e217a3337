"""Device-side schedule: a few days of the Hague-sized city with heros_advance_schedule (development aid / ncu target)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import heros_b200 as hb
from heros_b200 import properties, scenario
days = int(sys.argv[1]) if len(sys.argv) > 1 else 6
with open(os.path.join(ROOT, "tests", "golden", "params_alpha_distance.json")) as f:
    props = json.load(f)["properties"]
city = scenario.City(553667)
eng = hb.HerosEngine(hb.MODEL_DISTANCE, 111)
city.install(eng)
eng.set_transmission_distance(properties.dist_params(props))
eng.set_progression(properties.prog_params(props))
eng.seed_infections(100)
eng.set_schedule(scenario.compile_schedule(city), 111)
for d in range(days):
    t = time.perf_counter(); eng.advance_schedule(d); t1 = time.perf_counter(); s = eng.step(24.0 * (d + 1)); t2 = time.perf_counter()
    print(f"day {d}: advance {1e3*(t1-t):.3f} ms, step {1e3*(t2-t1):.3f} ms (gpu {s['gpu_ms']:.3f}), visits {s['visits']}")
