"""Replications side by side on one GPU (BASELINE.json configs[2]: "32-seed batched replications", the reference's own
practice of one JVM per seed, src/main/resources/server/*.sh): R engines (handles) with seeds 111 .. 111+R-1 on the same
city, every engine driven by its own host thread (the C ABI releases the GIL, a handle is single-threaded like a DSOL
replication), all of them with the device-side activity schedule so that nothing crosses PCIe.  A Hague-sized window
keeps the GPU busy for a fraction of its capacity only (the kernels are latency-bound), so independent replications
overlap.  Prints one JSON line per R with the aggregate person-hours/s (CUDA events of every engine stream, max over
engines, around the timed days).

usage: python tools/bench_replicas.py [--replicas 1,2,4,8] [--persons N] [--days K] [--preroll D]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--replicas", default="1,2,4,8")
    ap.add_argument("--persons", type=int, default=553667)
    ap.add_argument("--days", type=int, default=20)
    ap.add_argument("--preroll", type=int, default=10)
    ap.add_argument("--model", default="distance", choices=["area", "distance"])
    args = ap.parse_args()

    import torch
    import heros_b200 as hb
    from heros_b200 import properties, scenario

    name = "params_alpha_area.json" if args.model == "area" else "params_alpha_distance.json"
    with open(os.path.join(ROOT, "tests", "golden", name)) as f:
        props = json.load(f)["properties"]
    city = scenario.City(args.persons, **scenario.HAGUE_CALIBRATED)
    tables = scenario.compile_schedule(city)
    model = hb.MODEL_AREA if args.model == "area" else hb.MODEL_DISTANCE
    dev = torch.device("cuda", 0)

    def make(seed):
        eng = hb.HerosEngine(model, seed)
        city.install(eng)
        if model == hb.MODEL_AREA:
            eng.set_transmission_area(properties.area_params(props))
        else:
            eng.set_transmission_distance(properties.dist_params(props))
        eng.set_progression(properties.prog_params(props))
        eng.seed_infections(100)
        eng.set_schedule(tables, seed)
        return eng

    for R in [int(x) for x in args.replicas.split(",")]:
        engines = [make(111 + r) for r in range(R)]
        streams = [torch.cuda.ExternalStream(e.stream, device=dev) for e in engines]
        for e in engines:  # untimed days
            for d in range(args.preroll):
                e.advance_schedule(d)
                e.step(24.0 * (d + 1))
        torch.cuda.synchronize()
        start = [torch.cuda.Event(enable_timing=True) for _ in engines]
        end = [torch.cuda.Event(enable_timing=True) for _ in engines]
        go = threading.Barrier(R)

        def run(i):
            e = engines[i]
            go.wait()
            start[i].record(streams[i])
            for d in range(args.preroll, args.preroll + args.days):
                e.advance_schedule(d)
                e.step(24.0 * (d + 1))
                e.phase_counts()
            end[i].record(streams[i])
            end[i].synchronize()

        threads = [threading.Thread(target=run, args=(i,)) for i in range(R)]
        first = torch.cuda.Event(enable_timing=True)
        first.record()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        torch.cuda.synchronize()
        # the replications start together (barrier); the job ends when the last one ends
        seconds = max(first.elapsed_time(end[i]) for i in range(R)) * 1e-3 - min(first.elapsed_time(start[i]) for i in range(R)) * 1e-3
        infected = [int(e.phase_counts()[1:].sum()) for e in engines]
        print(json.dumps({"replicas": R, "persons": city.n_persons, "days": args.days,
                          "value": R * city.n_persons * 24.0 * args.days / seconds, "unit": "person-hours/s",
                          "ms_per_day_per_replica": 1e3 * seconds / args.days, "ms_per_replica_day": 1e3 * seconds / args.days / R,
                          "infected_per_seed": infected,
                          "what": "heros_advance_schedule + heros_step + counters per day, one host thread per engine"}),
              flush=True)
        for e in engines:
            e.close()


if __name__ == "__main__":
    main()
