"""Shared helpers of the test-suite: parameter fixtures, random scenarios, oracle / GPU runners, comparison."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import heros_b200 as hb  # noqa: E402
from heros_b200 import _lib, properties  # noqa: E402
from oracle import binding as ob  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_props(model: str) -> dict:
    """covidT_* / covidP.* values of the reference's alpha-area / alpha-distance.properties (tests/golden)."""
    name = "params_alpha_area.json" if model == "area" else "params_alpha_distance.json"
    with open(os.path.join(GOLDEN, name)) as f:
        p = json.load(f)["properties"]
    p["generic.diseasePropertiesModel"] = model
    return p


def thehague_location_types():
    with open(os.path.join(GOLDEN, "locationtypes_thehague.json")) as f:
        j = json.load(f)
    rows = j["rows"]
    names = [r[0] for r in rows]
    infect_sub = np.array([r[3].upper() == "TRUE" for r in rows], np.uint8)
    corr = np.array([float(r[5]) for r in rows], np.float64)
    return names, infect_sub, corr


def oracle_props(props: dict):
    """(AreaProps | None, DistProps | None, ProgProps) of the oracle from a property dict."""
    area = dist = None
    if any(k.startswith("covidT_area.") for k in props):
        area = ob.area_props({k: float(props["covidT_area." + k]) for k in properties.AREA_KEYS})
    if any(k.startswith("covidT_dist.") for k in props):
        dist = ob.dist_props({k: float(props["covidT_dist." + k]) for k in properties.DIST_KEYS})
    fractions, periods = properties.prog_tables(props)
    return area, dist, ob.prog_props(fractions, periods)


class Scenario:
    """A small synthetic city with a fixed list of stays (the visit stream the medlabs engine would produce)."""

    def __init__(self, seed: int, n_persons: int, n_households: int, n_workplaces: int, n_shops: int, days: int,
                 shop_visits_per_day: float = 1.0, tie_fraction: float = 0.1, short_gap_fraction: float = 0.05,
                 open_fraction: float = 0.02, n_parks: int = 0, park_nsub: int = 3):
        rng = np.random.default_rng(seed)
        self.days = days
        # location types: 0 Accommodation (sub-location = household), 1 Workplace, 2 Retail (single sub-location),
        # 3 (only with n_parks > 0) a type with infectInSublocation = false and several sub-locations: the
        # total-location branch (TArea:198-246 / TDist:189-246)
        self.type_infect_sub = np.array([1, 1, 1] + ([0] if n_parks else []), np.uint8)
        self.type_corr = np.array([1.0, 0.8, 1.25] + ([0.9] if n_parks else []), np.float64)
        n_home_loc = max(1, n_households // 4)
        home_nsub = np.full(n_home_loc, 0, np.int64)
        hh_loc = rng.integers(0, n_home_loc, n_households)
        hh_sub = np.zeros(n_households, np.int64)
        for h in range(n_households):
            hh_sub[h] = home_nsub[hh_loc[h]]
            home_nsub[hh_loc[h]] += 1
        home_nsub = np.maximum(home_nsub, 1)
        work_nsub = rng.integers(1, 6, n_workplaces)
        shop_nsub = np.ones(n_shops, np.int64)
        self.loc_type = np.concatenate([np.zeros(n_home_loc), np.ones(n_workplaces), np.full(n_shops, 2),
                                        np.full(n_parks, 3)]).astype(np.uint8)
        self.loc_nsub = np.concatenate([home_nsub, work_nsub, shop_nsub, np.full(n_parks, park_nsub)]).astype(np.int16)
        per_sub = np.concatenate([np.full(n_home_loc, 100.0), np.full(n_workplaces, 30.0),
                                  rng.uniform(50, 400, n_shops), np.full(n_parks, 60.0)])
        self.loc_area = (per_sub * self.loc_nsub).astype(np.float32)  # ConstructHerosModel.java:381
        self.n_loc = len(self.loc_type)
        w0, s0 = n_home_loc, n_home_loc + n_workplaces

        self.n_persons = n_persons
        self.age = rng.integers(0, 95, n_persons).astype(np.int8)
        self.female = (rng.random(n_persons) < 0.5).astype(np.uint8)
        hh = rng.integers(0, n_households, n_persons)
        work_loc = w0 + rng.integers(0, n_workplaces, n_persons)
        work_sub = (rng.integers(0, 1 << 30, n_persons) % self.loc_nsub[work_loc]).astype(np.int64)

        person, loc, sub, tin, tout = [], [], [], [], []

        def add(p, l, s, a, b):
            person.append(p); loc.append(l); sub.append(s); tin.append(a); tout.append(b)

        t_end = 24.0 * days
        for p in range(n_persons):
            t = 0.0
            here = (hh_loc[hh[p]], hh_sub[hh[p]])
            start = 0.0
            day = 0
            while True:
                # leave home in the morning
                leave = 24.0 * day + (8.0 if rng.random() < tie_fraction else rng.uniform(6.0, 9.5))
                if leave <= start:
                    leave = start + rng.uniform(0.1, 1.0)
                if leave >= t_end:
                    break
                add(p, here[0], here[1], start, leave)
                t = leave + rng.uniform(0.2, 0.4)  # travel: nowhere
                if self.age[p] >= 18 and self.age[p] < 67:
                    w_end = t + rng.uniform(3.0, 4.0)
                    add(p, work_loc[p], work_sub[p], t, w_end)
                    t = w_end
                    if rng.random() < 0.5:  # lunch at a shop, short hop
                        shop = s0 + rng.integers(0, n_shops)
                        gap = rng.uniform(0.001, 0.01) if rng.random() < short_gap_fraction else rng.uniform(0.05, 0.2)
                        add(p, shop, 0, t + gap, t + gap + rng.uniform(0.3, 0.8))
                        t = tout[-1] + rng.uniform(0.05, 0.2)
                    w_end = t + rng.uniform(3.0, 4.5)
                    add(p, work_loc[p], work_sub[p], t, w_end)
                    t = w_end + rng.uniform(0.2, 0.4)
                k = rng.poisson(shop_visits_per_day)
                for _ in range(k):
                    shop = s0 + rng.integers(0, n_shops)
                    zone = 0
                    if n_parks and rng.random() < 0.6:  # a walk in one of the zones of a park
                        shop = s0 + n_shops + rng.integers(0, n_parks)
                        zone = rng.integers(0, park_nsub)
                    d = rng.uniform(0.2, 1.0)
                    if rng.random() < short_gap_fraction:
                        d = rng.uniform(0.002, 0.02)
                    if n_parks and rng.random() < tie_fraction:
                        d = np.ceil((t + d) * 4.0) / 4.0 - t  # leave on the quarter hour: ties across the zones of a park
                    add(p, shop, zone, t, t + d)
                    t += d + rng.uniform(0.1, 0.3)
                if rng.random() < tie_fraction:
                    t = np.ceil(t)  # arrive home on the hour: ties across persons
                start = t
                day = int(start // 24.0) + 1 if start % 24.0 > 5.0 else int(start // 24.0)
                if start >= t_end:
                    break
            if start < t_end:
                # last stay at home, a few of them left open (still present when the run ends)
                add(p, here[0], here[1], start, np.inf if rng.random() < open_fraction else t_end + rng.uniform(1, 8))
        self.person = np.array(person, np.int32)
        self.loc = np.array(loc, np.int32)
        self.sub = np.array(sub, np.int16)
        self.t_in = np.array(tin, np.float64)
        self.t_out = np.array(tout, np.float64)
        order = rng.permutation(len(self.person))  # submission order is arbitrary
        for name in ("person", "loc", "sub", "t_in", "t_out"):
            setattr(self, name, getattr(self, name)[order])
        self.seeds = rng.choice(n_persons, size=max(2, n_persons // 20), replace=False).astype(np.int32)

    @property
    def n_visits(self):
        return len(self.person)


def run_oracle(scn: Scenario, props: dict, model: int, trigger: int, seed: int, days=None, policies=(),
               pre_infectious=None):
    area, dist, prog = oracle_props(props)
    sim = ob.Sim(model, trigger, seed)
    if model == ob.MODEL_AREA:
        sim.set_area(area)
    else:
        sim.set_dist(dist)
    sim.set_prog(prog)
    sim.set_location_types(scn.type_infect_sub, scn.type_corr)
    sim.set_locations(scn.loc_type, scn.loc_nsub, scn.loc_area)
    sim.set_persons(scn.age, scn.female)
    if getattr(scn, "role", None) is not None:
        sim.set_roles(scn.role)
    if getattr(scn, "rate_locations", None) is not None:  # (location ids, infection_rate_factor, infection_rate)
        sim.set_rate_locations(*scn.rate_locations)
    for (t, name, value) in policies:
        sim.set_parameter(name, value, t)
    sim.expose(scn.seeds, np.zeros(len(scn.seeds)))
    sim.add_visits(scn.person, scn.loc, scn.sub, scn.t_in, scn.t_out)
    days = scn.days if days is None else days
    counts = []
    for d in range(days):
        sim.run(24.0 * (d + 1))
        counts.append(sim.phase_counts())
    return dict(infections=sim.infections(), transitions=sim.transitions(), counts=np.array(counts),
                state=sim.agent_state(), evaluations=sim.n_evaluations(), draws=sim.n_draws(), sim=sim)


def make_engine(scn: Scenario, props: dict, model: int, trigger: int, seed: int, window_hours=24.0, policies=(),
                flags=_lib.FLAG_EXACT_STATS):
    eng = hb.HerosEngine(model, seed, trigger=trigger, window_hours=window_hours, flags=flags)
    eng.set_location_types(scn.type_infect_sub, scn.type_corr)
    eng.set_locations(scn.loc_type, scn.loc_nsub, scn.loc_area)
    eng.set_persons(scn.age, scn.female, getattr(scn, "role", None))
    if getattr(scn, "rate_locations", None) is not None:
        eng.set_rate_locations(*scn.rate_locations)
    if model == _lib.MODEL_AREA:
        eng.set_transmission_area(properties.area_params(props))
    else:
        eng.set_transmission_distance(properties.dist_params(props))
    eng.set_progression(properties.prog_params(props))
    for (t, name, value) in policies:
        eng.set_parameter(name, value, t)
    return eng


def run_gpu(scn: Scenario, props: dict, model: int, trigger: int, seed: int, days=None, window_hours=24.0,
            policies=(), submit="all", step_hours=24.0, flags=_lib.FLAG_EXACT_STATS):
    """submit: 'all' = every stay up-front; 'daily' = before each day only the stays that begin that day, stays still
    open at midnight first submitted as open (t_out = +inf) and re-submitted when they close."""
    eng = make_engine(scn, props, model, trigger, seed, window_hours, policies, flags)
    eng.expose(scn.seeds, np.zeros(len(scn.seeds)))
    days = scn.days if days is None else days
    stats = []
    counts = []
    if submit == "all":
        eng.submit_visits(scn.person, scn.loc, scn.sub, scn.t_in, scn.t_out)
    t = 0.0
    t_end = 24.0 * days
    while t < t_end:
        t1 = min(t + step_hours, t_end)
        if submit == "daily":
            begins = (scn.t_in >= t) & (scn.t_in < t1)
            closes_later = begins & (scn.t_out >= t1)
            t_out = np.where(closes_later, np.inf, scn.t_out)
            sel = np.nonzero(begins)[0]
            # stays opened in an earlier step that close in this one
            closing = np.nonzero((scn.t_in < t) & (scn.t_out >= t) & (scn.t_out < t1))[0]
            idx = np.concatenate([sel, closing])
            tout = np.concatenate([t_out[sel], scn.t_out[closing]])
            eng.submit_visits(scn.person[idx], scn.loc[idx], scn.sub[idx], scn.t_in[idx], tout)
        stats.append(eng.step(t1))
        if abs(t1 / 24.0 - round(t1 / 24.0)) < 1e-12:
            counts.append(eng.phase_counts())
        t = t1
    return dict(infections=eng.infections(), transitions=eng.transitions(), counts=np.array(counts),
                state=eng.agent_state(), stats=stats, engine=eng)


def sort_events(ev: dict, keys=("time", "person")):
    order = np.lexsort(tuple(ev[k] for k in reversed(keys)))
    return {k: v[order] for k, v in ev.items()}


P_RTOL = 1e-12   # north-star tolerance on infection probabilities (relative, fp64) ...
P_ATOL = 4.5e-16  # ... plus 4 ulp(1): p = 1 - exp(x) cancels, so |dp| >= ulp(1) whenever exp(x) differs in its last bit


def assert_same_run(gpu: dict, ora: dict):
    """Integer outputs identical, probabilities within the stated tolerance."""
    gi, oi = sort_events(gpu["infections"]), sort_events(ora["infections"])
    assert len(gi["person"]) == len(oi["person"]), (len(gi["person"]), len(oi["person"]))
    np.testing.assert_array_equal(gi["person"], oi["person"])
    np.testing.assert_array_equal(gi["loc"], oi["loc"])
    np.testing.assert_array_equal(gi["sub"], oi["sub"])
    np.testing.assert_array_equal(gi["time"], oi["time"])  # bit-exact event times
    np.testing.assert_allclose(gi["p"], oi["p"], rtol=P_RTOL, atol=P_ATOL)
    gt, ot = sort_events(gpu["transitions"]), sort_events(ora["transitions"])
    np.testing.assert_array_equal(gt["person"], ot["person"])
    np.testing.assert_array_equal(gt["phase"], ot["phase"])
    np.testing.assert_array_equal(gt["time"], ot["time"])
    np.testing.assert_array_equal(gpu["counts"], ora["counts"])
    np.testing.assert_array_equal(gpu["state"]["phase"], ora["state"]["phase"])
    np.testing.assert_array_equal(gpu["state"]["exposure"], ora["state"]["exposure"])
    ill = (ora["state"]["phase"] >= 1) & (ora["state"]["phase"] <= 5)
    np.testing.assert_array_equal(gpu["state"]["next_time"][ill], ora["state"]["next_time"][ill])
    np.testing.assert_array_equal(gpu["state"]["next_phase"][ill], ora["state"]["next_phase"][ill])


def occupancy_from_stays(st, first_sub, times):
    """numpy statement of heros_sample_occupancy: a stay counts at t iff t_in <= t < t_out"""
    pos = first_sub[st["loc"]] + st["sub"]
    per_sub = np.zeros((len(times), int(first_sub[-1])), np.int32)
    for j, t in enumerate(times):
        m = (st["t_in"] <= t) & (t < st["t_out"])
        np.add.at(per_sub[j], pos[m], 1)
    per_loc = np.add.reduceat(per_sub, first_sub[:-1], axis=1).astype(np.int32)
    return per_loc, per_sub


def policy_fixture(name: str):
    """(header, rows) of one of the reference's location policy files (tests/golden/location_policies.json, made by
    tools/make_policy_fixtures.py from data/thehague/policies/<name>.csv)."""
    with open(os.path.join(GOLDEN, "location_policies.json")) as f:
        entry = json.load(f)["files"][name]
    return entry["header"], entry["rows"]


def write_policy_csv(name: str, directory) -> str:
    """the fixture written back as the CSV file ``policies.LocationPolicyFile`` would name"""
    import csv
    header, rows = policy_fixture(name)
    path = os.path.join(str(directory), name + ".csv")
    with open(path, "w", newline="") as f:
        csv.writer(f).writerows([header] + rows)
    return path
