"""Rate-based locations (medlabs LocationProbBased, factory/ConstructHerosModel.java:274-295, 382-388; the rows of
data/thehague/epidemiology/infection_rates.csv).  The class is inside the un-vendored medlabs jar, so the rule is a stated
convention: a susceptible person leaving such a location at t after d hours is exposed iff
U(seed; person, t, TAG_RATE) < 1 - exp(-lambda d / 24), lambda = infection_rate per day, or infection_rate_factor x the
share of the reference group (role Worker) exposed the day before.  Here: the oracle against a plain-Python statement."""
import math
import os

import numpy as np

from oracle import binding as ob
from common import load_props, oracle_props

TAG_RATE = 0x500


def tiny_sim(seed, n_persons, roles):
    props = load_props("distance")
    _, dist, prog = oracle_props(props)
    sim = ob.Sim(ob.MODEL_DIST, ob.TRIGGER_EXIT_ONLY, seed)
    sim.set_dist(dist)
    sim.set_prog(prog)
    sim.set_location_types(np.array([1, 0], np.uint8), np.array([1.0, 1.0]))
    # location 0: an ordinary home with one sub-location; 1: a satellite with a fixed rate; 2: one that follows the reference group
    sim.set_locations(np.array([0, 1, 1], np.uint8), np.array([1, 1, 3], np.int16), np.array([100.0, 50.0, 50.0], np.float32))
    sim.set_persons(np.full(n_persons, 40, np.int8), np.zeros(n_persons, np.uint8))
    sim.set_roles(np.asarray(roles, np.uint8))
    return sim


def bits(t):
    return int(np.float64(t).view(np.uint64))


def test_fixed_rate_location_follows_the_stated_formula():
    seed, n = 99, 400
    rate = 0.6  # per day
    sim = tiny_sim(seed, n, np.full(n, 12))
    sim.set_rate_locations([1], [-1.0], [rate])
    rng = np.random.default_rng(1)
    t_in = 8.0 + rng.random(n) * 2.0
    dur = 1.0 + rng.random(n) * 9.0
    t_out = t_in + dur
    person = np.arange(n, dtype=np.int32)
    sim.add_visits(person, np.ones(n, np.int32), np.zeros(n, np.int16), t_in, t_out)
    sim.run(24.0)
    inf = sim.infections()
    want = []
    for i in range(n):
        p = 1.0 - math.exp(-(rate * ((t_out[i] - t_in[i]) / 24.0)))
        if ob.u01(seed, i, bits(t_out[i]), TAG_RATE) < p:
            want.append((i, t_out[i], p))
    assert 20 < len(want) < n
    order = np.argsort(inf["person"])
    assert inf["person"][order].tolist() == [w[0] for w in want]
    assert inf["time"][order].tolist() == [w[1] for w in want]
    np.testing.assert_allclose(inf["p"][order], [w[2] for w in want], rtol=1e-15)
    assert (inf["loc"] == 1).all() and sim.n_draws() == n and sim.n_evaluations() == 0  # no occupancy evaluation there
    # the infected are exposed at their exit time (float) and progress like everybody else
    st = sim.agent_state()
    assert (st["phase"][[w[0] for w in want]] == 1).all()
    np.testing.assert_array_equal(st["exposure"][[w[0] for w in want]], np.float32([w[1] for w in want]))


def test_factor_location_follows_the_reference_group_of_the_day_before():
    seed, n = 7, 600
    roles = np.where(np.arange(n) < 300, 7, 12)  # 300 Workers (the reference group), 300 satellite commuters
    factor = 0.9
    sim = tiny_sim(seed, n, roles)
    sim.set_rate_locations([2], [factor], [-1.0])
    # day 0: 60 Workers and 10 commuters are exposed by hand (seeding counts like any exposure)
    seeded = np.r_[np.arange(60), 300 + np.arange(10)].astype(np.int32)
    sim.expose(seeded, np.full(len(seeded), 3.0))
    commuters = np.arange(310, 600, dtype=np.int32)
    # the commuters visit the satellite on day 0 (share of the day before = 0: nobody can be infected) and on day 1
    for day in (0, 1):
        t_in = 24.0 * day + 8.0 + (commuters % 7) * 0.25
        sim.add_visits(commuters, np.full(len(commuters), 2, np.int32), (commuters % 3).astype(np.int16), t_in, t_in + 9.0)
    sim.run(24.0)
    assert len(sim.infections()["person"]) == 0
    sim.run(48.0)
    inf = sim.infections()
    share = 60 / 300
    want = []
    for p_id in commuters:
        t_out = 24.0 + 8.0 + (p_id % 7) * 0.25 + 9.0
        p = 1.0 - math.exp(-((factor * share) * (9.0 / 24.0)))
        if ob.u01(seed, int(p_id), bits(t_out), TAG_RATE) < p:
            want.append(int(p_id))
    assert 5 < len(want) < len(commuters)
    assert sorted(inf["person"].tolist()) == want
    assert (inf["loc"] == 2).all() and set(inf["sub"].tolist()) <= {0, 1, 2}


def test_infection_rates_file_of_the_reference_is_read():
    from heros_b200 import loaders
    path = "/root/reference/data/thehague/epidemiology/infection_rates.csv"
    if not os.path.exists(path):
        import pytest
        pytest.skip("the reference tree is not here")
    rows = loaders.read_infection_rates(path)
    assert len(rows["location"]) == 20 and rows["location"][0] == 143158 and rows["location"][-1] == 143177
    assert (rows["rate_factor"][:16] > 0).all() and (rows["rate"][:16] == -1.0).all()
    assert (rows["rate_factor"][16:] == -1.0).all() and rows["rate"][16] == 0.0009620898197102543
