"""The properties of the reference's published epidemic curves (tests/golden/seed_comparison.json, policy_comparison.json)
on the ENGINE: the synthetic-city run that tests/test_oracle.py compares with those curves on the CPU oracle, carried out
here on the GPU with the device-side schedule and the reference's lockdown rows through heros_set_closure_policy.  The
engine's infection list must be the oracle's, event for event; the profile properties then hold for the product, not
only for its checker."""
import json
import math
import os

import numpy as np
import pytest

import heros_b200 as hb
from common import GOLDEN, P_ATOL, P_RTOL, load_props, sort_events
from heros_b200 import properties, scenario
from test_oracle import SYNTHETIC_CITY, synthetic_city_infections, synthetic_city_lockdown_rows

pytestmark = pytest.mark.gpu


def engine_run(lockdown_day=None):
    cfg = SYNTHETIC_CITY
    city = scenario.City(cfg["n_persons"], seed=cfg["city_seed"])
    props = load_props("area")
    props["covidT_area.contagiousness"] = "1.0"
    eng = hb.HerosEngine(hb.MODEL_AREA, cfg["seed"])
    city.install(eng)
    eng.set_transmission_area(properties.area_params(props))
    eng.set_progression(properties.prog_params(props))
    eng.expose(cfg["first_cases"], np.zeros(len(cfg["first_cases"])))
    eng.set_schedule(scenario.compile_schedule(city), cfg["seed"])
    for d in range(cfg["days"]):
        if d == lockdown_day:
            for _, name, fraction_open, fraction_activities, alternative, _ in synthetic_city_lockdown_rows():
                if name in scenario.TYPE_ID:
                    eng.set_closure_policy(scenario.TYPE_ID[name], float(fraction_open), float(fraction_activities),
                                           scenario.TYPE_ID[alternative])
        eng.advance_schedule(d)
        eng.step(24.0 * (d + 1))
    inf = sort_events(eng.infections())
    eng.close()
    return inf


@pytest.mark.parametrize("lockdown_day", [None, 10])
def test_engine_reproduces_the_run_the_reference_curves_are_compared_with(lockdown_day):
    got = engine_run(lockdown_day)
    want = sort_events(synthetic_city_infections("area", lockdown_day))
    assert len(got["person"]) == len(want["person"]) > 500
    for k in ("person", "loc", "sub", "time"):
        np.testing.assert_array_equal(got[k], want[k])
    np.testing.assert_allclose(got["p"], want["p"], rtol=P_RTOL, atol=P_ATOL)

    t = got["time"]
    daily = np.bincount((t // 24.0).astype(int), minlength=16).astype(float)
    assert daily[0] == 0 and daily[1] == 0                       # 48 h latent period
    assert math.ceil(t.min() * 2) / 2 == 54.5                    # the sample of the first exposure in all four reference curves
    hist = np.bincount(np.ceil((t % 24.0) * 2).astype(int) % 48, minlength=48)
    assert hist[1:13].sum() == 0                                 # nothing between midnight and 06:00
    with open(os.path.join(GOLDEN, "seed_comparison.json")) as f:
        seeds = json.load(f)["seeds"]
    ref = np.sum([np.array(v["new_exposures_by_half_hour_of_day"]) for v in seeds.values()], axis=0)
    assert hist[13:15].sum() / hist.sum() > ref[13:15].sum() / ref.sum() > 0.05    # the area model's morning exposures

    def dip(new, d):
        return new[d] / math.sqrt(new[d - 1] * new[d + 1])

    assert dip(daily, 6) < 0.7                                   # Sunday of the first week, as in the reference
    if lockdown_day is None:
        assert dip(daily, 13) < 0.7
    else:
        free = np.bincount((synthetic_city_infections("area")["time"] // 24.0).astype(int), minlength=16)
        assert (daily[:10] == free[:10]).all()                   # one run until the policy
        assert 0.1 < daily[10:].sum() / free[10:].sum() < 0.5    # the reference: 0.20 over the same days
