"""Device-side activity schedule (heros_advance_schedule) against the host-side schedule generator that interprets the
same activity rows: the stays of every day must be the same set, bit for bit, and so must the epidemic they produce
(the phase held at midnight selects the next day's pattern, so the comparison covers the feedback too)."""
import numpy as np
import pytest

import heros_b200 as hb
from common import load_props, sort_events
from heros_b200 import properties, scenario

pytestmark = pytest.mark.gpu


def make(city, props, model):
    eng = hb.HerosEngine(model, 111)
    city.install(eng)
    if model == hb.MODEL_AREA:
        eng.set_transmission_area(properties.area_params(props))
    else:
        eng.set_transmission_distance(properties.dist_params(props))
    eng.set_progression(properties.prog_params(props))
    eng.seed_infections(60)
    return eng


def stay_table(st):
    order = np.lexsort((st["t_out"], st["t_in"], st["sub"], st["loc"], st["person"]))
    return np.stack([st["person"][order].astype(np.float64), st["loc"][order].astype(np.float64),
                     st["sub"][order].astype(np.float64), st["t_in"][order], st["t_out"][order]])


@pytest.mark.parametrize("model,capacity", [("area", False), ("distance", False), ("distance", True)])
def test_device_schedule_equals_host_generator(model, capacity):
    city = scenario.City(6000, seed=7)
    if capacity:
        # capacity-constrained locators: over-subscribed places shed a keyed share of their visits to the next candidate
        # (scenario.capacity_diversion -> heros_set_capacity_diversion); floor areas cut down so that many places saturate
        city.loc_area = (city.loc_area * np.where(scenario.TYPE_CAP_CONSTRAINED[city.loc_type] == 1, 0.05, 1.0)).astype(np.float32)
        share = scenario.capacity_diversion(city, 111)
        assert (share > 0.3).sum() > 5 and share.max() < 1.0
    props = load_props(model)
    props["covidT_area.contagiousness" if model == "area" else "covidT_dist.alpha"] = "6.0" if model == "area" else "2.5"
    kind = hb.MODEL_AREA if model == "area" else hb.MODEL_DISTANCE
    days = 16  # two weekends; the first seeds turn symptomatic (home all day) and change their pattern
    host, dev = make(city, props, kind), make(city, props, kind)
    gen = scenario.ScheduleGenerator(city, 111)
    dev.set_schedule(scenario.compile_schedule(city), 111)
    n_stays = 0
    for d in range(days):
        st = gen.generate_day(host.agent_state()["phase"])
        host.submit_visits(st["person"], st["loc"], st["sub"], st["t_in"], st["t_out"])
        dev.advance_schedule(d)
        a, b = stay_table(host.debug_stays()), stay_table(dev.debug_stays())
        assert a.shape == b.shape, (d, a.shape, b.shape)
        np.testing.assert_array_equal(a, b)   # same stays, bit-identical times
        n_stays += a.shape[1]
        host.step(24.0 * (d + 1))
        dev.step(24.0 * (d + 1))
        np.testing.assert_array_equal(host.phase_counts(), dev.phase_counts())
    assert n_stays > 100000
    hi, di = sort_events(host.infections()), sort_events(dev.infections())
    assert len(hi["person"]) > 200
    for k in ("person", "loc", "sub", "time", "p"):
        np.testing.assert_array_equal(hi[k], di[k])
    ht, dt = sort_events(host.transitions()), sort_events(dev.transitions())
    for k in ("person", "phase", "time"):
        np.testing.assert_array_equal(ht[k], dt[k])
    assert (host.agent_state()["phase"] >= 3).sum() > 10   # symptomatic / hospitalised persons changed their patterns
    host.close()
    dev.close()


def test_device_schedule_with_closure_policies_equals_host_generator():
    """LocationPolicy rows (partial closure 0.2 / 0.5 of workplaces, 0.1 / 0.1 of retail, full closure of the schools,
    reopening) through heros_set_closure_policy and through the host-side generator: the same stays bit for bit, the
    same epidemic."""
    city = scenario.City(6000, seed=11)
    props = load_props("distance")
    props["covidT_dist.alpha"] = "2.5"
    T = scenario.TYPE_ID
    host, dev = make(city, props, hb.MODEL_DISTANCE), make(city, props, hb.MODEL_DISTANCE)
    gen = scenario.ScheduleGenerator(city, 111)
    dev.set_schedule(scenario.compile_schedule(city), 111)

    class Model:          # the queue the LocationPolicy mirror needs; the engine receives the device-side calls
        engine = dev
    rows = [(2.0, "Workplace", 0.2, 0.5, "Accommodation"), (2.0, "Retail", 0.1, 0.1, "Accommodation"),
            (2.0, "PrimarySchool", 0.0, 0.0, "Accommodation"), (2.0, "SecondarySchool", 0.0, 0.0, "Accommodation"),
            (2.0, "BarRestaurant", 0.0, 1.0, "Accommodation"), (2.0, "Park", 1.0, 0.3, "Accommodation"),
            (9.0, "Workplace", 1.0, 1.0, "Workplace"), (9.0, "PrimarySchool", 1.0, 1.0, "PrimarySchool"),
            (9.0, "Park", 0.5, 1.0, "Accommodation")]
    for day, name, fo, fa, alt in rows:
        hb.LocationPolicy(Model, 24.0 * day, name, fo, fa, alt, "report")
    work_stays = []
    for d in range(12):
        due = [p for p in Model.locationPolicies if p.time <= 24.0 * d]
        for p in due:
            p.openCloseLocationType(gen)       # host-side schedule
        assert hb.apply_location_policies(Model, 24.0 * d) == len(due)   # device-side schedule
        st = gen.generate_day(host.agent_state()["phase"])
        work_stays.append(int((city.loc_type[st["loc"]] == T["Workplace"]).sum()))
        host.submit_visits(st["person"], st["loc"], st["sub"], st["t_in"], st["t_out"])
        dev.advance_schedule(d)
        a, b = stay_table(host.debug_stays()), stay_table(dev.debug_stays())
        assert a.shape == b.shape, (d, a.shape, b.shape)
        np.testing.assert_array_equal(a, b)
        if 3 <= d < 9:
            assert (city.loc_type[st["loc"]] == T["PrimarySchool"]).sum() == 0
            assert (city.loc_type[st["loc"]] == T["BarRestaurant"]).sum() == 0
        host.step(24.0 * (d + 1))
        dev.step(24.0 * (d + 1))
        np.testing.assert_array_equal(host.phase_counts(), dev.phase_counts())
    # closed from day 2, open again from day 9 (by then part of the workers is ill and stays at home anyway)
    assert work_stays[3] < 0.3 * work_stays[1] and work_stays[10] > 3 * work_stays[3]
    hi, di = sort_events(host.infections()), sort_events(dev.infections())
    assert len(hi["person"]) > 50
    for k in ("person", "loc", "sub", "time", "p"):
        np.testing.assert_array_equal(hi[k], di[k])
    with pytest.raises(hb.HerosError):
        dev.set_closure_policy(T["Retail"], 0.5, 0.5, T["Park"])      # neither home nor the type itself
    with pytest.raises(hb.HerosError):
        dev.set_closure_policy(T["Retail"], 1.5, 0.5, T["Accommodation"])
    with pytest.raises(hb.HerosError):
        dev.set_closure_policy(99, 0.5, 0.5, T["Accommodation"])
    host.close()
    dev.close()


def test_schedule_errors():
    city = scenario.City(800, seed=3)
    props = load_props("distance")
    eng = make(city, props, hb.MODEL_DISTANCE)
    with pytest.raises(hb.HerosError):
        eng.advance_schedule(0)               # heros_set_schedule has not been called
    with pytest.raises(hb.HerosError):
        eng.set_closure_policy(1, 0.0, 0.0, 0)   # needs the schedule tables
    tables = scenario.compile_schedule(city)
    bad = dict(tables)
    bad["pattern_count"] = tables["pattern_count"].copy()
    bad["pattern_count"][5] = 10 ** 6
    with pytest.raises(hb.HerosError):
        eng.set_schedule(bad, 1)
    missing = dict(tables)
    missing["pattern_count"] = tables["pattern_count"].copy()
    role = int(city.role[0])
    missing["pattern_count"][(3 * 16 + role) * 4: (3 * 16 + role) * 4 + 4] = 0   # no "0_Infected-Symptomatic_<role>"
    with pytest.raises(hb.HerosError) as err:                                    # HerosModel.java:497-501
        eng.set_schedule(missing, 1)
    assert "Week pattern key" in str(err.value) and "not found in week pattern map" in str(err.value)
    eng.set_schedule(tables, 1)
    eng.advance_schedule(0)
    eng.step(24.0)
    with pytest.raises(hb.HerosError):
        eng.advance_schedule(0)               # the day lies before the engine time
    eng.close()
