"""Checkpoint / resume: a run that is snapshotted, resumed in a fresh engine and continued must be the uninterrupted run
(and the oracle's), bit for bit -- host-driven stays with pending tickets at the snapshot, and the device-side schedule
with a closure policy and a parameter-change schedule in force."""
import numpy as np
import pytest

import heros_b200 as hb
from common import Scenario, assert_same_run, load_props, make_engine, run_oracle, sort_events
from heros_b200 import _lib, properties, scenario

pytestmark = pytest.mark.gpu


def submit_day(eng, scn, d):
    """the stays that begin on day d (still open at midnight: t_out = +inf) and the ones opened earlier that close on it"""
    t, t1 = 24.0 * d, 24.0 * (d + 1)
    begins = (scn.t_in >= t) & (scn.t_in < t1)
    t_out = np.where(begins & (scn.t_out >= t1), np.inf, scn.t_out)
    sel = np.nonzero(begins)[0]
    closing = np.nonzero((scn.t_in < t) & (scn.t_out >= t) & (scn.t_out < t1))[0]
    idx = np.concatenate([sel, closing])
    eng.submit_visits(scn.person[idx], scn.loc[idx], scn.sub[idx], scn.t_in[idx], np.concatenate([t_out[sel], scn.t_out[closing]]))


@pytest.mark.parametrize("model", ["area", "distance"])
def test_resumed_run_equals_the_uninterrupted_run_and_the_oracle(model):
    props = load_props(model)
    props["covidT_area.contagiousness" if model == "area" else "covidT_dist.alpha"] = "6.0" if model == "area" else "2.0"
    kind = _lib.MODEL_AREA if model == "area" else _lib.MODEL_DISTANCE
    policies = () if model == "area" else ((72.0, "psi", 2.0), (72.0, "mu", 0.3), (150.0, "mu", 0.0))
    scn = Scenario(seed=8, n_persons=900, n_households=300, n_workplaces=30, n_shops=6, days=8)
    ora = run_oracle(scn, props, kind, _lib.TRIGGER_EXIT_ONLY, 77, policies=policies)

    first = make_engine(scn, props, kind, _lib.TRIGGER_EXIT_ONLY, 77, policies=policies)
    first.expose(scn.seeds, np.zeros(len(scn.seeds)))
    counts = []
    for d in range(5):
        submit_day(first, scn, d)
        first.step(24.0 * (d + 1))
        counts.append(first.phase_counts())
    submit_day(first, scn, 5)                    # pending stays with their tickets are part of the snapshot
    blob = first.snapshot().copy()
    first.close()

    resumed = make_engine(scn, props, kind, _lib.TRIGGER_EXIT_ONLY, 77)   # static inputs only: no policies, no seeds
    resumed.restore(blob.tobytes())
    assert resumed.time == 120.0
    np.testing.assert_array_equal(resumed.phase_counts(), counts[-1])
    for d in range(5, 8):
        if d > 5:
            submit_day(resumed, scn, d)
        resumed.step(24.0 * (d + 1))
        counts.append(resumed.phase_counts())
    gpu = dict(infections=resumed.infections(), transitions=resumed.transitions(), counts=np.array(counts),
               state=resumed.agent_state())
    assert_same_run(gpu, ora)
    assert len(ora["infections"]["person"]) > 30
    times, series = resumed.count_series()
    assert len(times) == 8 and (series[-1] == counts[-1]).all()          # the compartment series continues

    other = Scenario(seed=9, n_persons=500, n_households=200, n_workplaces=20, n_shops=4, days=2)
    wrong = make_engine(other, props, kind, _lib.TRIGGER_EXIT_ONLY, 77)
    with pytest.raises(hb.HerosError):
        wrong.restore(blob)                      # other persons / locations
    with pytest.raises(hb.HerosError):
        resumed.restore(blob[:1000])             # truncated
    wrong.close()
    resumed.close()


def test_resume_with_device_schedule_and_closure_policy():
    city = scenario.City(5000, seed=17)
    props = load_props("distance")
    props["covidT_dist.alpha"] = "2.5"
    tables = scenario.compile_schedule(city)
    T = scenario.TYPE_ID

    def make():
        eng = hb.HerosEngine(hb.MODEL_DISTANCE, 111)
        city.install(eng)
        eng.set_transmission_distance(properties.dist_params(props))
        eng.set_progression(properties.prog_params(props))
        eng.set_schedule(tables, 111)
        return eng

    def day(eng, d):
        if d == 3:
            eng.set_closure_policy(T["Workplace"], 0.3, 0.5, T["Accommodation"])
            eng.set_closure_policy(T["PrimarySchool"], 0.0, 0.0, T["Accommodation"])
        if d == 8:
            eng.set_closure_policy(T["Workplace"], 1.0, 1.0, T["Workplace"])
        eng.advance_schedule(d)
        eng.step(24.0 * (d + 1))

    whole = make()
    whole.seed_infections(80)
    for d in range(11):
        day(whole, d)
    part = make()
    part.seed_infections(80)
    for d in range(6):
        day(part, d)
    blob = part.snapshot().copy()
    part.close()
    rest = make()
    rest.restore(blob)
    for d in range(6, 11):
        day(rest, d)
    a, b = sort_events(whole.infections()), sort_events(rest.infections())
    assert len(a["person"]) > 100
    for k in ("person", "loc", "sub", "time", "p"):
        np.testing.assert_array_equal(a[k], b[k])
    ta, tb = sort_events(whole.transitions()), sort_events(rest.transitions())
    for k in ("person", "phase", "time"):
        np.testing.assert_array_equal(ta[k], tb[k])
    np.testing.assert_array_equal(whole.phase_counts(), rest.phase_counts())
    sa, sb = whole.debug_stays(), rest.debug_stays()
    for k in sa:
        np.testing.assert_array_equal(np.sort(sa[k]), np.sort(sb[k]))
    whole.close()
    rest.close()
