"""Rate-based locations on the engine (heros_set_rate_locations -> k_rate_locations) against the oracle: a city in which
some workplaces are satellites -- two with a fixed daily rate, three that follow the reference group -- run day by day and
in 6 h windows; infection events (probabilities included), transitions and counts identical, and the occupancy kernels
never evaluate a satellite."""
import numpy as np
import pytest

from common import Scenario, assert_same_run, load_props, run_gpu, run_oracle
from heros_b200 import _lib

pytestmark = pytest.mark.gpu


def satellite_city(seed=31):
    scn = Scenario(seed=seed, n_persons=1500, n_households=520, n_workplaces=40, n_shops=5, days=10, shop_visits_per_day=1.5)
    rng = np.random.default_rng(seed)
    scn.role = np.where(rng.random(scn.n_persons) < 0.6, 7, 12).astype(np.uint8)  # Workers and satellite commuters
    w0 = int(np.nonzero(scn.loc_type == 1)[0][0])
    sat = np.arange(w0, w0 + 5, dtype=np.int32)   # the first five workplaces are satellites
    scn.rate_locations = (sat, np.array([-1.0, -1.0, 6.0, 9.0, 12.0]), np.array([0.8, 0.4, -1.0, -1.0, -1.0]))
    return scn, sat


@pytest.mark.parametrize("model,step_hours", [("distance", 24.0), ("area", 6.0)])
def test_rate_based_locations_equal_the_oracle(model, step_hours):
    scn, sat = satellite_city()
    props = load_props(model)
    props["covidT_area.contagiousness" if model == "area" else "covidT_dist.alpha"] = "1.5" if model == "area" else "0.3"
    kind = _lib.MODEL_AREA if model == "area" else _lib.MODEL_DISTANCE
    ora = run_oracle(scn, props, kind, _lib.TRIGGER_EXIT_ONLY, seed=17)
    gpu = run_gpu(scn, props, kind, _lib.TRIGGER_EXIT_ONLY, seed=17, window_hours=step_hours, step_hours=step_hours)
    assert_same_run(gpu, ora)
    inf = gpu["infections"]
    in_sat = np.isin(inf["loc"], sat)
    assert in_sat.sum() > 20 and (~in_sat).sum() > 100
    # both kinds of satellite infected somebody
    assert np.isin(inf["loc"], sat[:2]).any() and np.isin(inf["loc"], sat[2:]).any()
    gpu["engine"].close()


def test_factor_driven_location_refuses_a_window_across_midnight():
    scn, _ = satellite_city()
    props = load_props("distance")
    from common import make_engine
    eng = make_engine(scn, props, _lib.MODEL_DISTANCE, _lib.TRIGGER_EXIT_ONLY, 17, window_hours=36.0)
    eng.submit_visits(scn.person, scn.loc, scn.sub, scn.t_in, scn.t_out)
    with pytest.raises(Exception) as e:
        eng.step(36.0)
    assert "midnight" in str(e.value)
    eng.close()


def test_unordered_getter_returns_the_same_events():
    """heros_get_infections_unordered: the events as the device logged them (windows in order), fetched day by day with a
    running count -- the same set as the (time, person)-ordered getter."""
    from common import make_engine
    scn, _ = satellite_city(seed=33)
    props = load_props("distance")
    props["covidT_dist.alpha"] = "1.0"
    eng = make_engine(scn, props, _lib.MODEL_DISTANCE, _lib.TRIGGER_EXIT_ONLY, 17)
    eng.expose(scn.seeds, np.zeros(len(scn.seeds)))
    eng.submit_visits(scn.person, scn.loc, scn.sub, scn.t_in, scn.t_out)
    got = {k: [] for k in ("person", "loc", "sub", "time", "p")}
    n = 0
    for d in range(scn.days):
        eng.step(24.0 * (d + 1))
        new = eng.infections(first=n, ordered=False)
        assert ((new["time"] >= 24.0 * d) & (new["time"] < 24.0 * (d + 1))).all()   # the events of this window only
        n += len(new["person"])
        for k in got:
            got[k].append(new[k])
    got = {k: np.concatenate(v) for k, v in got.items()}
    want = eng.infections()
    assert len(want["person"]) == n > 100
    order = np.lexsort((got["person"], got["time"]))
    for k in got:
        np.testing.assert_array_equal(got[k][order], want[k])
    eng.close()


def test_window_results_one_window_behind():
    """heros_window_results / heros_get_infections_range: the counts and the events of every window of a run whose windows
    were all enqueued at once (heros_step_async), read window by window -- equal to a run that waits after every day."""
    from common import make_engine
    scn, _ = satellite_city(seed=35)
    props = load_props("distance")
    props["covidT_dist.alpha"] = "1.0"
    runs = []
    for lagged in (False, True):
        eng = make_engine(scn, props, _lib.MODEL_DISTANCE, _lib.TRIGGER_EXIT_ONLY, 17)
        eng.expose(scn.seeds, np.zeros(len(scn.seeds)))
        eng.submit_visits(scn.person, scn.loc, scn.sub, scn.t_in, scn.t_out)
        counts, events, n = [], [], 0
        if lagged:
            for d in range(scn.days):
                eng.step_async(24.0 * (d + 1))
            for d in range(scn.days):
                c, logged = eng.window_results(d)
                counts.append(c)
                events.append(eng.infections_range(n, logged - n))
                n = logged
            eng.wait()
        else:
            for d in range(scn.days):
                eng.step(24.0 * (d + 1))
                counts.append(eng.phase_counts())
                new = eng.infections(first=n, ordered=False)
                events.append(new)
                n += len(new["person"])
        runs.append((np.array(counts), events, n))
        eng.close()
    np.testing.assert_array_equal(runs[0][0], runs[1][0])
    assert runs[0][2] == runs[1][2] > 100
    for a, b in zip(runs[0][1], runs[1][1]):
        oa, ob = np.lexsort((a["person"], a["time"])), np.lexsort((b["person"], b["time"]))
        for k in a:
            np.testing.assert_array_equal(a[k][oa], b[k][ob])
