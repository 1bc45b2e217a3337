"""GPU parity: the windowed CUDA engine (through the C ABI) against the sequential CPU oracle on the same stays."""
import numpy as np
import pytest

from common import Scenario, assert_same_run, load_props, make_engine, run_gpu, run_oracle
import heros_b200 as hb
from heros_b200 import _lib

pytestmark = pytest.mark.gpu

MODELS = {"area": _lib.MODEL_AREA, "distance": _lib.MODEL_DISTANCE}
TRIGGERS = {"exit": _lib.TRIGGER_EXIT_ONLY, "both": _lib.TRIGGER_ENTER_AND_EXIT}


def hot_props(model):
    """Reference alpha parameters with the base contagiousness raised so that a few-hundred-person city produces
    hundreds of infections within two weeks."""
    p = load_props(model)
    if model == "area":
        p["covidT_area.contagiousness"] = "4.0"
    else:
        p["covidT_dist.alpha"] = "1.5"
    return p


@pytest.mark.parametrize("model", ["area", "distance"])
@pytest.mark.parametrize("trigger", ["exit", "both"])
def test_small_city_matches_oracle(model, trigger):
    scn = Scenario(seed=11, n_persons=600, n_households=220, n_workplaces=25, n_shops=6, days=14)
    props = hot_props(model)
    ora = run_oracle(scn, props, MODELS[model], TRIGGERS[trigger], seed=111)
    gpu = run_gpu(scn, props, MODELS[model], TRIGGERS[trigger], seed=111)
    assert len(ora["infections"]["person"]) > 50
    assert_same_run(gpu, ora)
    assert sum(s["evaluations"] for s in gpu["stats"]) == ora["evaluations"]
    assert sum(s["near_ties"] for s in gpu["stats"]) == 0


@pytest.mark.parametrize("model,trigger", [("area", "both"), ("distance", "exit")])
def test_daily_submission_with_open_stays(model, trigger):
    """The host submits day by day; stays still open at midnight go in with t_out = +inf and are closed later."""
    scn = Scenario(seed=12, n_persons=500, n_households=180, n_workplaces=20, n_shops=5, days=10, open_fraction=0.1)
    props = hot_props(model)
    ora = run_oracle(scn, props, MODELS[model], TRIGGERS[trigger], seed=7)
    gpu = run_gpu(scn, props, MODELS[model], TRIGGERS[trigger], seed=7, submit="daily")
    assert len(ora["infections"]["person"]) > 30
    assert_same_run(gpu, ora)


@pytest.mark.parametrize("window,step", [(6.0, 24.0), (12.0, 18.0), (47.0, 47.0), (24.0, 5.0)])
def test_window_length_does_not_change_the_result(window, step):
    """batch == sequential for any window below the latent period (48 h): same events as the oracle."""
    scn = Scenario(seed=13, n_persons=400, n_households=150, n_workplaces=15, n_shops=4, days=9)
    props = hot_props("area")
    ora = run_oracle(scn, props, _lib.MODEL_AREA, _lib.TRIGGER_ENTER_AND_EXIT, seed=3)
    gpu = run_gpu(scn, props, _lib.MODEL_AREA, _lib.TRIGGER_ENTER_AND_EXIT, seed=3, window_hours=window,
                  step_hours=step)
    gpu["counts"] = ora["counts"]  # day-end counts are only sampled when a step ends on a day boundary
    assert_same_run(gpu, ora)


def test_window_longer_than_latent_period_is_refused():
    import heros_b200 as hb
    scn = Scenario(seed=1, n_persons=50, n_households=20, n_workplaces=3, n_shops=1, days=1)
    with pytest.raises(hb.HerosError) as e:
        run_gpu(scn, hot_props("area"), _lib.MODEL_AREA, _lib.TRIGGER_EXIT_ONLY, seed=1, window_hours=48.0)
    assert e.value.code == _lib.ERR_UNSUPPORTED


def test_hot_sublocations_use_the_block_kernel_and_match():
    """A city with two shops: thousands of stays per sub-location and day, second-scale gaps, tied event times."""
    scn = Scenario(seed=14, n_persons=1500, n_households=500, n_workplaces=6, n_shops=2, days=8,
                   shop_visits_per_day=3.0, tie_fraction=0.3, short_gap_fraction=0.3)
    for model, trigger in (("area", "both"), ("distance", "exit"), ("distance", "both")):
        props = hot_props(model)
        if model == "area":
            props["covidT_area.contagiousness"] = "20.0"
        ora = run_oracle(scn, props, MODELS[model], TRIGGERS[trigger], seed=5)
        gpu = run_gpu(scn, props, MODELS[model], TRIGGERS[trigger], seed=5)
        assert sum(s["big_sublocations"] for s in gpu["stats"]) > 50
        assert len(ora["infections"]["person"]) > 100
        assert_same_run(gpu, ora)
        assert sum(s["evaluations"] for s in gpu["stats"]) == ora["evaluations"]


def test_disease_policy_psi_mu_changes_take_effect_at_their_time():
    """data/thehague/policies/MaskingDistance_10-40.csv scaled to days 3 and 6 (DiseasePolicy.java:29-43)."""
    scn = Scenario(seed=15, n_persons=500, n_households=180, n_workplaces=20, n_shops=5, days=9)
    props = hot_props("distance")
    policies = [(72.0, "psi", 2.0), (72.0, "mu", 0.6), (144.0, "psi", 1.0), (144.0, "mu", 0.0)]
    ora = run_oracle(scn, props, _lib.MODEL_DISTANCE, _lib.TRIGGER_EXIT_ONLY, seed=9, policies=policies)
    plain = run_oracle(scn, props, _lib.MODEL_DISTANCE, _lib.TRIGGER_EXIT_ONLY, seed=9)
    gpu = run_gpu(scn, props, _lib.MODEL_DISTANCE, _lib.TRIGGER_EXIT_ONLY, seed=9, policies=policies)
    assert len(plain["infections"]["person"]) != len(ora["infections"]["person"])  # the policy matters
    assert_same_run(gpu, ora)


def test_set_parameter_error_codes():
    scn = Scenario(seed=1, n_persons=50, n_households=20, n_workplaces=3, n_shops=1, days=1)
    from common import make_engine
    eng = make_engine(scn, hot_props("distance"), _lib.MODEL_DISTANCE, 0, 1)
    assert eng.set_parameter("psi", 2.0, 10.0) == _lib.OK
    assert eng.set_parameter("beta", 2.0, 10.0) == _lib.ERR_UNKNOWN_PARAMETER   # TDist:263
    eng.close()
    eng = make_engine(scn, hot_props("area"), _lib.MODEL_AREA, 0, 1)
    assert eng.set_parameter("psi", 2.0, 10.0) == _lib.ERR_UNKNOWN_PARAMETER    # TArea:251-255 rejects every name
    eng.close()


def test_generator_driven_city_with_phase_feedback():
    """The bench flow at small scale: stays come from the activity-schedule generator, which reads the engine's
    phases every midnight (symptomatic stay home, hospitalised move to the hospital)."""
    import heros_b200 as hb
    from heros_b200 import properties, scenario
    from common import oracle_props
    from oracle import binding as ob
    city = scenario.City(12000, seed=5)
    props = hot_props("distance")
    props["covidT_dist.alpha"] = "0.5"
    eng = hb.HerosEngine(hb.MODEL_DISTANCE, 111)
    city.install(eng)
    eng.set_transmission_distance(properties.dist_params(props))
    eng.set_progression(properties.prog_params(props))
    seeds = eng.seed_infections(30)
    area, dist, prog = oracle_props(props)
    sim = ob.Sim(ob.MODEL_DIST, ob.TRIGGER_EXIT_ONLY, 111)
    sim.set_dist(dist); sim.set_prog(prog)
    sim.set_location_types(city.type_infect_sub, city.type_correction)
    sim.set_locations(city.loc_type, city.loc_nsub, city.loc_area)
    sim.set_persons(city.age, city.female)
    sim.expose(seeds, np.zeros(len(seeds)))
    gen = scenario.ScheduleGenerator(city, 111)
    counts = []
    for d in range(12):
        phase = eng.agent_state()["phase"]
        np.testing.assert_array_equal(phase, sim.agent_state()["phase"])
        st = gen.generate_day(phase)
        eng.submit_visits(st["person"], st["loc"], st["sub"], st["t_in"], st["t_out"])
        eng.step(24.0 * (d + 1))
        sim.add_visits(st["person"], st["loc"], st["sub"], st["t_in"], st["t_out"])
        sim.run(24.0 * (d + 1))
        counts.append(eng.phase_counts())
    gpu = dict(infections=eng.infections(), transitions=eng.transitions(), counts=np.array(counts),
               state=eng.agent_state())
    t, series = eng.count_series()
    np.testing.assert_array_equal(series, np.array(counts))
    ora = dict(infections=sim.infections(), transitions=sim.transitions(), counts=np.array(counts),
               state=sim.agent_state())
    assert len(gpu["infections"]["person"]) > 100
    assert_same_run(gpu, ora)
    eng.close()


def test_seeding_is_the_keyed_bottom_k_and_respects_age_bounds():
    from heros_b200 import scenario
    from common import make_engine
    scn = Scenario(seed=2, n_persons=3000, n_households=900, n_workplaces=30, n_shops=5, days=1)
    eng = make_engine(scn, hot_props("area"), _lib.MODEL_AREA, 0, 111)
    chosen = eng.seed_infections(40, min_age=20, max_age=60)
    assert len(np.unique(chosen)) == 40 and (scn.age[chosen] >= 20).all() and (scn.age[chosen] <= 60).all()
    x, y, _, _ = scenario.philox4x32_10(np.arange(3000), 0, 0, 0x200, 111, 0)
    h = ((x << np.uint64(32)) | y)
    ok = (scn.age >= 20) & (scn.age <= 60)
    want = np.sort(np.nonzero(ok)[0][np.argsort(h[ok], kind="stable")[:40]])
    np.testing.assert_array_equal(chosen, want)
    c = eng.phase_counts()
    assert c[0] == 3000 - 40 and c[1] == 40
    st = eng.agent_state()
    assert (st["phase"][chosen] == 1).all() and (st["exposure"][chosen] == 0.0).all()
    assert ((st["next_time"][chosen] >= 60.0) & (st["next_time"][chosen] <= 91.2)).all()
    eng.close()


@pytest.mark.parametrize("model,trigger", [("area", "exit"), ("distance", "both")])
def test_fast_path_without_exact_stats_gives_identical_results(model, trigger):
    """Default mode: sub-locations where nobody can be infected only advance lastCalculation (no chain walk).  The
    infection events, trajectories and every lastCalculation must not change."""
    scn = Scenario(seed=16, n_persons=1200, n_households=420, n_workplaces=30, n_shops=3, days=10,
                   shop_visits_per_day=2.0, tie_fraction=0.3, short_gap_fraction=0.3)
    props = hot_props(model)
    ora = run_oracle(scn, props, MODELS[model], TRIGGERS[trigger], seed=21)
    gpu = run_gpu(scn, props, MODELS[model], TRIGGERS[trigger], seed=21, flags=0)
    assert_same_run(gpu, ora)
    assert sum(s["evaluations"] for s in gpu["stats"]) < ora["evaluations"]  # chains were skipped somewhere
    for loc in range(scn.n_loc):
        for sub in range(int(scn.loc_nsub[loc])):
            assert gpu["engine"].last_calculation(loc, sub) == ora["sim"].last_calc(loc, sub)


@pytest.mark.parametrize("model,trigger", [("area", _lib.TRIGGER_EXIT_ONLY), ("distance", _lib.TRIGGER_EXIT_ONLY),
                                           ("distance", _lib.TRIGGER_ENTER_AND_EXIT)])
def test_total_location_branch_in_windows(model, trigger):
    """Locations with infectInSublocation = false and >= 2 sub-locations (TArea:198-246 / TDist:189-246): every
    evaluation of a sub-location scans the persons of the whole location.  In the area model that branch can only
    infect when late-stage cases make the sum negative (sign quirk), in the distance model it infects normally."""
    scn = Scenario(seed=31, n_persons=700, n_households=260, n_workplaces=20, n_shops=2, days=12,
                   shop_visits_per_day=3.0, tie_fraction=0.3, n_parks=3, park_nsub=3)
    props = load_props(model)
    props["covidT_area.contagiousness" if model == "area" else "covidT_dist.alpha"] = "6.0" if model == "area" else "2.5"
    kind = _lib.MODEL_AREA if model == "area" else _lib.MODEL_DISTANCE
    ora = run_oracle(scn, props, kind, trigger, seed=13)
    gpu = run_gpu(scn, props, kind, trigger, seed=13)
    assert_same_run(gpu, ora)
    assert sum(s["evaluations"] for s in gpu["stats"]) == ora["evaluations"]
    park0 = int(np.nonzero(scn.loc_type == 3)[0][0])
    in_parks = (ora["infections"]["loc"] >= park0).sum()
    assert len(ora["infections"]["person"]) > 100
    if model == "distance":
        assert in_parks > 20          # the total branch infected people, in several zones
        assert len(np.unique(ora["infections"]["sub"][ora["infections"]["loc"] >= park0])) > 1
    gpu["engine"].close()


def test_one_crowded_sublocation_takes_the_global_scratch_class():
    """A single shop for the whole town: > 4096 stays and > 8192 movement events per day in one sub-location, which is
    beyond the shared-memory classes of the group kernels (class 4 works in global scratch); and a second run with a
    10 s threshold, where the bound on the evaluations per window (8642) exceeds what the shared-memory classes keep."""
    scn = Scenario(seed=15, n_persons=1600, n_households=550, n_workplaces=6, n_shops=1, days=5,
                   shop_visits_per_day=4.0, tie_fraction=0.3, short_gap_fraction=0.3)
    for model, trigger, threshold in (("distance", "exit", None), ("area", "both", "10")):
        props = hot_props(model)
        if model == "area":
            props["covidT_area.contagiousness"] = "20.0"
        if threshold:
            props["covidT_area.calculation_threshold"] = threshold
        ora = run_oracle(scn, props, MODELS[model], TRIGGERS[trigger], seed=6)
        gpu = run_gpu(scn, props, MODELS[model], TRIGGERS[trigger], seed=6)
        shop = scn.n_loc - 1
        per_day = ((scn.loc == shop).sum()) / scn.days
        assert per_day > 4096
        assert len(ora["infections"]["person"]) > 100
        assert_same_run(gpu, ora)
        assert sum(s["evaluations"] for s in gpu["stats"]) == ora["evaluations"]
        gpu["engine"].close()


@pytest.mark.parametrize("model", ["area", "distance"])
def test_helsinki_real_location_table(model):
    """BASELINE configs[1] as a parity case: the real Helsinki location table (69 508 locations, 431 470 sub-locations,
    among them 76 homes listed with 0 sub-locations and 3 563 workplaces with area = -1) with a
    synthetic population and the reference's activity schedule; engine vs oracle, event for event.  The degenerate rows
    must behave as Java's IEEE arithmetic does: NaN / negative probabilities infect nobody."""
    import os
    from common import GOLDEN, sort_events
    from heros_b200 import properties, scenario
    from oracle import binding as ob
    from common import oracle_props
    tab = dict(np.load(os.path.join(GOLDEN, "helsinki_locations.npz")))
    city = scenario.City(24000, seed=5, location_table=tab)
    props = load_props(model)
    props["covidT_area.contagiousness" if model == "area" else "covidT_dist.alpha"] = "8.0" if model == "area" else "2.5"
    kind = MODELS[model]
    area, dist, prog = oracle_props(props)
    sim = ob.Sim(kind, ob.TRIGGER_EXIT_ONLY, 111)
    sim.set_area(area) if model == "area" else sim.set_dist(dist)
    sim.set_prog(prog)
    sim.set_location_types(city.type_infect_sub, city.type_correction)
    sim.set_locations(city.loc_type, city.loc_nsub, city.loc_area)
    sim.set_persons(city.age, city.female)
    eng = hb.HerosEngine(kind, 111, flags=_lib.FLAG_EXACT_STATS)
    city.install(eng)
    if model == "area":
        eng.set_transmission_area(properties.area_params(props))
    else:
        eng.set_transmission_distance(properties.dist_params(props))
    eng.set_progression(properties.prog_params(props))
    seeds = np.arange(0, city.n_persons, 40, dtype=np.int32)
    sim.expose(seeds, np.zeros(len(seeds)))
    eng.expose(seeds, np.zeros(len(seeds)))
    gen = scenario.ScheduleGenerator(city, 111)
    in_bad_work = 0
    evaluations = 0
    for d in range(9):
        st = gen.generate_day(eng.agent_state()["phase"])
        in_bad_work += int((city.loc_area[st["loc"]] < 0).sum())
        sim.add_visits(st["person"], st["loc"], st["sub"], st["t_in"], st["t_out"])
        eng.submit_visits(st["person"], st["loc"], st["sub"], st["t_in"], st["t_out"])
        sim.run(24.0 * (d + 1))
        evaluations += eng.step(24.0 * (d + 1))["evaluations"]
        np.testing.assert_array_equal(eng.phase_counts(), sim.phase_counts())
    assert in_bad_work > 1000                                   # people did work in the area = -1 workplaces
    gi, oi = sort_events(eng.infections()), sort_events(sim.infections())
    assert len(oi["person"]) > 150
    for k in ("person", "loc", "sub", "time"):
        np.testing.assert_array_equal(gi[k], oi[k])
    np.testing.assert_allclose(gi["p"], oi["p"], rtol=1e-12, atol=4.5e-16)
    assert not (city.loc_area[oi["loc"]] < 0).any()             # nobody was infected in a negative-area workplace
    assert evaluations == sim.n_evaluations()
    gt, ot = sort_events(eng.transitions()), sort_events(sim.transitions())
    for k in ("person", "phase", "time"):
        np.testing.assert_array_equal(gt[k], ot[k])
    eng.close()


def test_logs_can_be_read_window_by_window():
    """heros_get_infections / heros_get_transitions with a cursor (what a host does after every window): the pieces read
    day by day are the complete, ordered logs."""
    scn = Scenario(seed=8, n_persons=600, n_households=220, n_workplaces=20, n_shops=4, days=10)
    props = hot_props("area")
    eng = make_engine(scn, props, _lib.MODEL_AREA, _lib.TRIGGER_EXIT_ONLY, 4)
    eng.expose(scn.seeds, np.zeros(len(scn.seeds)))
    eng.submit_visits(scn.person, scn.loc, scn.sub, scn.t_in, scn.t_out)
    inf_parts, tr_parts, n_inf, n_tr = [], [], 0, 0
    for d in range(scn.days):
        eng.step(24.0 * (d + 1))
        a, b = eng.infections(first=n_inf), eng.transitions(first=n_tr)
        n_inf += len(a["person"]); n_tr += len(b["person"])
        inf_parts.append(a); tr_parts.append(b)
        assert (a["time"] >= 24.0 * d).all() and (a["time"] < 24.0 * (d + 1)).all()
    whole_inf, whole_tr = eng.infections(), eng.transitions()
    assert len(whole_inf["person"]) > 50 and len(whole_tr["person"]) > len(whole_inf["person"])
    for k in whole_inf:
        np.testing.assert_array_equal(np.concatenate([p[k] for p in inf_parts]), whole_inf[k])
    for k in whole_tr:
        np.testing.assert_array_equal(np.concatenate([p[k] for p in tr_parts]), whole_tr[k])
    assert (np.diff(whole_tr["time"]) >= 0).all()
    eng.close()
