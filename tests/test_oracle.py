"""The CPU oracle against everything that can pin it without the (non-runnable) JVM reference:
Random123 known-answer vectors, libm, the worked example of the reference's Javadoc, hand-derived boundary cases."""
import json
import math
import os

import numpy as np
import pytest

from common import GOLDEN, load_props, oracle_props
from oracle import binding as ob

S, E, IA, IS, H, ICU, D, R = range(8)


def test_philox_random123_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        assert ob.philox4x32_10(ctr, key).tolist() == want


def test_u01_range_and_keying():
    us = [ob.u01(111, p, 0x4054000000000000, 0) for p in range(2000)]
    assert 0.0 <= min(us) and max(us) < 1.0
    assert abs(np.mean(us) - 0.5) < 0.03
    assert ob.u01(111, 5, 1234, 0) == ob.u01(111, 5, 1234, 0)
    assert ob.u01(111, 5, 1234, 0) != ob.u01(112, 5, 1234, 0)
    assert ob.u01(111, 5, 1234, 0) != ob.u01(111, 5, 1235, 0)
    assert ob.u01(111, 5, 1234, 0) != ob.u01(111, 5, 1234, 1)


def test_exp_within_one_ulp_of_libm_and_special_values():
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.uniform(-700, 700, 20000), rng.uniform(-2, 2, 20000), rng.uniform(-1e-6, 1e-6, 5000)])
    for x in xs:
        a, b = ob.exp(x), math.exp(x)
        assert abs(a - b) <= math.ulp(b), x
    assert ob.exp(0.0) == 1.0 and ob.exp(-0.0) == 1.0
    assert ob.exp(float("inf")) == float("inf") and ob.exp(float("-inf")) == 0.0
    assert math.isnan(ob.exp(float("nan")))
    assert ob.exp(710.0) == float("inf") and ob.exp(-746.0) == 0.0
    assert ob.exp(-745.0) > 0.0  # subnormal result


def test_java_max_propagates_nan():
    nan = float("nan")
    assert math.isnan(ob.lib().orc_java_max(nan, 1.0)) and math.isnan(ob.lib().orc_java_max(1.0, nan))
    assert ob.lib().orc_java_max(2.0, 1.0) == 2.0 and ob.lib().orc_java_max(-1.0, 1.0) == 1.0


def test_duration_sampling_matches_closed_forms():
    # Triangular(2.5, 3.4, 3.8) days -> hours; Uniform; Constant
    assert ob.sample_duration_hours(ob.DIST_TRIANGULAR, 2.5, 3.4, 3.8, 0.0) == 2.5 * 24.0
    assert ob.sample_duration_hours(ob.DIST_TRIANGULAR, 2.5, 3.4, 3.8, 1.0) == 3.8 * 24.0
    fc = (3.4 - 2.5) / (3.8 - 2.5)
    assert abs(ob.sample_duration_hours(ob.DIST_TRIANGULAR, 2.5, 3.4, 3.8, fc) - 3.4 * 24.0) < 1e-9
    us = np.linspace(0.001, 0.999, 999)
    xs = np.array([ob.sample_duration_hours(ob.DIST_TRIANGULAR, 2.5, 3.4, 3.8, u) for u in us])
    assert (np.diff(xs) > 0).all() and 60.0 <= xs.min() and xs.max() <= 91.2  # the 60-91.2 h band of SURVEY C2
    assert ob.sample_duration_hours(ob.DIST_UNIFORM, 1.0, 3.0, 0.0, 0.25) == 1.5 * 24.0
    assert ob.sample_duration_hours(ob.DIST_CONSTANT, 4.0, 0.0, 0.0, 0.9) == 96.0


def area(pB=0.51, beta=1.0, thr=60.0):
    return ob.AreaProps(pB, beta, 2.0, 3.4, 9.6, thr)


def one_call(props, area_m2, duration, n_sub=1, phases=(IS, S), exposure=(0.0, 0.0), now=3.4 * 24, corr=1.0,
             model=ob.MODEL_AREA, infect_sub=True, **kw):
    return ob.infect_people(model, props, 111, infect_sub, corr, area_m2, n_sub, list(range(len(phases))), phases,
                            exposure, now, duration, **kw)


def test_area_javadoc_worked_example():
    """Covid19TransmissionArea.java:112-121: 1 infectious at its peak + 1 susceptible, 1 h, 10 m2, p_B = 0.51."""
    assert one_call(area(), 10.0, 1.0)["p"] == pytest.approx(1 - math.exp(-0.051), rel=1e-14)
    assert one_call(area(), 10.0, 1.0)["p"] == pytest.approx(0.0497, abs=5e-4)          # "0.05 (5%)"
    assert one_call(area(), 5.0, 1.0)["p"] == pytest.approx(0.0970, abs=5e-4)           # "9.6%"
    assert one_call(area(), 10.0, 2.0)["p"] == pytest.approx(0.0970, abs=5e-4)          # 2 hours
    assert one_call(area(beta=0.5), 10.0, 1.0)["p"] == pytest.approx(0.0252, abs=5e-4)  # masks, "2.5%"
    assert one_call(area(), 1000.0, 1.0)["p"] == pytest.approx(5.1e-4, rel=1e-2)        # "1E-3" order


def test_area_triangle_boundaries():
    tmin, tmode, tmax = 48.0, 3.4 * 24.0, 9.6 * 24.0
    p = lambda te: one_call(area(), 10.0, 1.0, now=te)["p"]
    assert math.isnan(p(np.nextafter(tmin, 0)))            # not yet contagious: sum == 0 -> early return (TArea:178)
    assert math.isnan(p(tmin))                             # contribution exactly 0 at te == t_e_min
    assert p(np.nextafter(tmin, 1e9)) == 0.0 and p(tmin + 1e-6) > 0   # sum > 0 but 1 - exp(-1e-17) rounds to 0
    assert p(tmode) == pytest.approx(1 - math.exp(-0.051), rel=1e-14)   # peak belongs to the falling branch (>=)
    assert math.isnan(p(tmax))                             # (max - te) == 0 at te == t_e_max, inclusive bound
    assert p(tmax - 1e-6) > 0 and math.isnan(p(np.nextafter(tmax, 1e9)))


def test_area_record_lists_and_thresholds():
    # every ILL occupant is listed as infectious, contagious or not (TArea:175)
    r = one_call(area(), 10.0, 1.0, phases=(E, IS, S, R, D, H), exposure=(80.0, 0.0, 0.0, 0.0, 0.0, -500.0))
    assert r["calculated"] and r["infectious"].tolist() == [0, 1, 5]
    # duration below the threshold: not calculated; exactly the threshold: calculated (TArea:138 uses <)
    thr = 60.0 / 3600.0
    assert not one_call(area(), 10.0, np.nextafter(thr, 0))["calculated"]
    assert one_call(area(), 10.0, thr)["calculated"]
    # nobody susceptible / nobody ill
    assert one_call(area(), 10.0, 1.0, phases=(IS, R))["infected"].tolist() == []
    assert math.isnan(one_call(area(), 10.0, 1.0, phases=(S, S))["p"])


def test_area_float_exposure_time_is_widened_not_rounded_twice():
    # exposure is a float (ConstructHerosModel.java:745): te = now - (double)(float) t
    t_exp = 1000.1
    f = float(np.float32(t_exp))
    now = f + 60.0
    want = 1 - math.exp(-0.51 * 1.0 / 10.0 * ((now - f) - 48.0) / (81.6 - 48.0))
    got = one_call(area(), 10.0, 1.0, exposure=(t_exp, 0.0), now=now)["p"]
    assert got == pytest.approx(want, rel=1e-13)
    wrong = 1 - math.exp(-0.51 * 1.0 / 10.0 * ((now - t_exp) - 48.0) / (81.6 - 48.0))
    assert abs(got - wrong) > 1e-9


def test_area_degenerate_locations_follow_ieee():
    """76 Helsinki Accommodation rows have nb_sublocations = 0 and 3 563 Workplace rows area = -1."""
    r = one_call(area(), 0.0, 1.0, n_sub=0)   # area = 0 * 0 -> 0/0 = NaN -> factor NaN -> p NaN -> nobody infected
    assert r["calculated"] and math.isnan(r["p"]) and len(r["infected"]) == 0
    r = one_call(area(pB=50.0), -1.0, 1.0)    # negative area: factor > 0 -> p = 1 - exp(+x) < 0
    assert r["p"] < 0 and len(r["infected"]) == 0


def test_area_total_location_branch_quirks():
    """infectSub = false and nSub >= 2 (TArea:198-246): no minus sign (TArea:205), te/(mode-min) and
    1 - te/(max-mode) (TArea:219,221).  p <= 0 while the sum is positive; p > 0 only when late cases make it negative."""
    ids = [0, 1, 2]
    r = ob.infect_people(ob.MODEL_AREA, area(), 111, False, 1.0, 20.0, 2, [0], [S], [0.0], 60.0, 1.0,
                         all_ids=ids, all_phase=[IS, S, S], all_exposure=[0.0, 0.0, 0.0])
    want_sum = 60.0 / (81.6 - 48.0)
    assert r["p"] == pytest.approx(1 - math.exp(0.51 / 20.0 * want_sum), rel=1e-13) and r["p"] < 0
    assert r["infectious"].tolist() == [0] and len(r["infected"]) == 0
    # te = 200 h > max - mode = 148.8 h: contribution 1 - 200/148.8 < 0 -> p > 0 (KAT with sum < 0)
    r = ob.infect_people(ob.MODEL_AREA, area(pB=30.0), 111, False, 1.0, 20.0, 2, [0], [S], [0.0], 200.0, 1.0,
                         all_ids=ids, all_phase=[IS, S, S], all_exposure=[0.0, 0.0, 0.0])
    s = 1.0 - 200.0 / (230.4 - 81.6)
    assert s < 0 and r["p"] == pytest.approx(1 - math.exp(30.0 / 20.0 * s), rel=1e-13) and r["p"] > 0
    # the draws go over ALL persons of the location (TArea:234), with u = Philox(seed, person, now)
    want = [i for i in (1, 2) if ob.u01(111, i, np.float64(200.0).view(np.uint64), 0) < r["p"]]
    assert r["infected"].tolist() == want


def dist_props(**kw):
    p = dict(L=2.0, I=3.4, C=6.2, v_max=7.23, v_0=4.0, r=2.294, psi=1.0, alpha=0.05, mu=0.0, thr=60.0)
    p.update(kw)
    return ob.DistProps(p["L"], p["I"], p["C"], p["v_max"], p["v_0"], p["r"], p["psi"], p["alpha"], p["mu"], p["thr"])


def dist_expected(area_sub, n, duration, t, psi=1.0, alpha=0.05, mu=0.0):
    L, I, C, vmax, v0, r = 48.0, 81.6, 6.2 * 24.0, 7.23, 4.0, 2.294
    delta = math.sqrt(area_sub / n)
    sigma = 1.0 - 1.0 / (1.0 + math.exp(-3.0 * (max(delta, psi) - 1.5)))
    factor = sigma * alpha * (1 - mu) * (1 - mu)
    if L <= t < I:
        v = vmax * ((t - L) / (I - L))
    elif I <= t < I + C:
        v = vmax * ((I + C - t) / C)
    else:
        v = 0.0
    if v <= 0:
        return float("nan")
    pt = 1 / (1 + math.exp(-r * (v - v0)))
    return 1 - math.exp(-(factor * duration * pt))


def test_distance_formula_and_boundaries():
    for t in (50.0, 81.6, 100.0, 200.0):
        got = one_call(dist_props(), 30.0, 2.0, model=ob.MODEL_DIST, now=t)["p"]
        assert got == pytest.approx(dist_expected(30.0, 2, 2.0, t), rel=1e-12)
    I_C = 81.6 + 6.2 * 24.0
    assert math.isnan(one_call(dist_props(), 30.0, 2.0, model=ob.MODEL_DIST, now=48.0)["p"])   # v_t == 0 at t == L
    assert math.isnan(one_call(dist_props(), 30.0, 2.0, model=ob.MODEL_DIST, now=I_C)["p"])    # exclusive upper end
    # only occupants with v_t > 0 are listed (TDist:161-166), unlike the area model
    r = one_call(dist_props(), 30.0, 2.0, model=ob.MODEL_DIST, phases=(E, IS, S), exposure=(90.0, 0.0, 0.0), now=100.0)
    assert r["infectious"].tolist() == [1]
    # masks and distancing (DiseasePolicy psi / mu)
    got = one_call(dist_props(psi=2.0, mu=0.6), 30.0, 2.0, model=ob.MODEL_DIST, now=100.0)["p"]
    assert got == pytest.approx(dist_expected(30.0, 2, 2.0, 100.0, psi=2.0, mu=0.6), rel=1e-12)


def test_distance_degenerate_locations_follow_ieee():
    r = one_call(dist_props(), -1.0, 2.0, model=ob.MODEL_DIST, now=100.0)  # sqrt(negative) = NaN -> max -> NaN
    assert r["calculated"] and math.isnan(r["p"]) and len(r["infected"]) == 0
    r = one_call(dist_props(), 0.0, 2.0, n_sub=0, model=ob.MODEL_DIST, now=100.0)
    assert r["calculated"] and len(r["infected"]) == 0


def test_distance_total_branch_uses_undivided_area_and_sublocation_head_count():
    ids = [0, 1, 2, 3]
    r = ob.infect_people(ob.MODEL_DIST, dist_props(), 111, False, 1.0, 60.0, 3, [0, 1], [S, S], [0.0, 0.0], 100.0, 2.0,
                         all_ids=ids, all_phase=[S, S, IS, IS], all_exposure=[0.0, 0.0, 0.0, 10.0])
    # Delta = sqrt(60 / 2): un-divided area, head count of the SUB-location (TDist:196)
    delta = math.sqrt(60.0 / 2)
    sigma = 1.0 - 1.0 / (1.0 + math.exp(-3.0 * (max(delta, 1.0) - 1.5)))
    factor = sigma * 0.05
    s = 0.0
    for te in (100.0, 90.0):
        v = 7.23 * ((81.6 + 148.8 - te) / 148.8) if te >= 81.6 else 7.23 * ((te - 48.0) / 33.6)
        s += factor * 2.0 * (1 / (1 + math.exp(-2.294 * (v - 4.0))))
    assert r["p"] == pytest.approx(1 - math.exp(-s), rel=1e-12)
    assert r["infectious"].tolist() == [2, 3]


def small_sim(model=ob.MODEL_AREA, trigger=ob.TRIGGER_EXIT_ONLY, n=4):
    props = load_props("area" if model == ob.MODEL_AREA else "distance")
    a, d, prog = oracle_props(props)
    sim = ob.Sim(model, trigger, 111)
    sim.set_area(a) if model == ob.MODEL_AREA else sim.set_dist(d)
    sim.set_prog(prog)
    sim.set_location_types([1], [1.0])
    sim.set_locations([0, 0], [1, 2], [10.0, 60.0])
    sim.set_persons([30] * n, [0] * n)
    return sim


def test_driver_duration_accumulates_until_threshold_and_exit_sees_the_leaver():
    sim = small_sim()
    sim.expose([0], [0.0])
    # person 0 infectious from 48 h; persons 1, 2 visit location 0
    sim.add_visits([0, 1, 2], [0, 0, 0], [0, 0, 0], [50.0, 50.0, 50.5], [60.0, 51.0, 51.0 + 30.0 / 3600.0])
    sim.run(50.9)
    assert sim.n_evaluations() == 0          # enters do not trigger in EXIT_ONLY mode
    sim.run(51.005)
    assert sim.n_evaluations() == 1 and sim.last_calc(0, 0) == 51.0   # exit of person 1, duration 51 - 0
    sim.run(52.0)
    assert sim.n_evaluations() == 1          # exit of person 2 only 30 s later: below the 60 s threshold
    sim.run(61.0)
    assert sim.n_evaluations() == 2 and sim.last_calc(0, 0) == 60.0


def test_driver_enter_and_exit_mode_and_ties():
    sim = small_sim(trigger=ob.TRIGGER_ENTER_AND_EXIT)
    sim.add_visits([0, 1, 2], [0, 0, 0], [0, 0, 0], [10.0, 10.0, 12.0], [12.0, 14.0, 14.0])
    sim.run(100.0)
    # distinct event times 10, 12, 14 -> one calculated evaluation each (tied events have duration 0)
    assert sim.n_evaluations() == 3


def test_driver_progression_matches_golden_timing_band():
    """Seeds exposed at t = 0 enter I(A)/I(S) between 60 h and 91.2 h (first I at 61.5 h, all by 90 h)."""
    sim = small_sim(n=400)
    sim.expose(np.arange(400), np.zeros(400))
    sim.run(59.99)
    assert sim.phase_counts()[E] == 400
    sim.run(91.21)
    c = sim.phase_counts()
    assert c[E] == 0 and c[IA] + c[IS] == 400 and 0.3 < c[IA] / 400 < 0.62   # FractionAsymptomatic = 0.46
    sim.run(24.0 * 60)
    c = sim.phase_counts()
    assert c.sum() == 400 and c[R] + c[D] + c[H] + c[ICU] + c[IA] + c[IS] == 400
    tr = sim.transitions()
    assert (np.diff(tr["time"]) >= 0).all()   # fired in time order
    # a re-opened stay is closed by re-submitting it with the same t_in: no ghost occupant
    sim2 = small_sim()
    sim2.add_visits([1], [0], [0], [5.0], [np.inf])
    sim2.add_visits([1], [0], [0], [5.0], [9.0])
    sim2.add_visits([2], [0], [0], [9.5], [11.0])
    sim2.run(20.0)
    assert sim2.n_evaluations() == 2


def test_incubation_period_against_the_reference_curves():
    """The only reference-produced sample of a sojourn time: in seed comparison.xlsx the 100 persons exposed at t = 0 are
    the only ones that can be infectious before 54.5 h + 60 h, so I(A) + I(S) over hours 48..96 is 100 x the empirical
    CDF of the incubation period -- Triangular(2.5, 3.4, 3.8) days drawn by DSOL's DistTriangular and converted by
    medlabs' DurationDistribution(TimeUnit.DAY), both outside the reference tree.  Checked here:
    (1) the reference's sample against the closed-form triangular CDF in hours, (2) the oracle's own incubation times
    (its inverse-CDF statement, keyed Philox uniforms) against the reference's sample, (3) the asymptomatic split of
    the 4 x 100 initial cases against covidP.FractionAsymptomatic = 0.46 and against the oracle's split.
    The four sheets carry the SAME infectious-count series (the duration draws of the reference do not depend on
    generic.Seed; only the split does), so the incubation sample has 100 members, not 400."""
    with open(os.path.join(GOLDEN, "seed_comparison.json")) as f:
        seeds = json.load(f)["seeds"]
    series = {k: v["infectious_48_96h"] for k, v in seeds.items()}
    t = 48.0 + 0.5 * np.arange(len(series["111"]["asymptomatic"]))
    total = {k: np.array(v["asymptomatic"]) + np.array(v["symptomatic"]) for k, v in series.items()}
    for k in total:
        assert (total[k] == total["111"]).all() and total[k][-1] == 100 and total[k][t < 60.0].max() == 0
    ecdf = total["111"] / 100.0
    assert (np.diff(ecdf) >= 0).all() and ecdf[t >= 91.5].min() == 1.0     # support [60 h, 91.2 h]

    lo, mode, hi = 2.5 * 24, 3.4 * 24, 3.8 * 24

    def tri_cdf(x):
        x = np.clip(x, lo, hi)
        return np.where(x < mode, (x - lo) ** 2 / ((hi - lo) * (mode - lo)), 1 - (hi - x) ** 2 / ((hi - lo) * (hi - mode)))

    ks_crit_100 = 1.36 / math.sqrt(100)                       # 5 % two-sided, n = 100
    # (1) a count sampled at t holds the transitions of (t - 0.5, t]: the model CDF lies between F(t - 0.5) and F(t)
    d1 = np.maximum(ecdf - tri_cdf(t), tri_cdf(t - 0.5) - ecdf).max()
    assert d1 < ks_crit_100, d1
    # ... and the sample does tell shapes apart: a uniform or a symmetric triangular on the same support is rejected
    uni = np.clip((t - lo) / (hi - lo), 0, 1)
    mid = 0.5 * (lo + hi)
    sym = np.where(t < mid, 2 * (np.clip(t, lo, hi) - lo) ** 2 / (hi - lo) ** 2, 1 - 2 * (hi - np.clip(t, lo, hi)) ** 2 / (hi - lo) ** 2)
    assert np.abs(ecdf - uni).max() > ks_crit_100 and np.abs(ecdf - sym).max() > ks_crit_100

    # (2) the oracle: 4 000 persons exposed at t = 0, their E -> I transition times
    n = 4000
    sim = small_sim(n=n)
    sim.expose(np.arange(n), np.zeros(n))
    sim.run(96.0)
    tr = sim.transitions()
    onset = tr["time"][(tr["phase"] == IA) | (tr["phase"] == IS)]
    assert len(onset) == n and onset.min() >= lo and onset.max() <= hi
    oracle_cdf = np.searchsorted(np.sort(onset), t, side="right") / n
    d2 = np.abs(oracle_cdf - ecdf).max()
    assert d2 < 1.36 * math.sqrt((100 + n) / (100.0 * n)), d2
    assert np.abs(oracle_cdf - tri_cdf(t)).max() < 1.36 / math.sqrt(n)      # and the oracle against the closed form

    # (3) 184 of the 400 initial cases of the four sheets are asymptomatic: 0.46 on the nose
    asym = sum(v["asymptomatic"][-1] for v in series.values())
    sigma = math.sqrt(0.46 * 0.54 / 400)
    assert abs(asym / 400.0 - 0.46) < 2 * sigma
    oracle_asym = float((tr["phase"] == IA).sum()) / n
    assert abs(oracle_asym - asym / 400.0) < 2 * math.sqrt(0.46 * 0.54 * (1 / 400 + 1 / n))


def test_compartment_series_from_the_transition_log():
    """DiseaseMonitor mirror (ConstructHerosModel.java:145): the 0.5 h compartment series rebuilt from the transition
    log equals the live phase counters at every day boundary, keeps the population constant, and has the column
    layout of the golden spreadsheets."""
    from common import Scenario, load_props, run_oracle
    from heros_b200 import monitor
    scn = Scenario(seed=3, n_persons=500, n_households=180, n_workplaces=20, n_shops=4, days=12)
    props = load_props("area")
    props["covidT_area.contagiousness"] = "5.0"
    ora = run_oracle(scn, props, ob.MODEL_AREA, ob.TRIGGER_EXIT_ONLY, seed=9)
    times, counts = monitor.compartment_series(ora["transitions"], scn.n_persons, 24.0 * scn.days, 0.5)
    assert len(times) == 2 * 24 * scn.days + 1 and times[1] == 0.5
    assert (counts.sum(axis=1) == scn.n_persons).all()
    assert counts[0, 0] == scn.n_persons - len(scn.seeds) and counts[0, 1] == len(scn.seeds)   # seeds exposed at t = 0
    np.testing.assert_array_equal(counts[48::48], ora["counts"])      # every midnight = the engine's own counters
    assert (np.diff(counts[:, 0]) <= 0).all()                          # nobody becomes susceptible again
    assert monitor.HEADER == ("Time(h)", "Susceptible", "Exposed", "Infected-Asymptomatic", "Infected-Symptomatic",
                              "Hospitalized", "ICU", "Dead", "Recovered")


def test_driver_contract_against_the_reference_epidemic_curves():
    """The only numbers the reference publishes for this path are the epidemic curves of seed comparison.xlsx
    (tests/golden/seed_comparison.json, made by tools/make_golden_curves.py).  They pin three observable consequences
    of the driver contract the oracle restates, in all four seeds:
      * nobody is exposed before the seeds become contagious (48 h); the first secondary exposure comes with the first
        morning movements after that (54.5 h in the reference);
      * the seeds turn infectious between 60 h and 91.2 h (Triangular(2.5, 3.4, 3.8) days in hours);
      * exposures only happen at movement events: none between 00:00 and 06:00 of any day (the half-hour-of-day
        histogram of new exposures is empty there), while the bin of midnight itself is not (activities that end at
        24:00).
    The same three properties must hold for the oracle driven by the schedule generator on a synthetic city."""
    import heros_b200 as hb
    from heros_b200 import scenario
    with open(os.path.join(GOLDEN, "seed_comparison.json")) as f:
        gold = json.load(f)
    assert tuple(gold["header"]) == ("Time(h)",) + hb.PHASE_NAMES   # column order = phase registration order
    for seed, g in gold["seeds"].items():
        assert g["daily_counts"][0] == [553567, 100, 0, 0, 0, 0, 0, 0]
        assert all(sum(row) == 553667 for row in g["daily_counts"])
        assert 48.0 < g["first_exposure_h"] <= 55.0
        assert 60.0 <= g["first_infectious_h"] <= 62.0 and g["all_seeds_infectious_by_h"] <= 91.5
        hist = np.array(g["new_exposures_by_half_hour_of_day"])
        assert hist[1:13].sum() == 0 and hist[0] > 0 and hist[13:44].sum() > 0.99 * hist.sum()

    city = scenario.City(5000, seed=11)
    props = load_props("distance")
    props["covidT_dist.alpha"] = "2.0"
    _, dist, prog = oracle_props(props)
    sim = ob.Sim(ob.MODEL_DIST, ob.TRIGGER_EXIT_ONLY, 111)
    sim.set_dist(dist)
    sim.set_prog(prog)
    sim.set_location_types(city.type_infect_sub, city.type_correction)
    sim.set_locations(city.loc_type, city.loc_nsub, city.loc_area)
    sim.set_persons(city.age, city.female)
    seeds = np.arange(0, 5000, 50, dtype=np.int32)
    sim.expose(seeds, np.zeros(len(seeds)))
    gen = scenario.ScheduleGenerator(city, 111)
    for d in range(12):
        st = gen.generate_day(sim.agent_state()["phase"])
        sim.add_visits(st["person"], st["loc"], st["sub"], st["t_in"], st["t_out"])
        sim.run(24.0 * (d + 1))
    inf, tr = sim.infections(), sim.transitions()
    assert len(inf["time"]) > 300
    assert 48.0 < inf["time"].min() < 60.0                      # first morning after the seeds became contagious
    became_infectious = tr["time"][np.isin(tr["person"], seeds) & np.isin(tr["phase"], (2, 3))]
    assert len(became_infectious) == len(seeds) and 60.0 <= became_infectious.min() and became_infectious.max() <= 91.2
    hour = inf["time"] % 24.0
    assert not ((hour > 0.0) & (hour <= 6.0)).any()             # nothing happens while everybody sleeps


def test_sojourn_supports_against_the_reference_curves():
    """First appearance of Hospitalized / ICU / Dead / Recovered in the four reference curves against the lower ends of
    the sojourn-time supports (alpha parameters, days -> hours): nobody can be in hospital before incubation 2.5 d +
    PeriodSymptomaticToHospitalized 7 d = 228 h, in the ICU before + 1 d, dead before + 2 d more (deaths only through
    the ICU: FractionHospitalizedToDead = 0), recovered before 2.5 d + 12 d = 348 h -- and with about a hundred initial
    cases the first ones come within a few days of those bounds, which a wrong unit (hours for days) or a wrong
    transition table would not survive.  The oracle's own first times obey the same bounds."""
    with open(os.path.join(GOLDEN, "seed_comparison.json")) as f:
        seeds = json.load(f)["seeds"]
    lower = {"Hospitalized": 24 * (2.5 + 7), "ICU": 24 * (2.5 + 7 + 1), "Dead": 24 * (2.5 + 7 + 1 + 2), "Recovered": 24 * (2.5 + 12)}
    for g in seeds.values():
        first = g["first_time_positive_h"]
        for name, bound in lower.items():
            assert first[name] >= bound, (name, first[name])
        assert first["Hospitalized"] < lower["Hospitalized"] + 72 and first["Recovered"] < lower["Recovered"] + 48
        assert first["Hospitalized"] < first["ICU"] < first["Dead"]
    n = 20000
    sim = small_sim(n=n)
    sim.set_persons([70] * n, [0] * n)          # an age band with hospital, ICU and death fractions > 0
    sim.expose(np.arange(n), np.zeros(n))
    sim.run(24.0 * 40)
    tr = sim.transitions()
    for name, phase in (("Hospitalized", H), ("ICU", ICU), ("Dead", D), ("Recovered", R)):
        times = tr["time"][tr["phase"] == phase]
        slack = {"Hospitalized": 48, "ICU": 72, "Dead": 120, "Recovered": 48}[name]   # a minimum of sums of 2..4 triangulars
        assert len(times) > 0 and lower[name] <= times.min() < lower[name] + slack, (name, times.min())


def test_policy_runs_coincide_until_the_policy_as_in_the_reference():
    """out-flip-lockdown-no-10-20-comparison.xlsx (tests/golden/policy_comparison.json): seed 111 without policy, with the
    lockdown from day 10 and from day 20.  The three runs are identical sample for sample until the policy and differ
    from the 06:30 sample of the policy's day on (246.5 h / 486.5 h): a replication is a pure function of its seed, a
    policy event of midnight acts on that day's activities, and its first consequence is the first morning exposure.
    The oracle driven by the schedule generator must show the same: equal infection lists before the policy's day,
    a different list on it, no difference before the morning."""
    from heros_b200 import scenario
    with open(os.path.join(GOLDEN, "policy_comparison.json")) as f:
        gold = json.load(f)
    first = gold["first_differing_sample_h"]
    assert first == {"Day 10 vs Nothing": 246.5, "Day 20 vs Nothing": 486.5, "Day 10 vs Day 20": 246.5}
    assert gold["nothing_equals_seed_111_of_seed_comparison_until_h"] == 486.5    # that workbook's seed 111 locks down on day 20
    for day, h in ((10, first["Day 10 vs Nothing"]), (20, first["Day 20 vs Nothing"])):
        assert 24.0 * day + 6.0 <= h <= 24.0 * day + 7.0
    sus = gold["susceptible_at_day"]
    assert sus["Day 10"]["10"] == sus["Day 20"]["10"] == sus["Nothing"]["10"] and sus["Day 20"]["20"] == sus["Nothing"]["20"]
    assert sus["Day 10"]["120"] > sus["Day 20"]["120"] > sus["Nothing"]["120"]   # the earlier the lockdown, the fewer infected

    city = scenario.City(5000, seed=11)
    props = load_props("distance")
    props["covidT_dist.alpha"] = "2.0"
    _, dist, prog = oracle_props(props)
    policy_day, days = 5, 8

    def run(lockdown):
        sim = ob.Sim(ob.MODEL_DIST, ob.TRIGGER_EXIT_ONLY, 111)
        sim.set_dist(dist)
        sim.set_prog(prog)
        sim.set_location_types(city.type_infect_sub, city.type_correction)
        sim.set_locations(city.loc_type, city.loc_nsub, city.loc_area)
        sim.set_persons(city.age, city.female)
        seeds = np.arange(0, 5000, 50, dtype=np.int32)
        sim.expose(seeds, np.zeros(len(seeds)))
        gen = scenario.ScheduleGenerator(city, 111)
        for d in range(days):
            if lockdown and d == policy_day:          # the policy event of midnight comes before the day's movements
                for name in ("Workplace", "Retail", "BarRestaurant", "PrimarySchool", "SecondarySchool", "Recreation"):
                    gen.set_closure_policy(scenario.TYPE_ID[name], 0.1, 0.1, scenario.TYPE_ID["Accommodation"])
            st = gen.generate_day(sim.agent_state()["phase"])
            sim.add_visits(st["person"], st["loc"], st["sub"], st["t_in"], st["t_out"])
            sim.run(24.0 * (d + 1))
        inf = sim.infections()
        return np.stack([inf["time"], inf["person"].astype(np.float64), inf["loc"].astype(np.float64)], axis=1)

    free, locked = run(False), run(True)
    t0 = 24.0 * policy_day
    before_free, before_locked = free[free[:, 0] < t0], locked[locked[:, 0] < t0]
    assert len(before_free) > 50
    np.testing.assert_array_equal(before_free, before_locked)
    day_free = free[(free[:, 0] >= t0) & (free[:, 0] < t0 + 24.0)]
    day_locked = locked[(locked[:, 0] >= t0) & (locked[:, 0] < t0 + 24.0)]
    assert len(day_free) > 0
    only_one = set(map(tuple, day_free)) ^ set(map(tuple, day_locked))      # events of the day that only one run has
    first_diff = min(row[0] for row in only_one)
    assert t0 + 5.0 < first_diff < t0 + 24.0                   # nothing differs while everybody sleeps
    assert len(locked) < len(free)                             # and the lockdown slows the epidemic down


_CITY_RUNS = {}


def synthetic_city_exposure_times(model, lockdown_day=None):
    return synthetic_city_infections(model, lockdown_day)["time"]


SYNTHETIC_CITY = dict(n_persons=8000, city_seed=11, seed=111, days=16, first_cases=np.arange(0, 8000, 80, dtype=np.int32))


def synthetic_city_lockdown_rows():
    """The day-10 rows of the reference's HardLockdown_10-40.csv (tests/golden/location_policies.json)."""
    with open(os.path.join(GOLDEN, "location_policies.json")) as f:
        return [r for r in json.load(f)["files"]["HardLockdown_10-40"]["rows"] if float(r[0]) == 10.0]


def synthetic_city_infections(model, lockdown_day=None):
    """Infection events of 16 days of the oracle driven by the schedule generator on an 8 000-person synthetic city with 100
    initial cases (contagiousness / alpha = 1), optionally with the day-10 rows of the reference's HardLockdown_10-40.csv
    applied at midnight of ``lockdown_day``; kept for the tests that compare profiles with the reference curves."""
    if (model, lockdown_day) not in _CITY_RUNS:
        from heros_b200 import scenario
        city = scenario.City(8000, seed=11)
        props = load_props(model)
        props["covidT_area.contagiousness" if model == "area" else "covidT_dist.alpha"] = "1.0"
        area, dist, prog = oracle_props(props)
        sim = ob.Sim(ob.MODEL_AREA if model == "area" else ob.MODEL_DIST, ob.TRIGGER_EXIT_ONLY, 111)
        sim.set_area(area) if model == "area" else sim.set_dist(dist)
        sim.set_prog(prog)
        sim.set_location_types(city.type_infect_sub, city.type_correction)
        sim.set_locations(city.loc_type, city.loc_nsub, city.loc_area)
        sim.set_persons(city.age, city.female)
        first = np.arange(0, 8000, 80, dtype=np.int32)
        sim.expose(first, np.zeros(len(first)))
        gen = scenario.ScheduleGenerator(city, 111)
        lockdown = synthetic_city_lockdown_rows()
        for d in range(16):
            if d == lockdown_day:
                for _, name, fraction_open, fraction_activities, alternative, _ in lockdown:
                    if name in scenario.TYPE_ID:
                        gen.set_closure_policy(scenario.TYPE_ID[name], float(fraction_open), float(fraction_activities),
                                               scenario.TYPE_ID[alternative])
            st = gen.generate_day(sim.agent_state()["phase"])
            sim.add_visits(st["person"], st["loc"], st["sub"], st["t_in"], st["t_out"])
            sim.run(24.0 * (d + 1))
        _CITY_RUNS[(model, lockdown_day)] = {k: v.copy() for k, v in sim.infections().items()}
    return _CITY_RUNS[(model, lockdown_day)]


def test_lockdown_effect_has_the_size_the_reference_shows():
    """LocationType.setClosurePolicy is inside medlabs; the convention of DESIGN.md 3.4 stands in for it.  Its effect can
    be compared with the reference's: with HardLockdown_10-40 the reference counts 20 % of the no-policy exposures over
    days 10..15 (seed 111; 92 and 24 on days 10 and 11 against 566 and 835); the oracle driven by the schedule generator
    with the same policy rows applied on day 10 counts 32 % on a synthetic city (62 and 23 on days 10 and 11 against
    197 and 162), where households -- which a lockdown does not close -- weigh more (see the morning share above).
    Same direction, same order of magnitude; before day 10 the two runs are one."""
    with open(os.path.join(GOLDEN, "policy_comparison.json")) as f:
        daily = json.load(f)["susceptible_daily"]
    ref_free, ref_locked = -np.diff(daily["Nothing"])[:16], -np.diff(daily["Day 10"])[:16]
    assert (ref_free[:10] == ref_locked[:10]).all()
    ref_ratio = ref_locked[10:].sum() / ref_free[10:].sum()
    assert 0.1 < ref_ratio < 0.3
    free = np.bincount((synthetic_city_exposure_times("area") // 24.0).astype(int), minlength=16)
    locked = np.bincount((synthetic_city_exposure_times("area", lockdown_day=10) // 24.0).astype(int), minlength=16)
    assert (free[:10] == locked[:10]).all() and free[10:].sum() > 500
    ratio = locked[10:].sum() / free[10:].sum()
    assert 0.1 < ratio < 0.5, ratio
    assert locked[11] < 0.3 * free[11] and ref_locked[11] < 0.3 * ref_free[11]      # the second day of the lockdown


def test_morning_exposures_tell_which_model_made_the_reference_curves():
    """The reference does not record which transmission model produced seed comparison.xlsx.  Its hour-of-day profile of
    new exposures does: 7.3 % of them fall into the two half-hours after 06:00, when the first people leave home and the
    household sub-location is evaluated over the whole night.  With the oracle and the schedule generator on a synthetic
    city the area model puts a comparable or larger share there (households of 100 m2: p = 1 - exp(-p_B t / A) over
    ten hours), the distance model next to nothing (the distance factor sigma(sqrt(A / N)) of a home is ~1e-6) -- so
    the published curves are area-model runs, which is the model the seed-envelope comparison of DESIGN.md uses."""
    with open(os.path.join(GOLDEN, "seed_comparison.json")) as f:
        seeds = json.load(f)["seeds"]
    ref = np.sum([np.array(v["new_exposures_by_half_hour_of_day"]) for v in seeds.values()], axis=0)
    ref_share = ref[13:15].sum() / ref.sum()            # samples 06:30 and 07:00 = exposures of (06:00, 07:00]
    assert 0.05 < ref_share < 0.10
    share = {}
    for model in ("area", "distance"):
        t = synthetic_city_exposure_times(model)
        assert len(t) > 1000
        hist = np.bincount(np.ceil((t % 24.0) * 2).astype(int) % 48, minlength=48)   # the sample an exposure shows up in
        share[model] = hist[13:15].sum() / hist.sum()
        assert hist[1:13].sum() == 0
    assert share["area"] > ref_share > 10 * share["distance"], share
    # and the first secondary exposure of an area-model run shows in the same sample as in all four reference curves:
    # 54.5 h = the first departures from home (06:00-06:30) of the first morning after the 48 h latent period
    assert {g["first_exposure_h"] for g in seeds.values()} == {54.5}
    assert math.ceil(synthetic_city_exposure_times("area").min() * 2) / 2 == 54.5


def test_sundays_show_in_the_exposures_as_in_the_reference():
    """Day 0 of a run is a Monday and days 6, 13, ... take the Sunday week pattern (ConstructHerosModel.java:931-940:
    Monday..Thursday share the Monday pattern, Friday / Saturday / Sunday have their own).  In all four reference curves
    the new exposures of days 6 and 13 fall to about a third of the geometric mean of their neighbouring days, while the
    epidemic grows tenfold a week; the oracle driven by the schedule generator shows the same dips on the same days."""
    with open(os.path.join(GOLDEN, "seed_comparison.json")) as f:
        seeds = json.load(f)["seeds"]

    def dip(new, d):
        return new[d] / math.sqrt(new[d - 1] * new[d + 1])

    for g in seeds.values():
        new = -np.diff(np.array(g["daily_counts"])[:, 0]).astype(float)     # exposures of day d = S(d) - S(d + 1)
        assert new[0] == 0 and new[1] == 0                                   # nobody is contagious in the first 48 h
        assert dip(new, 6) < 0.5 and dip(new, 13) < 0.5
        assert sorted(range(3, 19), key=lambda d: dip(new, d))[:2] in ([6, 13], [13, 6])   # the two deepest dips of 16 days
    t = synthetic_city_exposure_times("area")
    new = np.bincount((t // 24.0).astype(int), minlength=16).astype(float)
    assert new[0] == 0 and new[1] == 0
    assert dip(new, 6) < 0.7 and dip(new, 13) < 0.7
    assert sorted(range(3, 15), key=lambda d: dip(new, d))[:2] in ([6, 13], [13, 6])


def test_progression_follows_the_reference_transition_table():
    """Covid19Progression.changeDiseasePhase (java :208-346) with the alpha parameters: only the transitions of the
    reference's table occur, every sojourn time lies inside the support of its Triangular distribution (days -> hours),
    and the branch frequencies match covidP.Fraction* of the person's age band (alpha-area.properties:47-108)."""
    props = load_props("area")
    _, _, prog = oracle_props(props)
    n_per_age = 6000
    ages = np.repeat([10, 35, 65, 75, 85], n_per_age).astype(np.int8)
    n = len(ages)
    sim = ob.Sim(ob.MODEL_AREA, ob.TRIGGER_EXIT_ONLY, 2024)
    sim.set_area(oracle_props(props)[0])
    sim.set_prog(prog)
    sim.set_location_types([1], [1.0])
    sim.set_locations([0], [1], [10.0])
    sim.set_persons(ages, (np.arange(n) % 2).astype(np.uint8))
    sim.expose(np.arange(n), np.zeros(n))
    sim.run(24.0 * 120)
    tr = sim.transitions()
    order = np.lexsort((tr["time"], tr["person"]))
    person, phase, time = tr["person"][order], tr["phase"][order], tr["time"][order]
    same = person[1:] == person[:-1]
    frm, to, dt = phase[:-1][same], phase[1:][same], (time[1:] - time[:-1])[same]
    who = person[1:][same]
    allowed = {(E, IA): (2.5, 3.8), (E, IS): (2.5, 3.8), (IA, R): (12, 20), (IS, H): (7, 11), (IS, R): (12, 20),
               (H, ICU): (1, 5), (H, D): (1, 5), (H, R): (11, 15), (ICU, D): (2, 6), (ICU, R): (28, 32)}
    seen = set(zip(frm.tolist(), to.tolist()))
    assert seen <= set(allowed), seen - set(allowed)
    assert (H, D) not in seen                                   # FractionHospitalizedToDead = 0.0
    for (a, b), (lo, hi) in allowed.items():
        m = (frm == a) & (to == b)
        if m.any():
            assert dt[m].min() >= 24.0 * lo - 1e-9 and dt[m].max() <= 24.0 * hi + 1e-9, (a, b, dt[m].min(), dt[m].max())
    first = phase[np.r_[True, ~same]]
    assert (first == E).all() and len(first) == n               # everybody's trajectory starts with Exposed at t = 0
    # everybody has reached an absorbing phase after 120 days
    c = sim.phase_counts()
    assert c[R] + c[D] == n

    def frac(a, b, sel):
        src = (frm == a) & sel[who]
        return ((to == b) & src).sum() / max(src.sum(), 1), src.sum()

    everyone = np.ones(n, bool)
    f, k = frac(E, IA, everyone)
    assert abs(f - 0.46) < 4 * np.sqrt(0.46 * 0.54 / k)        # FractionAsymptomatic
    want_s2h = {10: 0.02153, 35: 0.05044, 65: 0.44038, 75: 0.60867, 85: 0.32301}
    want_h2icu = {10: 0.00152, 35: 0.00921, 65: 0.14674, 75: 0.15508, 85: 0.01647}
    want_icu2d = {10: 0.0, 35: 0.0, 65: 0.08393, 75: 0.39731, 85: 0.63002}
    for age in (10, 35, 65, 75, 85):
        sel = ages == age
        for (a, b, want) in ((IS, H, want_s2h[age]), (H, ICU, want_h2icu[age]), (ICU, D, want_icu2d[age])):
            f, k = frac(a, b, sel)
            if k >= 30:
                assert abs(f - want) < 4.5 * np.sqrt(max(want * (1 - want), 1e-4) / k) + 1e-9, (age, a, b, f, want, k)
    # the golden curves show the same split: on day 10 of seed 111 I(A) : I(S) = 239 : 266 (0.473 asymptomatic)
    with open(os.path.join(GOLDEN, "seed_comparison.json")) as fjson:
        day10 = json.load(fjson)["seeds"]["111"]["daily_counts"][10]
    assert abs(day10[2] / (day10[2] + day10[3]) - 0.46) < 0.05


def test_single_calls_match_an_independent_restatement_of_the_formulas():
    """A second, independent statement of the two infectPeople formulas (sub-location branch), written here in plain
    Python from the Java text (Covid19TransmissionArea.java:138-196, Covid19TransmissionDistance.java:119-187), against
    the C oracle over random occupant sets, areas, durations and times: calculated flag, infectious list, pInfection
    (python's math.exp is libm, the oracle's exp is fdlibm: <= 1 ulp apart, hence the tolerance) and the infected list
    given the oracle's own keyed draws."""
    rng = np.random.default_rng(99)
    ill = {E, IA, IS, H, ICU}
    a_props = dict(pB=0.8, beta=0.9, tmin=48.0, tmode=81.6, tmax=230.4, thr=60.0 / 3600.0)
    d_props = dict(L=48.0, I=81.6, C=148.8, vmax=7.23, v0=4.0, r=2.294, psi=1.0, alpha=0.05, mu=0.1, thr=60.0 / 3600.0)

    def area_ref(phases, exposure, total_area, nsub, corr, duration, now):
        if duration < a_props["thr"]:
            return False, [], math.nan
        a = float(np.float32(total_area)) / nsub
        factor = -a_props["beta"] * a_props["pB"] * duration / (corr * a)
        if factor == 0.0:
            return True, [], math.nan
        s, infectious = 0.0, []
        for i, (ph, ex) in enumerate(zip(phases, exposure)):
            if ph in ill:
                te = now - float(np.float32(ex))
                c = 0.0
                if a_props["tmin"] <= te < a_props["tmode"]:
                    c = (te - a_props["tmin"]) / (a_props["tmode"] - a_props["tmin"])
                elif a_props["tmode"] <= te <= a_props["tmax"]:
                    c = (a_props["tmax"] - te) / (a_props["tmax"] - a_props["tmode"])
                s += c
                infectious.append(i)
        if s == 0.0:
            return True, infectious, math.nan
        return True, infectious, 1.0 - math.exp(factor * s)

    def dist_ref(phases, exposure, total_area, nsub, duration, now):
        if duration < d_props["thr"]:
            return False, [], math.nan
        a = float(np.float32(total_area)) / nsub
        delta = math.sqrt(a / len(phases))
        sigma = 1.0 - 1.0 / (1.0 + math.exp(-3.0 * (max(delta, d_props["psi"]) - 1.5)))
        factor = sigma * d_props["alpha"] * (1.0 - d_props["mu"]) * (1.0 - d_props["mu"])
        if factor == 0.0:
            return True, [], math.nan
        s, infectious = 0.0, []
        for i, (ph, ex) in enumerate(zip(phases, exposure)):
            if ph in ill:
                t = now - float(np.float32(ex))
                v = 0.0
                if d_props["L"] <= t < d_props["I"]:
                    v = d_props["vmax"] * ((t - d_props["L"]) / (d_props["I"] - d_props["L"]))
                elif d_props["I"] <= t < d_props["I"] + d_props["C"]:
                    v = d_props["vmax"] * ((d_props["I"] + d_props["C"] - t) / d_props["C"])
                if v > 0:
                    s += factor * duration * (1 / (1 + math.exp(-d_props["r"] * (v - d_props["v0"]))))
                    infectious.append(i)
        if s == 0.0:
            return True, infectious, math.nan
        return True, infectious, 1.0 - math.exp(-s)

    o_area = ob.AreaProps(0.8, 0.9, 2.0, 3.4, 9.6, 60.0)
    o_dist = ob.DistProps(2.0, 3.4, 6.2, 7.23, 4.0, 2.294, 1.0, 0.05, 0.1, 60.0)
    n_p = 0
    for trial in range(400):
        n = int(rng.integers(1, 40))
        phases = rng.choice([S, S, S, E, IA, IS, H, ICU, D, R], n).tolist()
        now = float(rng.uniform(0.0, 400.0))
        exposure = np.where(rng.random(n) < 0.2, 0.0, rng.uniform(max(now - 300.0, 0.0), now + 1e-9, n)).astype(np.float32)
        nsub = int(rng.integers(1, 6))
        total_area = float(np.float32(rng.uniform(5.0, 400.0) * nsub))
        corr = float(rng.choice([1.0, 0.8, 1.25]))
        duration = float(rng.choice([rng.uniform(0.0, 0.03), rng.uniform(0.02, 12.0)]))
        ids = list(range(n))
        for model in (ob.MODEL_AREA, ob.MODEL_DIST):
            if model == ob.MODEL_AREA:
                got = ob.infect_people(model, o_area, 7, True, corr, total_area, nsub, ids, phases, exposure, now, duration)
                calc, infectious, p = area_ref(phases, exposure, total_area, nsub, corr, duration, now)
            else:
                got = ob.infect_people(model, o_dist, 7, True, 1.0, total_area, nsub, ids, phases, exposure, now, duration)
                calc, infectious, p = dist_ref(phases, exposure, total_area, nsub, duration, now)
            assert got["calculated"] == calc, (trial, model)
            assert got["infectious"].tolist() == infectious, (trial, model)
            if math.isnan(p):
                assert math.isnan(got["p"]) and len(got["infected"]) == 0, (trial, model, got["p"])
                continue
            n_p += 1
            assert got["p"] == pytest.approx(p, rel=1e-12, abs=4.5e-16), (trial, model)
            bits = int(np.float64(now).view(np.uint64))
            want_infected = [i for i in ids if phases[i] == S and ob.u01(7, i, bits, 0) < got["p"]]
            assert got["infected"].tolist() == want_infected, (trial, model)
    assert n_p > 200


def brute_force_run(model, trigger, scn, seed=4711):
    """(events of the brute-force statement, infections of the oracle) for one scenario"""
    props = load_props(model)
    props["covidT_area.contagiousness" if model == "area" else "covidT_dist.alpha"] = "25.0" if model == "area" else "3.0"
    a, d, prog = oracle_props(props)
    kind = ob.MODEL_AREA if model == "area" else ob.MODEL_DIST
    sim = ob.Sim(kind, ob.TRIGGER_EXIT_ONLY if trigger == "exit" else ob.TRIGGER_ENTER_AND_EXIT, seed)
    sim.set_area(a) if model == "area" else sim.set_dist(d)
    sim.set_prog(prog)
    sim.set_location_types(scn.type_infect_sub, scn.type_corr)
    sim.set_locations(scn.loc_type, scn.loc_nsub, scn.loc_area)
    sim.set_persons(scn.age, scn.female)
    sim.expose(scn.seeds, np.zeros(len(scn.seeds)))
    sim.add_visits(scn.person, scn.loc, scn.sub, scn.t_in, scn.t_out)
    horizon = 24.0 * scn.days
    sim.run(horizon)
    got = sim.infections()
    order = np.lexsort((got["person"], got["time"]))
    got = {k: v[order] for k, v in got.items()}

    # ---- the brute-force statement: one movement at a time, occupancy kept as plain sets ----
    pB, beta, tmin, tmode, tmax, thr = 25.0, 1.0, 48.0, 81.6, 230.4, 60.0 / 3600.0
    L, I, Cc, vmax, v0, r, psi, alpha, mu = 48.0, 81.6, 148.8, 7.23, 4.0, 2.294, 1.0, 3.0, 0.0
    susceptible = np.ones(scn.n_persons, bool)
    susceptible[scn.seeds] = False
    exposure = np.zeros(scn.n_persons, np.float32)
    is_ill = ~susceptible                       # the seeds stay ill for the whole run (E -> I(A) / I(S), never out in 4 days)
    key = scn.loc.astype(np.int64) * 1000 + scn.sub
    occupants, last_calc, events = {}, {}, []
    moves = [(float(scn.t_out[i]), 0, int(key[i]), int(scn.person[i]), i) for i in range(scn.n_visits)
             if scn.t_out[i] > scn.t_in[i] and np.isfinite(scn.t_out[i]) and scn.t_out[i] < horizon]
    moves += [(float(scn.t_in[i]), 1, int(key[i]), int(scn.person[i]), i) for i in range(scn.n_visits)
              if scn.t_out[i] > scn.t_in[i] and scn.t_in[i] < horizon]

    def contagion(persons, t, duration, factor):
        total = 0.0
        for p in sorted(persons):               # the oracle sums in person order (Trove's order is not reproducible)
            if not is_ill[p]:
                continue
            te = t - float(exposure[p])
            if model == "area":
                if tmin <= te < tmode:
                    total += (te - tmin) / (tmode - tmin) if factor < 0 else te / (tmode - tmin)
                elif tmode <= te <= tmax:
                    total += (tmax - te) / (tmax - tmode) if factor < 0 else 1.0 - te / (tmax - tmode)
            else:
                v = 0.0
                if L <= te < I:
                    v = vmax * ((te - L) / (I - L))
                elif I <= te < I + Cc:
                    v = vmax * ((I + Cc - te) / Cc)
                if v > 0:
                    total += factor * duration * (1 / (1 + math.exp(-r * (v - v0))))
        return total

    for t, enter, k, who, i in sorted(moves):   # exits before enters at equal times, then by sub-location, then by person
        here = occupants.setdefault(k, set())
        if not enter or trigger == "both":
            duration = t - last_calc.get(k, 0.0)
            if duration >= thr:                 # else not calculated: the interval keeps accumulating (C6)
                last_calc[k] = t
                loc = int(scn.loc[i])
                nsub, total_area = int(scn.loc_nsub[loc]), float(np.float32(scn.loc_area[loc]))
                corr = float(scn.type_corr[scn.loc_type[loc]])
                whole = not scn.type_infect_sub[scn.loc_type[loc]] and nsub >= 2      # the total-location branch
                group = set(here)               # a leaver is still counted, an enterer not yet (C5)
                if whole:
                    group = set().union(*(occupants.get(loc * 1000 + z, set()) for z in range(nsub)))
                p_inf = None
                if model == "area":
                    factor = (beta * pB * duration / (corr * total_area)) if whole else (-beta * pB * duration / (corr * (total_area / nsub)))
                    total = contagion(group, t, duration, factor)
                    if factor != 0.0 and total != 0.0:
                        p_inf = 1.0 - math.exp(factor * total)
                elif len(here) > 0:             # sqrt(area / 0) = inf with nobody in the sub-location: factor 0
                    delta = math.sqrt((total_area if whole else total_area / nsub) / len(here))
                    factor = (1.0 - 1.0 / (1.0 + math.exp(-3.0 * (max(delta, psi) - 1.5)))) * alpha * (1.0 - mu) * (1.0 - mu)
                    total = contagion(group, t, duration, factor)
                    if factor != 0.0 and total != 0.0:
                        p_inf = 1.0 - math.exp(-total)
                if p_inf is not None:
                    bits = int(np.float64(t).view(np.uint64))
                    for p in sorted(group):
                        if susceptible[p] and ob.u01(seed, p, bits, 0) < p_inf:
                            susceptible[p] = False
                            is_ill[p] = True
                            exposure[p] = np.float32(t)     # exposure time is a float (C4)
                            # the event is reported with the sub-location of the evaluation (the InfectionRecord's
                            # location); in the total-location branch the person may stand in another zone
                            events.append((t, p, loc, int(scn.sub[i]), p_inf))
        if enter:
            here.add(who)
        else:
            here.discard(who)
    events.sort()
    return events, got


@pytest.mark.parametrize("model,trigger", [("area", "exit"), ("distance", "exit"), ("area", "both"), ("distance", "both")])
def test_driver_matches_a_brute_force_statement_of_the_contract(model, trigger):
    """The oracle's event-by-event driver against a second statement of the driver contract
    written here as a plain loop over the movements in time order with the occupancy kept as sets (exits before enters
    at equal times; enters trigger only with ENTER_AND_EXIT): per sub-location lastCalculation with the threshold rule,
    the leaver still counted and the enterer not yet, the formula of the model -- sub-location and total-location
    branch, the parks of the scenario have infectInSublocation = false and three zones --, keyed draws, a person is
    infected once.  Four days with the seeds
    exposed at t = 0: nobody infected later turns contagious inside the run (48 h latent period from an exposure
    >= 48 h), so the comparison does not involve the progression draws."""
    from common import Scenario
    scn = Scenario(seed=21, n_persons=160, n_households=60, n_workplaces=8, n_shops=3, days=4, tie_fraction=0.25,
                   short_gap_fraction=0.15, n_parks=2, park_nsub=3)
    events, got = brute_force_run(model, trigger, scn)
    assert len(events) > 10 and len(events) == len(got["person"])
    assert [e[1] for e in events] == got["person"].tolist()
    assert [e[0] for e in events] == got["time"].tolist()
    assert [e[2] for e in events] == got["loc"].tolist() and [e[3] for e in events] == got["sub"].tolist()
    np.testing.assert_allclose([e[4] for e in events], got["p"], rtol=1e-12, atol=4.5e-16)
