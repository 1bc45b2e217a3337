"""Capacity-constrained locators (NearestLocatorCap / RandomLocatorCap and their Choice forms,
ConstructHerosModel.java:998-1020; capConstrained / capPersonsPerM2 / sizeFactor of locationtypes.csv).  The rule lives in
medlabs; the convention here (scenario.capacity_diversion): a constrained place whose unconstrained peak occupancy exceeds
its capacity sheds the share 1 - capacity / peak of the visits a *Cap locator sends to it to the next candidate."""
import numpy as np

from heros_b200 import loaders, scenario


def peak_occupancy(st, n_loc):
    loc = st["loc"].astype(np.int64)
    t_out = np.where(np.isinf(st["t_out"]), 1e9, st["t_out"])
    ev_loc, ev_t = np.r_[loc, loc], np.r_[st["t_in"], t_out]
    ev_d = np.r_[np.ones(len(loc), np.int64), -np.ones(len(loc), np.int64)]
    order = np.lexsort((ev_d, ev_t, ev_loc))
    run = np.cumsum(ev_d[order])
    peak = np.zeros(n_loc)
    np.maximum.at(peak, ev_loc[order], run)
    return peak


def test_the_shipped_schedule_uses_capacity_constrained_locators():
    pat = scenario.load_patterns()
    capped = [a for rows in pat["patterns"].values() for a in rows if a["locator"] & loaders.LOCATOR_CAP]
    plain = [a for rows in pat["patterns"].values() for a in rows if a["locator"] in (3, 4)]
    assert len(capped) > 200 and {a["locator"] & 0xF for a in capped} == {3, 4} and len(plain) > 0
    assert list(scenario.TYPE_CAP_CONSTRAINED[[scenario.TYPE_ID[n] for n in ("Accommodation", "Workplace", "Supermarket", "Park")]]) == [0, 0, 1, 1]
    assert scenario.TYPE_SIZE_FACTOR[scenario.TYPE_ID["Supermarket"]] == 9.7


def test_over_subscribed_places_shed_their_share_to_the_next_candidate():
    city = scenario.City(20000, seed=4, box_m=(2300.0, 1700.0))   # Hague density: several candidates within 2.5 km
    # the busiest supermarket loses nine tenths of its floor: it cannot hold its unconstrained peak, its neighbours can
    # take what it sheds
    constrained = scenario.TYPE_CAP_CONSTRAINED[city.loc_type] == 1
    markets = np.nonzero(city.loc_type == scenario.TYPE_ID["Supermarket"])[0]
    assert len(markets) >= 4
    free = scenario.ScheduleGenerator(city, 111).generate_day(np.zeros(city.n_persons, np.uint8))
    small = markets[[np.argmax(np.bincount(free["loc"], minlength=city.n_locations)[markets])]]
    city.loc_area[small] *= np.float32(0.1)
    share = scenario.capacity_diversion(city, 111)
    cap, peak0 = scenario.location_capacity(city), peak_occupancy(free, city.n_locations)
    np.testing.assert_array_equal(city.peak_unconstrained[constrained], peak0[constrained])
    over = peak0 > cap
    assert over[small].all() and not over[~constrained].any()
    np.testing.assert_allclose(share[over], 1.0 - cap[over] / peak0[over])
    assert (share[~over] == 0).all() and (share[small] > 0.15).all()
    # with the shares in force the same day sends fewer persons to the small supermarkets, more to the others ...
    held = scenario.ScheduleGenerator(city, 111).generate_day(np.zeros(city.n_persons, np.uint8))
    v0 = np.bincount(free["loc"], minlength=city.n_locations)
    v1 = np.bincount(held["loc"], minlength=city.n_locations)
    assert (v1[small] < v0[small]).all() and v1[small].sum() < 0.6 * v0[small].sum()
    big = np.setdiff1d(markets, small)
    assert v1[big].sum() > v0[big].sum()
    peak1 = peak_occupancy(held, city.n_locations)
    assert np.maximum(peak1 - cap, 0)[small].sum() < 0.7 * (peak0 - cap)[small].sum()   # the overflow shrinks
    # ... nobody vanishes: the visits went to other supermarkets; elsewhere only the places change that somebody now
    # reaches from a different supermarket (the next search starts from where the person is)
    assert v1[markets].sum() == v0[markets].sum()
    other = np.ones(city.n_locations, bool)
    other[markets] = False
    assert (v1[other] != v0[other]).mean() < 0.005 and abs(int(v1[other].sum()) - int(v0[other].sum())) < 5


def test_without_saturation_nothing_changes():
    city = scenario.City(5000, seed=9)
    city.loc_area = (city.loc_area * 1000.0).astype(np.float32)
    a = scenario.ScheduleGenerator(city, 5).generate_day(np.zeros(city.n_persons, np.uint8))
    assert not scenario.capacity_diversion(city, 5).any()
    b = scenario.ScheduleGenerator(city, 5).generate_day(np.zeros(city.n_persons, np.uint8))
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
