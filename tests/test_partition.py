"""Geographic partition of one city: Morton-ordered contiguous shards of equal expected load, persons owned
by the shard of their home -- and a shard's schedule generator produces exactly the stays the whole city's generator
produces for the shard's persons (draws are keyed by global ids), so cutting a city cannot change a run."""
import numpy as np
import pytest

from heros_b200 import scenario, sharded


@pytest.fixture(scope="module")
def city():
    return scenario.City(24000, seed=5, work_anywhere_fraction=0.1)


def test_morton_codes_order_points_along_the_z_curve():
    x = np.array([0.0, 1.0, 0.0, 1.0, 0.49, 0.51])
    y = np.array([0.0, 0.0, 1.0, 1.0, 0.49, 0.51])
    c = sharded.morton_codes(x, y)
    # quadrants in Z order: (0,0) < (1,0) < (0,1) < (1,1); the two points near the centre fall into the first / last quadrant
    assert list(np.argsort(c)) == [0, 4, 1, 2, 5, 3]


@pytest.mark.parametrize("n_shards", [1, 2, 3, 8])
def test_partition_is_contiguous_balanced_and_owned_by_home(city, n_shards):
    part = sharded.partition_city(city, n_shards)
    s = part.shard_of_location
    assert s.min() == 0 and s.max() == n_shards - 1 and len(s) == city.n_locations
    order = np.argsort(sharded.morton_codes(city.loc_x, city.loc_y), kind="stable")
    assert (np.diff(s[order]) >= 0).all()                       # contiguous pieces of the curve
    load = sharded.expected_load(city)
    per = np.array([load[s == r].sum() for r in range(n_shards)])
    assert per.max() / per.mean() < 1.02                        # equal expected load
    np.testing.assert_array_equal(part.shard_of_person, s[city.home_loc])
    # the new numbering makes every shard contiguous and is a permutation
    assert sorted(part.loc_new.tolist()) == list(range(city.n_locations))
    assert sorted(part.person_new.tolist()) == list(range(city.n_persons))
    for r in range(n_shards):
        ids = part.loc_new[s == r]
        assert ids.min() == part.loc_bases[r] and ids.max() == part.loc_bases[r + 1] - 1
        pid = part.person_new[part.shard_of_person == r]
        assert pid.min() == part.person_bases[r] and pid.max() == part.person_bases[r + 1] - 1


def test_shard_generators_reproduce_the_whole_city(city):
    part = sharded.partition_city(city, 3)
    whole = part.global_city()
    gen_w = scenario.ScheduleGenerator(whole, 111)
    shards = [part.shard_city(r) for r in range(3)]
    gens = [scenario.ScheduleGenerator(c, 111) for c in shards]
    for c in shards:  # homes are own, foreign locations sit behind the own ones
        assert (c.home_loc < c.n_locations).all() and len(c.loc_type) == c.n_locations + len(c.ext_global)
    n_remote = 0
    for d in range(3):
        phase = np.zeros(whole.n_persons, np.uint8)
        if d == 2:
            phase[::7] = 3   # a day with symptomatic persons (they stay at home)
            phase[5::97] = 4  # ... and a few in hospital (nearest hospital: may lie in another shard)
        ref = gen_w.generate_day(phase)
        rows = []
        for r, (c, g) in enumerate(zip(shards, gens)):
            st = g.generate_day(phase[part.person_bases[r]:part.person_bases[r + 1]])
            own, rem, off, _ = sharded.split_stays(c, st, r, part.person_bases, part.loc_bases)
            n_remote += len(rem["person"])
            assert off[-1] == len(rem["person"]) and off[r + 1] == off[r]        # nothing remote goes to oneself
            dest = np.searchsorted(part.loc_bases[1:], rem["loc"], side="right")
            assert (np.diff(dest) >= 0).all() and (dest != r).all()
            assert ((own["loc"] >= part.loc_bases[r]) & (own["loc"] < part.loc_bases[r + 1])).all()
            for cols in (own, rem):
                rows.append(np.c_[cols["person"], cols["loc"], cols["sub"], cols["t_in"], cols["t_out"]])
        got = np.concatenate(rows)
        want = np.c_[ref["person"], ref["loc"], ref["sub"], ref["t_in"], ref["t_out"]]
        got, want = got[np.lexsort(got.T[::-1])], want[np.lexsort(want.T[::-1])]
        np.testing.assert_array_equal(got, want)
    assert n_remote > 1000


def test_open_remote_stays_are_resent_until_closed():
    c = object.__new__(scenario.City)
    c.n_locations, c.ext_global = 2, np.array([7, 9], np.int64)
    bases_p, bases_l = np.array([0, 10, 20]), np.array([0, 2, 10])
    day0 = dict(person=np.array([1, 2], np.int32), loc=np.array([2, 3], np.int32), sub=np.zeros(2, np.int16),
                t_in=np.array([20.0, 21.0]), t_out=np.array([np.inf, 23.0]))
    own, rem, off, open_next = sharded.split_stays(c, day0, 0, bases_p, bases_l)
    assert len(own["person"]) == 0 and rem["loc"].tolist() == [7, 9] and open_next["person"].tolist() == [1]
    # next day: the closed version of the open stay arrives -> the open copy is dropped, nothing stays open
    day1 = dict(person=np.array([1], np.int32), loc=np.array([2], np.int32), sub=np.zeros(1, np.int16),
                t_in=np.array([20.0]), t_out=np.array([30.0]))
    _, rem, _, open_next = sharded.split_stays(c, day1, 0, bases_p, bases_l, open_next)
    assert rem["t_out"].tolist() == [30.0] and len(open_next["person"]) == 0
    # ... or it does not: the open stay is sent again
    _, rem, _, open_next2 = sharded.split_stays(c, {k: v[:0] for k, v in day1.items()}, 0, bases_p, bases_l,
                                                dict(person=np.array([1], np.int32), loc=np.array([7], np.int32),
                                                     sub=np.zeros(1, np.int16), t_in=np.array([20.0]), t_out=np.array([np.inf])))
    assert rem["person"].tolist() == [1] and np.isinf(rem["t_out"]).all() and open_next2["person"].tolist() == [1]


def test_europe_cities_are_linked_by_commuters_with_global_ids():
    """bench.py --config europe (BASELINE configs[3]): cities of their own size, one per GPU; the commuters' workplaces are
    foreign locations of the city table, their stays come out of the schedule generator with the other city's global ids."""
    import os
    import sys
    import types
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    args = types.SimpleNamespace(persons=20000, commute=0.1)
    world = 3
    shards = [bench.europe_city(args, r, world) for r in range(world)]
    for r, (c, pb, lb) in enumerate(shards):
        np.testing.assert_array_equal(pb, shards[0][1])
        np.testing.assert_array_equal(lb, shards[0][2])
        assert pb[r + 1] - pb[r] == c.n_persons and lb[r + 1] - lb[r] == c.n_locations
        far = c.work_loc >= c.n_locations
        assert 0.02 * c.n_persons < far.sum() < 0.1 * c.n_persons            # a tenth of the workers
        gid = c.ext_global[c.work_loc[far] - c.n_locations]
        owner = np.searchsorted(lb[1:], gid, side="right")
        assert (owner != r).all() and set(owner.tolist()) == set(range(world)) - {r}   # every other city is a destination
        st = scenario.ScheduleGenerator(c, 111).generate_day(np.zeros(c.n_persons, np.uint8))
        own, rem, off, _ = sharded.split_stays(c, st, r, pb, lb)
        assert len(rem["person"]) > 0.5 * far.sum() and off[r + 1] == off[r]
        assert ((rem["person"] >= pb[r]) & (rem["person"] < pb[r + 1])).all()
        assert (np.searchsorted(lb[1:], rem["loc"], side="right") != r).all()
        # the sub-location of a remote stay exists in the city that owns the workplace
        for dest in range(world):
            rows = slice(off[dest], off[dest + 1])
            if dest != r and off[dest + 1] > off[dest]:
                other = shards[dest][0]
                local = rem["loc"][rows] - lb[dest]
                assert (rem["sub"][rows] < np.maximum(other.loc_nsub[local], 1)).all()
                assert (other.loc_type[local] == scenario.TYPE_ID["Workplace"]).all()


def test_partition_properties_hold_for_random_point_sets():
    """Property test (hypothesis): whatever the coordinates and weights, the shards are contiguous pieces of the Morton
    curve, every location lies in exactly one of them, and no shard carries more than its share plus one location's weight."""
    from hypothesis import given, settings, strategies as st

    class Points:
        pass

    @settings(max_examples=40, deadline=None)
    @given(st.integers(1, 400), st.integers(1, 9), st.integers(0, 2 ** 31 - 1), st.booleans())
    def check(n, n_shards, seed, clustered):
        rng = np.random.default_rng(seed)
        c = Points()
        c.n_locations = n
        c.loc_x = (rng.normal(0, 1, n) if clustered else rng.random(n) * 1000).astype(np.float32)
        c.loc_y = (rng.normal(0, 1, n) if clustered else np.zeros(n)).astype(np.float32)   # also a degenerate line
        c.home_loc = rng.integers(0, n, 3 * n)
        w = rng.random(n) * rng.choice([1.0, 50.0], n) + 1e-3
        part = sharded.partition_city(c, n_shards, weights=w)
        s = part.shard_of_location
        assert s.min() >= 0 and s.max() < n_shards and len(s) == n
        order = np.argsort(sharded.morton_codes(c.loc_x, c.loc_y), kind="stable")
        assert (np.diff(s[order]) >= 0).all()
        per = np.bincount(s, weights=w, minlength=n_shards)
        assert per.max() <= w.sum() / n_shards + w.max() + 1e-9
        assert part.loc_bases[-1] == n and part.person_bases[-1] == len(c.home_loc)
        np.testing.assert_array_equal(np.sort(part.loc_new), np.arange(n))
        np.testing.assert_array_equal(part.shard_of_person, s[c.home_loc])

    check()
