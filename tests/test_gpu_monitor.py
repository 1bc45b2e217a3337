"""heros_sample_occupancy (the hourly persons-per-location / per-sub-location samples of HerosModel.locationDump,
model/HerosModel.java:266-299) against the same count made with numpy from the stays the engine holds."""
import gzip

import numpy as np
import pytest

import heros_b200 as hb
from common import load_props, occupancy_from_stays
from heros_b200 import properties, scenario

pytestmark = pytest.mark.gpu


def make(city, props):
    eng = hb.HerosEngine(hb.MODEL_DISTANCE, 111)
    city.install(eng)
    eng.set_transmission_distance(properties.dist_params(props))
    eng.set_progression(properties.prog_params(props))
    eng.seed_infections(20)
    return eng


@pytest.mark.parametrize("source", ["host", "device"])
def test_sampled_occupancy_equals_the_count_over_the_stays(source):
    city = scenario.City(5000, seed=13)
    eng = make(city, load_props("distance"))
    gen = scenario.ScheduleGenerator(city, 111)
    if source == "device":
        eng.set_schedule(scenario.compile_schedule(city), 111)
    first = eng.location_first_sub
    assert first[-1] == np.maximum(city.loc_nsub.astype(np.int64), 1).sum()
    for d in range(3):
        if source == "host":
            st = gen.generate_day(eng.agent_state()["phase"])
            eng.submit_visits(st["person"], st["loc"], st["sub"], st["t_in"], st["t_out"])
        else:
            eng.advance_schedule(d)
        times = np.r_[24.0 * d + np.arange(24.0), 24.0 * d + 8.25, 24.0 * d + 23.999]
        times.sort()
        per_loc, per_sub = eng.sample_occupancy(times)
        held = eng.debug_stays()
        ref_loc, ref_sub = occupancy_from_stays(held, first, times)
        np.testing.assert_array_equal(per_sub, ref_sub)
        np.testing.assert_array_equal(per_loc, ref_loc)
        assert 0.98 * city.n_persons <= per_loc[3].sum() <= city.n_persons   # 03:00: (nearly) everybody is at home
        assert (per_loc.sum(axis=1) <= city.n_persons).all()
        only_loc, none = eng.sample_occupancy(times[:2], sublocations=False)
        assert none is None
        np.testing.assert_array_equal(only_loc, ref_loc[:2])
        eng.step(24.0 * (d + 1))
    with pytest.raises(hb.HerosError):
        eng.sample_occupancy([5.0, 4.0])                    # not ascending
    with pytest.raises(hb.HerosError):
        eng.sample_occupancy([])
    eng.n_locations -= 1                                    # wrong output size
    with pytest.raises(hb.HerosError):
        eng.sample_occupancy([80.0])
    eng.n_locations += 1
    eng.close()


def test_location_dump_written_from_the_engine(tmp_path):
    city = scenario.City(2000, seed=2)
    eng = make(city, load_props("distance"))
    eng.set_schedule(scenario.compile_schedule(city), 7)
    dump = hb.LocationDump(eng, city.loc_type, str(tmp_path))
    for d in range(2):
        eng.advance_schedule(d)
        assert dump.dump_hours(24.0 * d, 24.0 * (d + 1)) == 24
        eng.step(24.0 * (d + 1))
    dump.close()
    with gzip.open(tmp_path / "locationNrPersons.csv.gz", "rt") as f:
        lines = f.read().splitlines()
    assert lines[0] == '"Time(h)","LocationNr","NrPersons"' and len(lines) == 1 + 48 * city.n_locations
    rows = np.array([[float(x) for x in line.split(",")] for line in lines[1:]])
    assert 0.98 * city.n_persons <= rows[rows[:, 0] == 27.0][:, 2].sum() <= city.n_persons
    eng.close()
