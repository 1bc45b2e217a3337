"""CPU-side checks: the C-ABI library loads and exports every symbol include/heros_gpu.h declares (no compute calls),
struct layouts, host-side parameter parsing, the schedule generator, the reference-API mirror's error behaviour."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import heros_b200 as hb
from common import GOLDEN, ROOT, load_props, occupancy_from_stays, write_policy_csv
from heros_b200 import _lib, properties, scenario


def header_functions():
    text = open(os.path.join(ROOT, "include", "heros_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(heros_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = header_functions()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in heros_gpu.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared        # the ctypes table covers the header exactly
    assert lib.heros_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_the_header():
    assert C.sizeof(_lib.Config) == 40
    assert C.sizeof(_lib.AreaParams) == 6 * 8 and C.sizeof(_lib.DistParams) == 10 * 8
    assert C.sizeof(_lib.Duration) == 32
    assert C.sizeof(_lib.ProgParams) == 5 * 2 * 128 * 8 + 10 * 32
    assert C.sizeof(_lib.StepStats) == 27 * 8


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(hb.HerosError) as e:
        hb.HerosEngine(hb.MODEL_AREA, 1)
    assert e.value.code == _lib.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_create_rejects_bad_configs_before_touching_the_device():
    lib = _lib.load()
    h = C.c_void_p()
    for cfg in (_lib.Config(99, 0, 0, 0, 1, 24.0, 0, 0), _lib.Config(1, 7, 0, 0, 1, 24.0, 0, 0),
                _lib.Config(1, 0, 5, 0, 1, 24.0, 0, 0), _lib.Config(1, 0, 0, 0, 1, 0.0, 0, 0)):
        assert lib.heros_create(C.byref(cfg), C.byref(h)) == _lib.ERR_INVALID
        assert lib.heros_last_error(None)
    assert lib.heros_create(None, C.byref(h)) == _lib.ERR_INVALID
    assert lib.heros_destroy(None) == _lib.OK


def test_properties_of_the_reference_parse_to_its_values():
    p = load_props("area")
    a = properties.area_params(p)
    assert (a.contagiousness, a.beta, a.t_e_min, a.t_e_mode, a.t_e_max, a.calculation_threshold) == \
           (1.0, 1.0, 2.0, 3.4, 9.6, 60.0)                       # alpha-area.properties:7-24
    d = properties.dist_params(load_props("distance"))
    assert (d.L, d.I, d.C, d.v_max, d.v_0, d.r, d.psi, d.alpha, d.mu) == (2.0, 3.4, 6.2, 7.23, 4.0, 2.294, 1.0, 0.05, 0.0)
    fr, periods = properties.prog_tables(p)
    assert fr.shape == (5, 2, 128)
    assert (fr[0] == 0.46).all()                                  # FractionAsymptomatic
    assert fr[1, 0, 19] == 0.02153 and fr[1, 1, 20] == 0.01648 and fr[1, 0, 65] == 0.44038 and fr[1, 0, 100] == 0.12687
    assert fr[1, 0, 101] == 0.0                                   # outside every band
    assert fr[2, 0, 95] == 0.0 and fr[3].max() == 0.0 and fr[4, 0, 49] == 0.0 and fr[4, 1, 90] == 0.67882
    assert periods[0] == (_lib.DIST_TRIANGULAR, 2.5, 3.4, 3.8) and periods[9] == (_lib.DIST_TRIANGULAR, 28.0, 30.0, 32.0)


def test_conditional_probability_and_distribution_syntax():
    t = properties.parse_conditional_probability("gender{M:0.45, F:0.5}")
    assert (t[0] == 0.45).all() and (t[1] == 0.5).all()
    t = properties.parse_conditional_probability("age{0-19: 0.8, 20-55: 0.5, 56-100: 0.3}")
    assert t[0, 19] == 0.8 and t[1, 20] == 0.5 and t[0, 100] == 0.3
    t = properties.parse_conditional_probability("age{0-12: 0.000, 10-19: 0.1}")   # overlapping: first band wins
    assert t[0, 11] == 0.0 and t[0, 13] == 0.1
    assert properties.parse_distribution("Uniform(2,4)") == (_lib.DIST_UNIFORM, 2.0, 4.0, 0.0)
    assert properties.parse_distribution("Constant(3)") == (_lib.DIST_CONSTANT, 3.0, 0.0, 0.0)
    assert properties.parse_distribution(" triangular( 1 ,3, 5 ) ")[0] == _lib.DIST_TRIANGULAR
    with pytest.raises(NotImplementedError):
        properties.parse_distribution("TruncatedNormal(12.0, 2.3, 12.0, 14.0)")
    with pytest.raises(ValueError):
        properties.parse_distribution("garbage")


def test_location_type_fixture_matches_the_generator_tables():
    with open(os.path.join(GOLDEN, "locationtypes_thehague.json")) as f:
        rows = json.load(f)["rows"]
    assert [r[0] for r in rows] == scenario.TYPE_NAMES          # type id = row order (ConstructHerosModel.java:233)
    assert [int(r[3].upper() == "TRUE") for r in rows] == scenario.TYPE_INFECT_SUB.tolist()
    assert [float(r[5]) for r in rows] == scenario.TYPE_CORRECTION.tolist()


def test_vectorised_philox_matches_the_oracle():
    from oracle import binding as ob
    x, y, z, w = scenario.philox4x32_10(np.array([0, 0xffffffff, 0x243f6a88]), np.array([0, 0xffffffff, 0x85a308d3]),
                                        np.array([0, 0xffffffff, 0x13198a2e]), np.array([0, 0xffffffff, 0x03707344]),
                                        0, 0)
    assert [int(x[0]), int(y[0]), int(z[0]), int(w[0])] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    for person in (0, 5, 70000):
        a, b, _, _ = scenario.philox4x32_10(np.array([person]), 7, 0, 0x300, 111, 0)
        got = float(scenario.u01_from_words(a, b)[0])
        assert got == ob.u01(111, person, 7, 0x300)


def test_schedule_generator_is_deterministic_and_consistent():
    city = scenario.City(3000, seed=3)
    assert city.loc_nsub.min() >= 1 and (city.home_sub < city.loc_nsub[city.home_loc]).all()
    has_work = city.work_loc >= 0
    assert (city.work_sub[has_work] < city.loc_nsub[city.work_loc[has_work]]).all()
    days = []
    for rep in range(2):
        gen = scenario.ScheduleGenerator(city, 111)
        phase = np.zeros(city.n_persons, np.uint8)
        phase[:40] = 3    # Infected-Symptomatic: stay home
        phase[40:50] = 4  # Hospitalized
        phase[50:55] = 6  # Dead
        days.append([gen.generate_day(phase) for _ in range(3)])
    for a, b in zip(*days):
        for k in a:
            np.testing.assert_array_equal(a[k], b[k])
    d0, d1, d2 = days[0]
    closed = np.isfinite(d1["t_out"])
    assert (d1["t_out"][closed] > d1["t_in"][closed]).all()
    assert (d1["t_out"][closed] <= 48.0).all() and (d1["t_out"][closed] >= 24.0).all()
    assert (d1["sub"] < city.loc_nsub[d1["loc"]]).all()
    # per person the stays of a day do not overlap
    for d in (d1, d2):
        o = np.lexsort((d["t_in"], d["person"]))
        p, a, b = d["person"][o], d["t_in"][o], d["t_out"][o]
        same = p[1:] == p[:-1]
        assert (a[1:][same] >= b[:-1][same]).all()
    # symptomatic persons do not move on day 1; the dead have left the model
    moved = np.unique(d1["person"])
    assert not np.isin(np.arange(0, 40), moved).any()
    assert not np.isin(np.arange(50, 55), np.unique(d2["person"])).any()
    hosp = d0["loc"][np.isin(d0["person"], np.arange(40, 50)) & ~np.isfinite(d0["t_out"])]
    assert (city.loc_type[hosp] == scenario.TYPE_ID["Hospital"]).all()
    assert 3.0 < len(d1["person"]) / city.n_persons < 6.0


def test_mirror_classes_keep_the_reference_error_behaviour(capsys):
    # DiseasePolicy rejects anything but covidT_dist.psi / mu with a message, no exception (DiseasePolicy.java:21-25)
    class FakeModel:
        def getDiseaseTransmission(self):
            raise AssertionError("must not be reached")
    hb.DiseasePolicy(FakeModel(), 240.0, "covidT_area.beta", 0.5)
    assert "Unrecognized variable name covidT_area.beta" in capsys.readouterr().err
    with pytest.raises(ValueError):
        hb.HerosModel({"generic.diseasePropertiesModel": "sir"})


def test_helsinki_location_fixture_is_the_reference_table():
    """tests/golden/helsinki_locations.npz (tools/make_helsinki_fixture.py) is data/helsinki/locations.csv.gz, the one
    real location table the reference ships: row count, sub-location total and the quirk rows SURVEY.md 8a A12 lists."""
    tab = dict(np.load(os.path.join(GOLDEN, "helsinki_locations.npz")))
    assert len(tab["type"]) == 69508 and int(tab["nsub"].sum()) == 431470 and int(tab["nsub"].max()) == 650
    acc, wp = tab["type"] == scenario.TYPE_ID["Accommodation"], tab["type"] == scenario.TYPE_ID["Workplace"]
    assert int(acc.sum()) == 20512 and int(wp.sum()) == 42169
    assert int(((tab["nsub"] == 0) & acc).sum()) == 76 and int((tab["nsub"] == 0).sum()) == 76
    assert int(((tab["area_per_sub"] < 0) & wp).sum()) == 3563 and (tab["nsub"][tab["area_per_sub"] < 0] == 1).all()
    parks = tab["type"] == scenario.TYPE_ID["Park"]
    assert (tab["nsub"][parks] == 1).all() and (tab["area_per_sub"][parks] == 300000.0).all()
    assert 60.12 < tab["lat"].min() and tab["lat"].max() < 60.30 and 24.83 < tab["lon"].min() and tab["lon"].max() < 25.25
    city = scenario.City(4000, seed=5, location_table=tab)
    assert city.n_locations == 69508 and int(city.loc_nsub.sum()) == 431470
    assert (city.loc_area[tab["nsub"] == 0] == 0.0).all()          # totalArea = area * 0 (ConstructHerosModel.java:381)
    assert (city.home_sub < np.maximum(city.loc_nsub[city.home_loc], 1)).all()


def test_compiled_schedule_tables_are_consistent():
    """scenario.compile_schedule (what heros_set_schedule takes) against the activity rows it was compiled from."""
    city = scenario.City(1500, seed=2)
    pat = scenario.load_patterns()
    t = scenario.compile_schedule(city, pat)
    assert t["activities"].dtype.itemsize == C.sizeof(_lib.Activity) == 48
    start, count = t["pattern_start"].reshape(8, 16, 4), t["pattern_count"].reshape(8, 16, 4)
    assert count.max() <= 32 and (start + count <= len(t["activities"])).all()
    rows = pat["patterns"]["Susceptible|worker|Monday"]
    ph, role = pat["phases"].index("Susceptible"), pat["roles"].index("worker") + 1
    got = t["activities"][start[ph, role, 0]:start[ph, role, 0] + count[ph, role, 0]]
    assert [int(a["kind"]) for a in got] == [r["kind"] for r in rows]
    assert [float(a["max"]) for a in got] == [r["max"] for r in rows]
    assert (count[pat["phases"].index("Dead")] == 0).all()                      # no pattern: a dead person leaves the model
    hospital = t["activities"][start[4, role, 0]:start[4, role, 0] + count[4, role, 0]]
    assert len(hospital) >= 1                                                    # Hospitalized: to the nearest hospital
    cs = t["choice_start"]
    assert (np.diff(cs) > 0).all() and cs[-1] == len(t["choice_type"]) == len(t["choice_cum"])
    for c in range(len(cs) - 1):
        assert (np.diff(t["choice_cum"][cs[c]:cs[c + 1]]) >= 0).all()
    assert t["nearest"].shape == (city.n_locations, len(scenario.TYPE_NAMES)) and t["cand"].shape[2] == t["n_candidates"]


class _QueueModel:
    """what LocationPolicy needs of the model on a box without a GPU: somewhere to queue"""


def test_location_policy_files_close_and_reopen_location_types(capsys, tmp_path):
    """data/thehague/policies/HardLockdown_Realistic_10-40.csv (tests/golden/location_policies.json, tools/make_policy_fixtures.py)
    through the LocationPolicy mirror into the host-side schedule: closed types are not visited from day 10 on, the
    partially open ones keep a subset of their locations and of their activities, everything is back on day 40."""
    city = scenario.City(4000, seed=5)
    T = scenario.TYPE_ID
    model = _QueueModel()
    pol = hb.schedule_location_policies(model, write_policy_csv("HardLockdown_Realistic_10-40", tmp_path))
    assert len(pol) == 26 and all(p.scheduled for p in pol)
    assert pol[0].time == 240.0 and pol[0].locationType == T["Workplace"] and pol[0].fractionOpen == 0.2
    assert pol[0].alternativeLocationType == T["Accommodation"] and pol[0].reportAsLocationName == "WorkToHome"
    hb.LocationPolicy(model, 0.0, "Casino", 0.0, 0.0, "Accommodation", "StayHome")        # LocationPolicy.java:24-28
    hb.LocationPolicy(model, 0.0, "Retail", 0.0, 0.0, "Moon", "StayHome")                  # :30-34
    err = capsys.readouterr().err
    assert "Did not recognize locationType Casino" in err and "Did not recognize alternativeLocationType Moon" in err
    assert len(model.locationPolicies) == 26

    gen, free = scenario.ScheduleGenerator(city, 111), scenario.ScheduleGenerator(city, 111)
    phase = np.zeros(city.n_persons, np.uint8)
    per_day, per_day_free = [], []
    for d in range(42):
        fired = hb.apply_location_policies(model, 24.0 * d, gen)
        assert fired == (13 if d in (10, 40) else 0)
        st, sf = gen.generate_day(phase), free.generate_day(phase)
        per_day.append(st)
        per_day_free.append(sf)
        if d < 10:   # before the lockdown: the same stays
            for k in st:
                np.testing.assert_array_equal(st[k], sf[k])

    def count(st, name):
        return int((city.loc_type[st["loc"]] == T[name]).sum())

    for d in range(11, 40):   # day 10 still closes the stays opened on day 9
        for name in ("Mall", "BarRestaurant", "PrimarySchool", "University", "Recreation", "Park"):
            assert count(per_day[d], name) == 0, (d, name)
    work_lock = sum(count(per_day[d], "Workplace") for d in range(14, 19))      # Mon-Fri
    work_free = sum(count(per_day_free[d], "Workplace") for d in range(14, 19))
    assert 0.03 * work_free < work_lock < 0.25 * work_free                       # 0.2 x 0.5 of the work stays, roughly
    # the workplaces still visited are a fixed subset (keyed by the location), not a fresh draw every day
    open_places = [np.unique(per_day[d]["loc"][city.loc_type[per_day[d]["loc"]] == T["Workplace"]]) for d in (14, 15, 16)]
    all_places = np.unique(np.concatenate(open_places))
    free_places = np.unique(per_day_free[14]["loc"][city.loc_type[per_day_free[14]["loc"]] == T["Workplace"]])
    assert len(all_places) < 0.35 * len(free_places)
    for d in (40, 41):
        assert count(per_day[d], "PrimarySchool") > 0 or d % 7 >= 5
    # day 41: the stays that begin that day are the ones of a run without any policy
    tables = []
    for st in (per_day[41], per_day_free[41]):
        m = st["t_in"] >= 24.0 * 41
        o = np.lexsort((st["t_in"][m], st["person"][m]))
        tables.append(np.stack([st[k][m][o].astype(np.float64) for k in ("person", "loc", "sub", "t_in", "t_out")]))
    np.testing.assert_array_equal(tables[0], tables[1])
    with pytest.raises(ValueError):
        gen.set_closure_policy(T["Retail"], 0.5, 0.5, T["Park"])       # only home or the type itself
    with pytest.raises(ValueError):
        gen.set_closure_policy(T["Retail"], 1.5, 0.5, T["Accommodation"])


def test_disease_policy_file_reader(tmp_path):
    import csv
    calls = []
    with open(os.path.join(GOLDEN, "policy_masking_distance_10_40.json")) as f:   # data/thehague/policies/MaskingDistance_10-40.csv
        fixture = json.load(f)
    path = str(tmp_path / "MaskingDistance_10-40.csv")
    with open(path, "w", newline="") as f:
        csv.writer(f).writerows([fixture["header"]] + fixture["rows"])

    class Transmission:
        def setParameter(self, name, value, at_time=None):
            calls.append((name, value, at_time))

    class Model:
        def getDiseaseTransmission(self):
            return Transmission()

    hb.schedule_disease_policies(Model(), path)
    assert calls == [("psi", 2.0, 240.0), ("mu", 0.6, 240.0), ("psi", 1.0, 960.0), ("mu", 0.0, 960.0)]


def test_location_dump_files_have_the_reference_layout(tmp_path):
    """HerosModel.locationDump (HerosModel.java:266-299): header lines, one row per location and hour, sub-location rows
    for the indices 1 .. nSub-1 only."""
    import gzip
    city = scenario.City(1500, seed=9)
    gen = scenario.ScheduleGenerator(city, 5)
    st = gen.generate_day(np.zeros(city.n_persons, np.uint8))

    class FakeEngine:
        location_n_sub = city.loc_nsub
        location_first_sub = np.concatenate([[0], np.cumsum(np.maximum(city.loc_nsub.astype(np.int64), 1))])

        def sample_occupancy(self, times):
            return occupancy_from_stays(st, self.location_first_sub, times)

    dump = hb.LocationDump(FakeEngine(), city.loc_type, str(tmp_path))
    assert dump.dump_hours(0.0, 24.0) == 24
    dump.close()
    with gzip.open(tmp_path / "locationNrPersons.csv.gz", "rt") as f:
        loc_lines = f.read().splitlines()
    with gzip.open(tmp_path / "sublocationNrPersons.csv.gz", "rt") as f:
        sub_lines = f.read().splitlines()
    assert loc_lines[0] == '"Time(h)","LocationNr","NrPersons"'
    assert sub_lines[0] == '"Time(h)","LocationNr","SubLocationNr","NrPersons"'
    assert len(loc_lines) == 1 + 24 * city.n_locations
    assert len(sub_lines) == 1 + 24 * int(np.maximum(city.loc_nsub.astype(np.int64) - 1, 0).sum())
    rows = np.array([[float(x) for x in line.split(",")] for line in loc_lines[1:]])
    assert rows[0, 0] == 0.0 and loc_lines[1].startswith("0.0,") and loc_lines[-1].startswith("23.0,")
    # everybody is somewhere at 03:00 (at home), fewer are inside a location at 08:30-ish travel times
    at3 = rows[rows[:, 0] == 3.0]
    assert at3[:, 2].sum() == city.n_persons
    first = rows[rows[:, 0] == 0.0]
    assert (np.diff(city.loc_type[first[:, 1].astype(int)].astype(int)) >= 0).all()      # grouped by location type
    srow = np.array([[float(x) for x in line.split(",")] for line in sub_lines[1:]])
    assert srow[:, 2].min() == 1 and (srow[:, 2] < city.loc_nsub[srow[:, 1].astype(int)]).all()
    per_loc, per_sub = occupancy_from_stays(st, FakeEngine.location_first_sub, [12.0])
    noon = srow[srow[:, 0] == 12.0]
    pos = FakeEngine.location_first_sub[noon[:, 1].astype(int)] + noon[:, 2].astype(int)
    np.testing.assert_array_equal(noon[:, 3].astype(np.int32), per_sub[0, pos])


def test_schedule_generator_with_commuters_into_another_tile():
    """bench.py's multi-GPU setup: some workers of a tile work in the next tile (location codes beyond the own table);
    the generator -- with and without a closure policy in force -- sends them there / keeps them home."""
    city = scenario.City(3000, seed=21)
    other = scenario.City(3000, seed=22, origin_m=(12000.0, 0.0))
    n = city.add_commuters(other, 0.2, seed=5)
    assert n > 0 and (city.work_loc >= city.n_locations).sum() == n
    gen = scenario.ScheduleGenerator(city, 111)
    phase = np.zeros(city.n_persons, np.uint8)
    st = gen.generate_day(phase)
    remote = st["loc"] >= city.n_locations
    assert remote.sum() > 0.5 * n                                       # Monday: the commuters go to work
    glob = city.global_location_ids(st["loc"], 1000000, 2000000)
    assert (glob[remote] >= 2000000).all() and (glob[~remote] < 2000000).all()
    gen.set_closure_policy(scenario.TYPE_ID["Workplace"], 0.0, 0.0, scenario.TYPE_ID["Accommodation"])
    st = gen.generate_day(phase)                                        # own workplaces closed; the other tile's are not ours to close
    ty = city.loc_type[st["loc"][st["loc"] < city.n_locations]]
    opened_today = st["t_in"][st["loc"] < city.n_locations] >= 24.0
    assert ((ty == scenario.TYPE_ID["Workplace"]) & opened_today).sum() == 0
    assert ((st["loc"] >= city.n_locations) & (st["t_in"] >= 24.0)).sum() > 0.5 * n


def test_loaders_read_the_reference_file_formats(tmp_path, capsys):
    """locations.csv.gz / people.csv.gz in the column layout of data/<city>/ (ConstructHerosModel.java:300-395, :450-758):
    ids are mapped to dense indices, households get their sub-location in order of first appearance, the rows the
    reference rejects are reported on stderr with its messages and skipped."""
    import csv
    import gzip
    from heros_b200 import loaders
    loc_rows = [["location_id", "lon", "lat", "nb_sublocations", "location_category", "area", "extra"],
                [100, 4.30, 52.07, 3, "Accommodation", 100.0, "x"], [101, 4.31, 52.08, 2, "Accommodation", 80.0, "x"],
                [205, 4.32, 52.07, 4, "Workplace", 30.0, "x"], [300, 4.30, 52.09, 6, "Kindergarten", 50.0, "x"],
                [301, 4.33, 52.06, 10, "PrimarySchool", 75.0, "x"], [400, 4.31, 52.07, 1, "Supermarket", 200.0, "x"],
                [401, 4.32, 52.08, 1, "Casino", 10.0, "x"], [500, 4.30, 52.07, 1, "Park", 300000.0, "x"]]
    lp = str(tmp_path / "locations.csv.gz")
    with gzip.open(lp, "wt", newline="") as f:
        csv.writer(f).writerows(loc_rows)
    people_rows = [["person_id", "household_id", "age", "home_id", "workplace_id", "social_role"],
                   [1, 10, 34.0, 100, 205, 7], [2, 10, 36, 100, 205.0, 7], [3, 11, 70, 100, "", 8],
                   [4, 12, 4, 100, 300, 2], [5, 13, 8, 100, 301, 3],          # household 13: 4th in a home of 3 dwellings
                   [6, 20, 30, 999, 205, 7],                                   # unknown home
                   [7, 20, 30, 101, "", 7],                                    # worker without a work location
                   [8, 20, 5, 101, 301, 2],                                    # kindergarten child registered at a primary school
                   [9, 21, 130, 101, "", 9],                                   # age out of range: reported, kept
                   [10, 21, 1, 101, "", 1], [11, 22, 40, 101, 205, 42]]        # unknown role
    pp = str(tmp_path / "people.csv.gz")
    with gzip.open(pp, "wt", newline="") as f:
        csv.writer(f).writerows(people_rows)
    locs = loaders.read_locations(lp)
    err = capsys.readouterr().err
    assert "Location type Casino not found on line 8" in err
    assert locs["location_id"].tolist() == [100, 101, 205, 300, 301, 400, 500]
    assert locs["type"].tolist() == [scenario.TYPE_ID[n] for n in ("Accommodation", "Accommodation", "Workplace", "Kindergarten",
                                                                     "PrimarySchool", "Supermarket", "Park")]
    assert locs["nsub"].tolist() == [3, 2, 4, 6, 10, 1, 1] and locs["area_per_sub"][2] == 30.0
    people = loaders.read_people(pp, locs)
    err = capsys.readouterr().err
    assert "homeId 999 not found in the location map on line 7" in err
    assert "No work location [-1] for Worker on line 8" in err
    assert "workSchoolId 301 not a kindergarten in the location map on line 9" in err
    assert "Person 9 has age 130 on row 10" in err
    assert "has more sublocations (4) than defined" in err and "social role 42 not recognized" in err
    assert people["person_id"].tolist() == [1, 2, 3, 4, 5, 9, 10]
    assert people["home_loc"].tolist() == [0, 0, 0, 0, 0, 1, 1]
    assert people["home_sub"].tolist() == [0, 0, 1, 2, 3, 1, 1]      # household 20 took dwelling 0 of home 101 before being dropped
    assert people["work_loc"].tolist() == [2, 2, -1, 3, 4, -1, -1] and people["role"].tolist() == [7, 7, 8, 2, 3, 9, 1]
    assert set(people["female"].tolist()) <= {0, 1}
    with pytest.raises(ValueError):
        loaders._columns(["person_id", "age"], ("person_id", "household_id"), "Person")
    city = loaders.load_city(lp, pp)
    capsys.readouterr()
    assert city.n_persons == 7 and city.n_locations == 7 and city.person_id.tolist() == [1, 2, 3, 4, 5, 9, 10]
    assert city.loc_area[2] == np.float32(30.0 * 4)                      # total surface = area * nb_sublocations as float
    assert (city.work_sub[:2] < 4).all() and city.work_sub[2] == 0
    # the tables are what the engine and the schedule take
    tables = scenario.compile_schedule(city)
    assert tables["nearest"].shape == (7, len(scenario.TYPE_NAMES)) and tables["work_sub"].shape == (7,)
    assert tables["nearest"][0, scenario.TYPE_ID["Supermarket"]] == 5
    gen = scenario.ScheduleGenerator(city, 111)
    st = gen.generate_day(np.zeros(7, np.uint8))
    assert (st["sub"] < np.maximum(city.loc_nsub[st["loc"]], 1)).all() or (st["sub"][st["loc"] == 0] <= 3).all()
    assert 2 in st["loc"].tolist()                                        # the workers went to work


def test_location_loader_reproduces_the_helsinki_fixture():
    """data/helsinki/locations.csv.gz through loaders.read_locations == tests/golden/helsinki_locations.npz (only where the
    reference tree is present: this container, not the GPU box)."""
    from heros_b200 import loaders
    path = "/root/reference/data/helsinki/locations.csv.gz"
    if not os.path.exists(path):
        pytest.skip("the reference tree is not on this machine")
    locs = loaders.read_locations(path)
    fix = np.load(os.path.join(GOLDEN, "helsinki_locations.npz"))
    assert len(locs["type"]) == 69508 and locs["location_id"].tolist() == list(range(69508))
    for k in ("type", "nsub", "area_per_sub", "lat", "lon"):
        np.testing.assert_array_equal(locs[k], fix[k])


def _write_xlsx(path, sheet_name, rows):
    """a minimal workbook with inline strings and numbers (what heros_b200.xlsx reads)"""
    import zipfile
    from xml.sax.saxutils import escape

    def col(i):
        s = ""
        i += 1
        while i:
            i, r = divmod(i - 1, 26)
            s = chr(65 + r) + s
        return s
    body = []
    for r, row in enumerate(rows, start=1):
        cells = []
        for c, v in enumerate(row):
            if v is None:
                continue
            ref = f"{col(c)}{r}"
            if isinstance(v, (int, float)):
                cells.append(f'<c r="{ref}"><v>{v}</v></c>')
            else:
                cells.append(f'<c r="{ref}" t="inlineStr"><is><t>{escape(str(v))}</t></is></c>')
        body.append(f'<row r="{r}">{"".join(cells)}</row>')
    ns = "http://schemas.openxmlformats.org/spreadsheetml/2006/main"
    rel = "http://schemas.openxmlformats.org/officeDocument/2006/relationships"
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("xl/workbook.xml", f'<workbook xmlns="{ns}" xmlns:r="{rel}"><sheets><sheet name="{sheet_name}" sheetId="1" r:id="rId1"/></sheets></workbook>')
        z.writestr("xl/_rels/workbook.xml.rels", '<Relationships xmlns="http://schemas.openxmlformats.org/package/2006/relationships">'
                   '<Relationship Id="rId1" Type="x" Target="worksheets/sheet1.xml"/></Relationships>')
        z.writestr("xl/worksheets/sheet1.xml", f'<worksheet xmlns="{ns}"><sheetData>{"".join(body)}</sheetData></worksheet>')


def test_activity_schedule_workbook_reader(tmp_path):
    """rows in the column layout of data/<city>/activities/activityschedules*.xlsx (ConstructHerosModel.java:763-904)"""
    from heros_b200 import loaders
    header = ["policy", "phase", "day", "role", "name", "class", "until", "dist", "mode", "min", "max", None, "locator",
              None, "travel locator", "max distance", "types"]
    rows = [header,
            ["0", "Susceptible", "Monday", "worker", "sleep", "UntilFixedTimeActivity", 7, None, None, None, None, None, "HomeLocator"],
            ["0", "Susceptible", "Monday", "worker", "to work", "TravelActivityDistanceBased", None, None, None, 0.2, 0.5, None, None, None, "WorkLocator"],
            ["0", "Susceptible", "Monday", "worker", "work", "StochasticDurationActivity", None, "triangular", 8, 7, 9, None, "WorkLocator"],
            ["0", "Susceptible", "Monday", "worker", "shop", "FixedDurationActivity", None, None, 0.5, None, None, None, "RandomLocatorCap", None, None,
             2500, "LocationType.Supermarket:0.7; LocationType.Retail:0.3"],
            ["0", "Susceptible", "Monday", "worker", "home", "UntilFixedTimeActivity", 24, None, None, None, None, None, "HomeLocator"],
            ["0", "Susceptible", "Monday", "worker", "sleep", "UntilFixedTimeActivity", 7, None, None, None, None, None, "HomeLocator"],   # second copy of the block
            ["1", "Susceptible", "Monday", "worker", "other policy", "UntilFixedTimeActivity", 24, None, None, None, None, None, "HomeLocator"],
            ["0", "Infected-Symptomatic", "Monday", "worker", "ill", "UntilFixedTimeActivity", 24, None, None, None, None, None, "HomeLocator"]]
    path = str(tmp_path / "activityschedules.xlsx")
    _write_xlsx(path, "activityschedules", rows)
    pat = loaders.read_activity_schedules(path)
    day = pat["patterns"]["Susceptible|worker|Monday"]
    assert [a["name"] for a in day] == ["sleep", "to work", "work", "shop", "home"]           # cut at the first "until 24"
    assert [a["kind"] for a in day] == [2, 1, 0, 0, 2] and [a["locator"] for a in day] == [0, 1, 1, 4 | loaders.LOCATOR_CAP, 0]  # the Cap form keeps its flag
    assert day[1]["dist"] == 1 and (day[1]["min"], day[1]["max"]) == (0.2, 0.5)                # travel time ~ U(min, max)
    assert day[2]["dist"] == 2 and (day[2]["min"], day[2]["mode"], day[2]["max"]) == (7.0, 8.0, 9.0)
    assert day[3]["dist"] == 0 and day[3]["min"] == day[3]["max"] == 0.5 and day[3]["max_distance"] == 2500.0
    assert pat["choices"][day[3]["choice"]] == [["Supermarket", 0.7], ["Retail", 0.3]]
    assert list(pat["patterns"]) == ["Susceptible|worker|Monday", "Infected-Symptomatic|worker|Monday"]
    bad = [header, ["0", "Susceptible", "Monday", "worker", "x", "TeleportActivity", 24, None, None, None, None, None, "HomeLocator"]]
    _write_xlsx(path, "activityschedules", bad)
    with pytest.raises(ValueError):
        loaders.read_activity_schedules(path)


def test_activity_schedule_reader_reproduces_the_packaged_patterns():
    """data/thehague/activities/activityschedules_cap.xlsx through loaders.read_activity_schedules == the packaged
    activity_patterns_thehague.json (only where the reference tree is present)."""
    from heros_b200 import loaders
    path = "/root/reference/data/thehague/activities/activityschedules_cap.xlsx"
    if not os.path.exists(path):
        pytest.skip("the reference tree is not on this machine")
    got = loaders.read_activity_schedules(path)
    want = scenario.load_patterns()
    assert got["patterns"] == want["patterns"] and got["choices"] == want["choices"]
    assert got["roles"] == want["roles"] and got["phases"] == want["phases"]


def _validate(city, tables, **override):
    """heros_validate_schedule_tables (host code of the library, no device needed) on the tables of a city"""
    lib = _lib.load()
    t = dict(tables)
    cols = dict(n_sub=city.loc_nsub, home_loc=city.home_loc, home_sub=city.home_sub, work_loc=city.work_loc,
                work_sub=t["work_sub"], nearest=t["nearest"], n_cand=t["n_cand"], cand=t["cand"],
                choice_start=t["choice_start"], choice_type=t["choice_type"])
    acc = int(override.pop("acc", t["accommodation_type"]))
    cols.update(override)
    dt = dict(n_sub=np.int16, home_loc=np.int32, home_sub=np.int16, work_loc=np.int32, work_sub=np.int16, nearest=np.int32,
              n_cand=np.int32, cand=np.int32, choice_start=np.int32, choice_type=np.int32)
    a = {k: np.ascontiguousarray(v, dt[k]) for k, v in cols.items()}
    p = lambda k: a[k].ctypes.data_as(C.c_void_p)
    msg = C.create_string_buffer(200)
    rc = lib.heros_validate_schedule_tables(city.n_locations, p("n_sub"), city.n_persons, p("home_loc"), p("home_sub"),
                                            p("work_loc"), p("work_sub"), int(t["n_types"]), int(t["n_candidates"]),
                                            p("nearest"), p("n_cand"), p("cand"), len(a["choice_start"]) - 1,
                                            p("choice_start"), p("choice_type"), acc, msg, 200)
    return rc, msg.value.decode()


def test_schedule_tables_are_range_checked_on_the_host():
    """What heros_set_schedule refuses before a kernel could index out of bounds; the tables of the synthetic cities and
    of the Helsinki city pass."""
    tab = dict(np.load(os.path.join(GOLDEN, "helsinki_locations.npz")))
    for city in (scenario.City(800, seed=3), scenario.City(6000, seed=7), scenario.City(20000, seed=5, location_table=tab)):
        tables = scenario.compile_schedule(city)
        assert _validate(city, tables) == (0, "")
    city = scenario.City(800, seed=3)
    tables = scenario.compile_schedule(city)
    n_loc = city.n_locations

    def broken(name, index, value, source=None):
        arr = np.array(source if source is not None else (tables[name] if name in tables else getattr(city, name))).copy()
        arr[index] = value
        return {name: arr}

    worker = int(np.nonzero(city.work_loc >= 0)[0][0])
    cases = [(broken("home_loc", 5, n_loc), "home location"), (broken("home_loc", 5, -1), "home location"),
             (broken("home_sub", 5, 300), "home location"), (broken("work_loc", worker, n_loc + 7), "work location lies outside"),
             (broken("work_sub", worker, 200), "work sub-location"), (broken("nearest", (3, 6), n_loc), "nearest-location table"),
             (broken("n_cand", (3, 6), 9), "candidate count"), (broken("n_cand", (3, 6), -1), "candidate count"),
             (broken("choice_type", 0, 22), "unknown location type"), (dict(acc=40), "unknown accommodation type")]
    home = int(city.home_loc[0])
    ty = int(np.nonzero(tables["n_cand"][home] > 0)[0][0])
    cases.append((broken("cand", (home, ty, 0), n_loc + 1), "candidate table"))
    for override, text in cases:
        rc, msg = _validate(city, tables, **override)
        assert rc == _lib.ERR_INVALID and text in msg, (override.keys(), rc, msg)
    # -1 = "none" stays legal everywhere it means that
    ok = {**broken("work_loc", worker, -1), **broken("nearest", (3, 6), -1)}
    assert _validate(city, tables, **ok)[0] == 0


def test_java_and_python_bindings_agree_with_the_header():
    """integration/java/eu/heros/gpu/HerosGpuNative.java (generated by tools/gen_java_binding.py; the Panama FFM handles a
    maintainer of the reference adds) is up to date with include/heros_gpu.h, and the ctypes table of the Python binding
    describes the same prototypes: same functions, same arity, pointers where the header has pointers, 32 / 64-bit
    integers and doubles where it has those."""
    import importlib.util
    import subprocess
    import sys
    gen = os.path.join(ROOT, "tools", "gen_java_binding.py")
    assert subprocess.run([sys.executable, gen, "--check"]).returncode == 0
    spec = importlib.util.spec_from_file_location("gen_java_binding", gen)
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    protos = g.prototypes(open(os.path.join(ROOT, "include", "heros_gpu.h")).read())
    assert sorted(n for _, n, _ in protos) == header_functions() == sorted(_lib.SIGNATURES)
    java = open(g.OUT).read()
    kind = {C.c_int32: "JAVA_INT", C.c_uint32: "JAVA_INT", C.c_int64: "JAVA_LONG", C.c_uint64: "JAVA_LONG",
            C.c_double: "JAVA_DOUBLE", C.c_int16: "JAVA_SHORT", C.c_float: "JAVA_FLOAT"}
    for ret, name, params in protos:
        restype, argtypes = _lib.SIGNATURES[name]
        assert len(argtypes) == len(params), name
        layouts = [g.layout(ret)] + [g.layout(t) for t, _ in params]
        assert f'fn("{name}",\n            FunctionDescriptor.of({", ".join(layouts)}));' in java, name
        for ctype, lay in zip([restype] + list(argtypes), layouts):
            got = kind.get(ctype, "ADDRESS")     # c_void_p, c_char_p, POINTER(...) are addresses
            assert got == lay, (name, ctype, lay)
    # the struct layouts: same sizes as the ctypes mirrors (natural alignment, no implicit padding), every field named
    all_structs = g.structs(open(os.path.join(ROOT, "include", "heros_gpu.h")).read())
    mirrors = dict(heros_config=_lib.Config, heros_area_params=_lib.AreaParams, heros_dist_params=_lib.DistParams,
                   heros_duration=_lib.Duration, heros_prog_params=_lib.ProgParams, heros_step_stats=_lib.StepStats,
                   heros_activity=_lib.Activity)
    assert set(all_structs) == set(mirrors)
    for name, mirror in mirrors.items():
        assert g.struct_size(name, all_structs) == C.sizeof(mirror), name
        assert f"{name.upper()}_LAYOUT = MemoryLayout.structLayout(" in java
        assert [f for f, _, _ in all_structs[name]] == [f for f, *_ in mirror._fields_], name


def test_hand_written_java_layer_uses_the_generated_handles_consistently():
    """integration/java/eu/heros/gpu/{HerosGpu, GpuCovid19Transmission, GpuCovid19Progression, GpuWindow}.java cannot be
    compiled here (no JDK); what can be checked is checked: every handle and layout they name exists in the generated
    HerosGpuNative.java, every invokeExact passes as many arguments as the C prototype has parameters, its int result is
    taken (invokeExact is signature-exact), and the struct fields addressed by name exist in the header's structs."""
    import importlib.util
    import re
    gen = os.path.join(ROOT, "tools", "gen_java_binding.py")
    spec = importlib.util.spec_from_file_location("gen_java_binding", gen)
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    header = open(os.path.join(ROOT, "include", "heros_gpu.h")).read()
    arity = {"heros_" + n[len("heros_"):]: len(params) for _, n, params in g.prototypes(header)}
    java_dir = os.path.dirname(g.OUT)
    native = open(g.OUT).read()
    handle_of = dict(re.findall(r"MethodHandle (\w+) = fn\(\"(\w+)\"", native))
    layouts = set(re.findall(r"StructLayout (\w+) =", native))
    fields = set(re.findall(r'withName\("(\w+)"\)', native))
    files = [f for f in sorted(os.listdir(java_dir)) if f.endswith(".java") and f != os.path.basename(g.OUT)]
    assert {"HerosGpu.java", "GpuCovid19Transmission.java", "GpuCovid19Progression.java", "GpuWindow.java"} <= set(files)
    calls = 0
    for name in files:
        src = open(os.path.join(java_dir, name)).read()
        for ref in re.findall(r"HerosGpuNative\.(\w+)", src):
            assert ref in handle_of or ref in layouts, (name, ref)
        for f in re.findall(r'groupElement\("(\w+)"\)', src) + re.findall(r'offset\("(\w+)"\)', src):
            assert f in fields, (name, f)
        for m in re.finditer(r"HerosGpuNative\.(\w+)\.invokeExact\(", src):
            depth, commas, i = 1, 0, m.end()
            empty = src[i:].lstrip().startswith(")")
            while depth:
                c = src[i]
                depth += c in "([{"
                depth -= c in ")]}"
                commas += c == "," and depth == 1
                i += 1
            n_args = 0 if empty else commas + 1
            assert n_args == arity[handle_of[m.group(1)]], (name, m.group(1), n_args)
            before = src[max(0, m.start() - 40):m.start()]
            assert re.search(r"\((int|MemorySegment)\)\s*$", before), (name, m.group(1), "result of invokeExact not taken")
            calls += 1
    assert calls >= 15
