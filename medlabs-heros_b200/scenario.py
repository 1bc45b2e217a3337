"""Synthetic populations and the host-side activity-schedule generator.

The reference's population blobs (people.csv.gz / locations.csv.gz of The Hague, Haaglanden, Europe) are not shipped
(/root/reference/.MISSING_LARGE_BLOBS), so every benchmark configuration runs on a synthetic city that follows the
documented schema (metadata/DataDictionary.xlsx; ConstructHerosModel.java:300-395, 450-758):

  * location mix, sub-location counts and per-sub-location areas scaled from the one real table that is shipped
    (data/helsinki/locations.csv.gz: 69 508 locations, 431 470 sub-locations; per-category statistics in CATEGORIES),
  * location type ids = row order of data/thehague/locations/locationtypes.csv,
  * social roles 1..15 by age (ConstructHerosModel.java:565-743), households mapped to Accommodation sub-locations
    (:531-552),
  * daily stays produced from the reference's own activity schedule rows (policy 0 of activityschedules_cap.xlsx,
    compiled to data/activity_patterns_thehague.json by tools/make_schedule_fixture.py) with counter-based Philox
    draws keyed by (seed, person, day, activity), so the stream is reproducible on any device.

This module is host-side set-up code (numpy); it stands in for the medlabs activity engine the Java host keeps.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))

# type ids = row order of data/thehague/locations/locationtypes.csv
TYPE_NAMES = ["Accommodation", "Workplace", "Retail", "Mall", "BarRestaurant", "FoodBeverage", "Supermarket",
              "Kindergarten", "PrimarySchool", "SecondarySchool", "College", "University", "Religion", "Police",
              "FireStation", "Pharmacy", "Healthcare", "Hospital", "Recreation", "Park", "Satellite accommodation",
              "Satellite workplace"]
TYPE_ID = {n: i for i, n in enumerate(TYPE_NAMES)}
# locationtypes.csv columns infectSub / contagiousRateFactor (Park .. Satellite* have infectSub = FALSE)
TYPE_INFECT_SUB = np.array([1] * 19 + [0, 0, 0], np.uint8)
TYPE_CORRECTION = np.ones(22, np.float64)
# columns capConstrained / capPersonsPerM2 / sizeFactor of the same file
TYPE_CAP_CONSTRAINED = np.array([0, 0, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 1, 1, 1, 1, 0, 0], np.uint8)
TYPE_CAP_PER_M2 = np.array([.25, .5, .25, .1, .5, .5, .2, .5, .5, .5, .5, .5, .5, .5, .5, .5, .25, .25, .2, .1, .25, .25])
TYPE_SIZE_FACTOR = np.array([1, 1, 1, 7.4, 1, 1.3, 9.7, 1, 1, 1, 1, 1, 1, 1, 1, 3.6, 1, 1, 1.9, 1, 1, 1], np.float64)
LOCATOR_CAP = 0x10  # flag on an activity's locator: the capacity-constrained form (NearestLocatorCap, RandomLocatorCap ...)

# (category, share of locations, sub-locations per location, m2 per sub-location) from data/helsinki/locations.csv.gz
CATEGORIES = [
    ("Workplace", 42169, None, 30.0), ("Accommodation", 20512, None, 100.0), ("BarRestaurant", 2255, 1, 200.0),
    ("Retail", 2239, 1, 100.0), ("Park", 506, 1, 300000.0), ("Kindergarten", 356, 6, 50.0),
    ("FoodBeverage", 355, 1, 100.0), ("Recreation", 293, 4, 250.0), ("Healthcare", 233, 3, 100.0),
    ("PrimarySchool", 173, 10, 75.0), ("Supermarket", 89, 1, 200.0), ("SecondarySchool", 74, 30, 75.0),
    ("Pharmacy", 67, 1, 75.0), ("Religion", 48, 1, 250.0), ("Mall", 37, 50, 100.0), ("University", 28, 100, 100.0),
    ("FireStation", 25, 10, 100.0), ("Hospital", 24, 650, 75.0), ("College", 22, 50, 100.0), ("Police", 3, 25, 30.0),
]
HELSINKI_LOCATIONS = 69508

# The free parameters of the synthetic population, fitted (tools/calibrate_city.py, profiles/r02_calibration.json) so that
# the area-model epidemic of the synthetic Hague with the reference's HardLockdown_20 policy follows the curves the
# reference publishes for seeds 111-114 (seed comparison.xlsx): persons per household, number of workplaces relative to
# the Helsinki table (0.23: about 14 workers per workplace instead of 3), mean number of rooms per workplace.
HAGUE_CALIBRATED = dict(household_size=2.2, workplace_scale=0.23, workplace_sub_mean=1.75)

ROLE_INFANT, ROLE_KINDERGARTEN, ROLE_PRIMARY, ROLE_SECONDARY, ROLE_COLLEGE, ROLE_UNIVERSITY, ROLE_WORKER, \
    ROLE_PENSIONER, ROLE_UNEMPLOYED, ROLE_WEEKEND_WORKER, ROLE_ESSENTIAL_WORKER = range(1, 12)

TAG_SCHEDULE = 0x300
NONE = -1


# ---------------------------------------------------------------------------------------------------------------------
# vectorised Philox4x32-10 (same function as csrc/heros_math.cuh)
# ---------------------------------------------------------------------------------------------------------------------
def philox4x32_10(c0, c1, c2, c3, k0, k1):
    u64, u32 = np.uint64, np.uint32
    c0 = np.asarray(c0, u32).astype(u64)
    shape = c0.shape
    c1 = np.broadcast_to(np.asarray(c1, u32), shape).astype(u64)
    c2 = np.broadcast_to(np.asarray(c2, u32), shape).astype(u64)
    c3 = np.broadcast_to(np.asarray(c3, u32), shape).astype(u64)
    k0 = u64(int(k0) & 0xFFFFFFFF)
    k1 = u64(int(k1) & 0xFFFFFFFF)
    m0, m1, mask = u64(0xD2511F53), u64(0xCD9E8D57), u64(0xFFFFFFFF)
    for _ in range(10):
        p0 = m0 * c0
        p1 = m1 * c2
        n0 = (p1 >> u64(32)) ^ c1 ^ k0
        n2 = (p0 >> u64(32)) ^ c3 ^ k1
        c0, c1, c2, c3 = n0, p1 & mask, n2, p0 & mask
        k0 = (k0 + u64(0x9E3779B9)) & mask
        k1 = (k1 + u64(0xBB67AE85)) & mask
    return c0, c1, c2, c3


def u01_from_words(x, y):
    return ((((x << np.uint64(32)) | y) >> np.uint64(11)).astype(np.float64)) * 2.0 ** -53


# ---------------------------------------------------------------------------------------------------------------------
# city
# ---------------------------------------------------------------------------------------------------------------------
class City:
    """Static tables of a synthetic city, in the layout heros_set_location_types / _locations / _persons take."""

    def __init__(self, n_persons: int, seed: int = 20240807, n_locations: Optional[int] = None,
                 box_m=(12000.0, 9000.0), origin_m=(0.0, 0.0), n_random_candidates: int = 8,
                 location_table: Optional[Dict[str, np.ndarray]] = None,
                 person_table: Optional[Dict[str, np.ndarray]] = None,
                 work_anywhere_fraction: float = 1.0, n_work_neighbours: int = 64,
                 household_size: float = 2.15, workplace_scale: float = 1.0, workplace_sub_mean: float = 1.75,
                 type_scale: Optional[Dict[str, float]] = None):
        """``location_table``: a real location table (type, nsub, area_per_sub, lat, lon -- e.g. the Helsinki fixture
        of tests/golden, or ``loaders.read_locations``) instead of the synthetic one.  ``person_table``: real persons
        (age, role, home_loc, home_sub, work_loc as location indices, optionally female and work_sub --
        ``loaders.read_people``) instead of the synthetic population; needs ``location_table``."""
        rng = np.random.default_rng(seed)
        # share of the workers whose workplace is drawn from the whole city; the others work at one of the
        # n_work_neighbours workplaces nearest to their home (1.0: the Hague-sized benchmark city, where everybody may work
        # anywhere; < 1: cities much larger than a commute, BASELINE configs[4])
        self.work_anywhere_fraction = float(work_anywhere_fraction)
        self.n_work_neighbours = int(n_work_neighbours)
        if person_table is not None:
            if location_table is None:
                raise ValueError("a person table refers to the rows of a location table")
            n_persons = len(person_table["age"])
        self.n_persons = int(n_persons)
        self.type_infect_sub = TYPE_INFECT_SUB.copy()
        self.type_correction = TYPE_CORRECTION.copy()
        # knobs of the synthetic population that the reference's missing people / location files would fix (persons per
        # household, number of workplaces and rooms per workplace relative to the Helsinki table, number of places of a type)
        self.calibration = dict(household_size=household_size, workplace_scale=workplace_scale,
                                workplace_sub_mean=workplace_sub_mean, type_scale=dict(type_scale or {}))
        n_households = max(1, int(round(n_persons / household_size)))
        if location_table is not None:
            # columns as the loader reads them (ConstructHerosModel.java:300-395); rows with nb_sublocations = 0 and
            # area = -1 are kept as they are
            loc_type = np.asarray(location_table["type"], np.uint8).copy()
            loc_nsub = np.asarray(location_table["nsub"], np.int64).copy()
            area_sub = np.asarray(location_table["area_per_sub"], np.float64).copy()
            n_loc = len(loc_type)
            lat, lon = np.asarray(location_table["lat"], np.float64), np.asarray(location_table["lon"], np.float64)
            x = (lon - lon.mean()) * 111320.0 * np.cos(np.deg2rad(lat.mean()))  # metres, equirectangular
            y = (lat - lat.mean()) * 110540.0
            acc = np.nonzero(loc_type == TYPE_ID["Accommodation"])[0]
            # households by the number of dwellings of the building; a few live in buildings listed with 0 sub-locations
            w = np.maximum(loc_nsub[acc], 0).astype(np.float64) + 0.05
            hh_building = acc[rng.choice(len(acc), n_households, p=w / w.sum())]
            hh_sub = rng.integers(0, 1 << 30, n_households) % np.maximum(loc_nsub[hh_building], 1)
        else:
            if n_locations is None:
                n_locations = int(round(143178 * n_persons / 553667))  # The Hague: 143 178 locations for 553 667 persons
            scale = n_locations / HELSINKI_LOCATIONS
            types, nsubs, per_sub_area = [], [], []
            for name, count, nsub, area in CATEGORIES:
                k = max(1, int(round(count * scale * (type_scale or {}).get(name, 1.0))))
                if name == "Accommodation":
                    k = max(1, min(k, n_households))
                    sub = np.zeros(k, np.int64)  # filled below from the households
                elif name == "Workplace":
                    k = max(1, int(round(k * workplace_scale)))
                    sub = np.minimum(1 + np.floor(rng.exponential(workplace_sub_mean, k)).astype(np.int64), 100)
                else:
                    sub = np.full(k, nsub, np.int64)
                types.append(np.full(k, TYPE_ID[name], np.uint8))
                nsubs.append(sub)
                per_sub_area.append(np.full(k, area))
            loc_type = np.concatenate(types)
            loc_nsub = np.concatenate(nsubs)
            area_sub = np.concatenate(per_sub_area)
            n_loc = len(loc_type)
            x = origin_m[0] + rng.random(n_loc) * box_m[0]
            y = origin_m[1] + rng.random(n_loc) * box_m[1]

            # households -> Accommodation sub-locations in order of first appearance (ConstructHerosModel.java:531-552)
            acc = np.nonzero(loc_type == TYPE_ID["Accommodation"])[0]
            hh_building = acc[rng.integers(0, len(acc), n_households)]
            order = np.argsort(hh_building, kind="stable")
            hh_sub = np.zeros(n_households, np.int64)
            sorted_b = hh_building[order]
            first = np.r_[True, sorted_b[1:] != sorted_b[:-1]]
            start = np.maximum.accumulate(np.where(first, np.arange(n_households), 0))
            hh_sub[order] = np.arange(n_households) - start
            np.add.at(loc_nsub, hh_building, 1)
            loc_nsub[acc] = np.maximum(loc_nsub[acc], 1)

        if person_table is not None:
            age = np.asarray(person_table["age"], np.int64)
            role = np.asarray(person_table["role"], np.int64)
            home_loc = np.asarray(person_table["home_loc"], np.int64)
            home_sub = np.asarray(person_table["home_sub"], np.int64)
            work_loc = np.asarray(person_table["work_loc"], np.int64)
            self.household = np.asarray(person_table.get("household", np.zeros(n_persons)), np.int32)
            if "work_sub" in person_table:
                work_sub = np.asarray(person_table["work_sub"], np.int64)
            else:
                # where in the work / school building a person sits is medlabs' choice (not visible in the reference):
                # a keyed hash of the person, fixed for the run
                hx = philox4x32_10(np.arange(n_persons), 0, 3, TAG_SCHEDULE, seed & 0xFFFFFFFF, seed >> 32)[0]
                work_sub = np.where(work_loc >= 0, hx.astype(np.int64) % np.maximum(loc_nsub[np.maximum(work_loc, 0)], 1), 0)
        else:
            age, role, home_loc, home_sub, work_loc, work_sub = self._synthetic_persons(
                rng, n_persons, n_households, hh_building, hh_sub, loc_type, loc_nsub, x, y)
        from scipy.spatial import cKDTree
        # nearest / random candidate tables of EVERY location: the reference builds NearestLocator / RandomLocator over a
        # CurrentLocator (ConstructHerosModel.java:997, 1012) -- nearest to, or within the maximum distance of, the place
        # the person is in when the activity starts (its home, its workplace, the shop it just left ...)
        self.n_random_candidates = K = n_random_candidates
        n_types = len(TYPE_NAMES)
        nearest = np.full((n_loc, n_types), NONE, np.int32)
        cand = np.full((n_loc, n_types, K), NONE, np.int32)
        n_cand = np.zeros((n_loc, n_types), np.int32)
        pts = np.c_[x, y]
        for t in range(n_types):
            places = np.nonzero(loc_type == t)[0]
            if t in (TYPE_ID["Accommodation"], TYPE_ID["Workplace"]) or len(places) == 0:
                continue
            tree = cKDTree(np.c_[x[places], y[places]])
            kk = min(K, len(places))
            d, j = tree.query(pts, k=kk)
            d = d.reshape(n_loc, kk)
            j = j.reshape(n_loc, kk)
            nearest[:, t] = places[j[:, 0]]
            ok = d <= 2500.0  # RandomLocator max distance (xlsx column P)
            n_cand[:, t] = ok.sum(axis=1)
            cand[:, t, :kk] = np.where(ok, places[j], NONE)
        self.nearest, self.cand, self.n_cand = nearest, cand, n_cand

        self.loc_type = loc_type.astype(np.uint8)
        self.loc_nsub = loc_nsub.astype(np.int16)
        self.loc_area = (area_sub * loc_nsub).astype(np.float32)  # totalArea = area * nb_sublocations (java :381)
        self.loc_x, self.loc_y = x.astype(np.float32), y.astype(np.float32)
        self.n_locations = n_loc
        self.age = age.astype(np.int8)
        self.female = (rng.random(n_persons) < 0.5).astype(np.uint8)  # gender drawn U01 < 0.5 (java :513)
        if person_table is not None and "female" in person_table:
            self.female = np.asarray(person_table["female"], np.uint8)
        self.role = role.astype(np.uint8)
        self.home_loc = home_loc.astype(np.int32)
        self.home_sub = home_sub.astype(np.int16)
        self.work_loc = work_loc.astype(np.int32)
        self.work_sub = work_sub.astype(np.int16)

    def _synthetic_persons(self, rng, n_persons, n_households, hh_building, hh_sub, loc_type, loc_nsub, x, y):
        """ages, roles, households, work / school places of the synthetic population"""
        age = self._draw_ages(rng, n_persons)
        role = self._roles(rng, age)
        adults = np.nonzero(age >= 18)[0]
        rng.shuffle(adults)
        hh = np.empty(n_persons, np.int64)
        heads = adults[:min(len(adults), n_households)]
        hh[heads] = np.arange(len(heads))
        rest = np.setdiff1d(np.arange(n_persons), heads, assume_unique=False)
        hh[rest] = rng.integers(0, n_households, len(rest))
        self.household = hh.astype(np.int32)
        home_loc = hh_building[hh]
        home_sub = hh_sub[hh]

        # work / school places
        work_loc = np.full(n_persons, NONE, np.int64)
        work_sub = np.zeros(n_persons, np.int64)
        wp = np.nonzero(loc_type == TYPE_ID["Workplace"])[0]
        workers = np.nonzero(np.isin(role, (ROLE_WORKER, ROLE_WEEKEND_WORKER, ROLE_ESSENTIAL_WORKER)))[0]
        weights = np.maximum(loc_nsub[wp], 1) / np.maximum(loc_nsub[wp], 1).sum()
        work_loc[workers] = wp[rng.choice(len(wp), len(workers), p=weights)]
        from scipy.spatial import cKDTree
        if self.work_anywhere_fraction < 1.0 and len(workers) and len(wp):
            rng_w = np.random.default_rng(rng.integers(0, 1 << 62))  # its own stream: the draws above stay what they were
            near = workers[rng_w.random(len(workers)) >= self.work_anywhere_fraction]
            k = min(self.n_work_neighbours, len(wp))
            _, j = cKDTree(np.c_[x[wp], y[wp]]).query(np.c_[x[home_loc[near]], y[home_loc[near]]], k=k)
            j = j.reshape(len(near), k)
            work_loc[near] = wp[j[np.arange(len(near)), rng_w.integers(0, k, len(near))]]
        for r, tname in ((ROLE_KINDERGARTEN, "Kindergarten"), (ROLE_PRIMARY, "PrimarySchool"),
                         (ROLE_SECONDARY, "SecondarySchool")):
            who = np.nonzero(role == r)[0]
            places = np.nonzero(loc_type == TYPE_ID[tname])[0]
            if len(who) and len(places):
                _, j = cKDTree(np.c_[x[places], y[places]]).query(np.c_[x[home_loc[who]], y[home_loc[who]]])
                work_loc[who] = places[j]
        for r, tname in ((ROLE_COLLEGE, "College"), (ROLE_UNIVERSITY, "University")):
            who = np.nonzero(role == r)[0]
            places = np.nonzero(loc_type == TYPE_ID[tname])[0]
            if len(who) and len(places):
                work_loc[who] = places[rng.integers(0, len(places), len(who))]
        has = work_loc >= 0
        work_sub[has] = rng.integers(0, 1 << 30, has.sum()) % np.maximum(loc_nsub[work_loc[has]], 1)

        return age, role, home_loc, home_sub, work_loc, work_sub

    @staticmethod
    def _draw_ages(rng, n):
        # coarse Dutch age pyramid
        edges = np.array([0, 4, 6, 12, 18, 23, 30, 40, 50, 60, 67, 75, 85, 96])
        share = np.array([.040, .022, .066, .070, .062, .095, .125, .130, .140, .085, .080, .060, .025])
        band = rng.choice(len(share), n, p=share / share.sum())
        return (edges[band] + np.floor(rng.random(n) * (edges[band + 1] - edges[band]))).astype(np.int64)

    @staticmethod
    def _roles(rng, age):
        n = len(age)
        u = rng.random(n)
        role = np.full(n, ROLE_PENSIONER, np.int64)
        role[age < 4] = ROLE_INFANT
        role[(age >= 4) & (age < 6)] = ROLE_KINDERGARTEN
        role[(age >= 6) & (age < 12)] = ROLE_PRIMARY
        role[(age >= 12) & (age < 18)] = ROLE_SECONDARY
        young = (age >= 18) & (age < 23)
        role[young] = np.select([u[young] < .3, u[young] < .6], [ROLE_COLLEGE, ROLE_UNIVERSITY], ROLE_WORKER)
        adult = (age >= 23) & (age < 67)
        role[adult] = np.select([u[adult] < .62, u[adult] < .74, u[adult] < .82],
                                [ROLE_WORKER, ROLE_UNEMPLOYED, ROLE_WEEKEND_WORKER], ROLE_ESSENTIAL_WORKER)
        return role

    def add_commuters(self, other: "City", fraction: float, seed: int = 77) -> int:
        """Gives ``fraction`` of the workers of this city tile a workplace in the neighbouring tile ``other`` (cross-shard
        travel of a geographically sharded run).  A remote workplace is coded as ``n_locations + its id in
        the other tile``; ``global_location_ids`` turns the codes into global ids.  Returns the number of commuters."""
        rng = np.random.default_rng(seed)
        workers = np.nonzero(self.role == ROLE_WORKER)[0]
        pick = workers[rng.random(len(workers)) < fraction]
        wp = np.nonzero(other.loc_type == TYPE_ID["Workplace"])[0]
        weights = other.loc_nsub[wp].astype(np.float64)
        dest = wp[rng.choice(len(wp), len(pick), p=weights / weights.sum())]
        self.work_loc[pick] = (self.n_locations + dest).astype(np.int32)
        self.work_sub[pick] = (rng.integers(0, 1 << 30, len(pick)) % other.loc_nsub[dest]).astype(np.int16)
        self.n_commuters = len(pick)
        return len(pick)

    def global_location_ids(self, loc: np.ndarray, own_base: int, other_base: Optional[int] = None) -> np.ndarray:
        """Location codes of the schedule generator -> global location ids of a sharded run.  Codes beyond the city's own
        table are foreign locations: entries of ``ext_global`` (a shard of ``partition_city``), or -- ``add_commuters`` --
        locations of the one neighbouring tile whose first global id is ``other_base``."""
        loc = loc.astype(np.int64)
        ext = getattr(self, "ext_global", None)
        if ext is not None:
            far = loc >= self.n_locations
            return np.where(far, ext[np.where(far, loc - self.n_locations, 0)], loc + own_base).astype(np.int32)
        if other_base is None:  # a city on its own: every location is its own
            return (loc + own_base).astype(np.int32)
        return np.where(loc >= self.n_locations, loc - self.n_locations + other_base, loc + own_base).astype(np.int32)

    def install(self, engine):
        """heros_set_location_types / heros_set_locations / heros_set_persons (a shard: its own locations and persons)."""
        n = self.n_locations
        engine.set_location_types(self.type_infect_sub, self.type_correction)
        engine.set_locations(self.loc_type[:n], self.loc_nsub[:n], self.loc_area[:n], self.loc_y[:n], self.loc_x[:n])
        engine.set_persons(self.age, self.female, self.role, self.home_loc, self.home_sub, self.work_loc)


# ---------------------------------------------------------------------------------------------------------------------
# activity schedules -> stays
# ---------------------------------------------------------------------------------------------------------------------
def load_patterns(path: Optional[str] = None) -> dict:
    path = path or os.path.join(PKG_DIR, "data", "activity_patterns_thehague.json")
    with open(path) as f:
        return json.load(f)


ACTIVITY_DTYPE = np.dtype([("kind", np.int32), ("locator", np.int32), ("choice", np.int32), ("dist", np.int32),
                           ("min", np.float64), ("mode", np.float64), ("max", np.float64), ("until", np.float64)])
DAY_TYPES = ("Monday", "Friday", "Saturday", "Sunday")  # Mon-Thu share the Monday pattern


def compile_schedule(city: "City", patterns: Optional[dict] = None) -> Dict[str, np.ndarray]:
    """The activity patterns and locator tables of a city as the flat arrays heros_set_schedule takes: the same rows the
    host-side ScheduleGenerator interprets, so that the device-side schedule (heros_advance_schedule) produces the same
    stays bit for bit."""
    pat = patterns or load_patterns()
    acts, start, count = [], np.zeros((8, 16, 4), np.int32), np.zeros((8, 16, 4), np.int32)
    for ph, ph_name in enumerate(pat["phases"]):
        for r, role_name in enumerate(pat["roles"], start=1):
            for d, day in enumerate(DAY_TYPES):
                rows = pat["patterns"].get(f"{ph_name}|{role_name}|{day}") or pat["patterns"].get(f"{ph_name}|{role_name}|Monday")
                if not rows:
                    continue
                start[ph, r, d], count[ph, r, d] = len(acts), len(rows)
                for a in rows:
                    acts.append((a["kind"], a["locator"], a["choice"], a["dist"], a["min"], a["mode"], a["max"], a["until"]))
    c_start, c_type, c_cum = [0], [], []
    for items in pat["choices"]:
        c_type += [TYPE_ID[name] for name, _ in items]
        c_cum += np.cumsum([p for _, p in items]).tolist()
        c_start.append(len(c_type))
    return dict(activities=np.array(acts, ACTIVITY_DTYPE), pattern_start=start.ravel(), pattern_count=count.ravel(),
                choice_start=np.array(c_start, np.int32), choice_type=np.array(c_type, np.int32),
                choice_cum=np.array(c_cum, np.float64), n_types=len(TYPE_NAMES), accommodation_type=TYPE_ID["Accommodation"],
                n_candidates=city.n_random_candidates, nearest=city.nearest, n_cand=city.n_cand, cand=city.cand,
                work_sub=city.work_sub, divert=getattr(city, "divert", None))


def location_capacity(city: "City") -> np.ndarray:
    """Persons a location holds at the same time: capPersonsPerM2 x totalArea x sizeFactor of its type where the type is
    capConstrained (data/<city>/locations/locationtypes.csv columns 8-10), else unlimited."""
    ty = city.loc_type.astype(np.int64)
    cap = TYPE_CAP_PER_M2[ty] * city.loc_area.astype(np.float64) * TYPE_SIZE_FACTOR[ty]
    return np.where(TYPE_CAP_CONSTRAINED[ty] == 1, np.maximum(cap, 1.0), np.inf)


def capacity_diversion(city: "City", seed: int, patterns: Optional[dict] = None) -> np.ndarray:
    """The capacity-constrained locators (NearestLocatorCap / RandomLocatorCap and their Choice forms,
    ConstructHerosModel.java:998-1020) act inside medlabs; their rule is not in the reference tree.  Convention here,
    shared by the host generator and the device-side schedule and free of any dependence on the order in which persons are
    processed: one ordinary Monday of the healthy population is generated WITHOUT constraint and the largest number of
    persons present at the same time is taken for every location (peak); a constrained location with peak > capacity
    sheds the share 1 - capacity / peak of the visits a ``*Cap`` locator sends to it -- which visits is decided by a
    keyed uniform per (person, day, activity) -- to the next candidate of the same search (next-nearest place of the
    type; the next of the random candidates).  Returns that share per location; ``city.divert`` is set to it."""
    city.divert = np.zeros(len(city.loc_type), np.float64)
    gen = ScheduleGenerator(city, seed, patterns)
    st = gen.generate_day(np.zeros(city.n_persons, np.uint8))
    cap = location_capacity(city)
    sel = np.isfinite(cap[st["loc"]])
    loc = st["loc"][sel].astype(np.int64)
    t_in, t_out = st["t_in"][sel], np.where(np.isinf(st["t_out"][sel]), 1e9, st["t_out"][sel])
    ev_loc = np.concatenate([loc, loc])
    ev_t = np.concatenate([t_in, t_out])
    ev_d = np.concatenate([np.ones(len(loc), np.int64), -np.ones(len(loc), np.int64)])
    order = np.lexsort((ev_d, ev_t, ev_loc))  # by location, time, leaving before entering
    run = np.cumsum(ev_d[order])              # every location's events sum to zero: the count restarts by itself
    peak = np.zeros(len(city.loc_type), np.float64)
    np.maximum.at(peak, ev_loc[order], run.astype(np.float64))
    with np.errstate(divide="ignore", invalid="ignore"):
        divert = np.where(peak > cap, 1.0 - cap / np.maximum(peak, 1.0), 0.0)
    city.divert = divert
    city.peak_unconstrained = peak
    return divert


class ScheduleGenerator:
    """Produces, day by day, the stays (person, location, sub-location, t_in, t_out) of every person from the week
    patterns ``0_<phase>_<role>`` (HerosModel.java:484-504; Mon-Thu use the Monday pattern, ConstructHerosModel.java
    :931-940).  The phase that selects the pattern is the one held at midnight; a pattern switch therefore takes effect
    at the next day boundary (driver contract C15 is not observable in the reference; this is the convention shared
    with the windowed engine).  Travel time is spent outside every location; consecutive activities at the
    same place form one stay (contract C7)."""

    def __init__(self, city: City, seed: int, patterns: Optional[dict] = None):
        self.city = city
        self.seed = int(seed)
        self.pat = patterns or load_patterns()
        n = city.n_persons
        # everybody is at home from t = 0 (open stay)
        self.cur_loc = city.home_loc.astype(np.int64).copy()
        self.cur_sub = city.home_sub.astype(np.int64).copy()
        self.cur_tin = np.zeros(n, np.float64)
        self.open_reported = np.zeros(n, bool)
        self.day = 0
        self._choice_tables = [self._compile_choice(c) for c in self.pat["choices"]]
        self._day_names = ["Monday"] * 4 + ["Friday", "Saturday", "Sunday"]
        # closure policy per location type: fraction of its locations that stay open, fraction of the activities that
        # still take place in an open one (LocationPolicy.java:39-46)
        self.frac_open = np.ones(len(TYPE_NAMES), np.float64)
        self.frac_act = np.ones(len(TYPE_NAMES), np.float64)
        # a shard of partition_city draws with the GLOBAL ids of its persons and locations, so that the stays of a person
        # do not depend on how the city was cut
        self.person_key = getattr(city, "person_gid", None)
        self.loc_key = getattr(city, "loc_gid", None)

    def _pkey(self, idx):
        return idx if self.person_key is None else self.person_key[idx]

    def set_closure_policy(self, location_type: int, fraction_open: float, fraction_activities: float,
                           alternative_type: int) -> None:
        """``LocationType.setClosurePolicy(fractionOpen, fractionActivities, alternativeLocationType, reportAs)`` as the
        reference's policy files use it (LocationPolicy.java:45; data/*/policies/*.csv: the alternative is always
        Accommodation = the person's home, or the type itself = reopened).  medlabs' own rule is not in the reference
        tree; the convention shared with the device-side schedule is: a location of the type is open iff its keyed
        uniform (seed, location) is < fractionOpen (the same locations stay open while the fraction does not change, and
        a smaller fraction keeps a subset); an activity that would take place in an open location takes place iff its
        keyed uniform (seed, person, day, activity) is < fractionActivities; otherwise the person goes home.  A change
        applies to the days generated after it."""
        ty = int(location_type)
        if not 0 <= ty < len(TYPE_NAMES):
            raise ValueError("unknown location type")
        if int(alternative_type) == ty:
            fraction_open = fraction_activities = 1.0  # "Workplace -> Workplace": reopened
        elif int(alternative_type) != TYPE_ID["Accommodation"]:
            raise ValueError("the alternative of a closed location type is the person's home (Accommodation)")
        if not (0.0 <= fraction_open <= 1.0 and 0.0 <= fraction_activities <= 1.0):
            raise ValueError("fractions must lie in [0, 1]")
        self.frac_open[ty], self.frac_act[ty] = float(fraction_open), float(fraction_activities)

    def _closed(self, idx, day_act, dl, ty):
        """True where the activity of persons ``idx`` at location ``dl`` (type ``ty``, arrays) falls to the closure
        policy of the type."""
        ty = np.asarray(ty, np.int64)
        fo, fa = self.frac_open[ty], self.frac_act[ty]
        if not ((fo < 1.0) | (fa < 1.0)).any():
            return np.zeros(len(idx), bool)
        lk = np.maximum(dl, 0) if self.loc_key is None else self.loc_key[np.maximum(dl, 0)]
        x_loc = philox4x32_10(lk, 0, 2, TAG_SCHEDULE, self.seed, self.seed >> 32)[0]
        x_act = philox4x32_10(self._pkey(idx), day_act, 1, TAG_SCHEDULE, self.seed, self.seed >> 32)[0]
        u_loc = x_loc.astype(np.float64) * 2.0 ** -32
        u_act = x_act.astype(np.float64) * 2.0 ** -32
        return (dl >= 0) & ~((u_loc < fo) & (u_act < fa))

    @staticmethod
    def _compile_choice(items):
        types = np.array([TYPE_ID[name] for name, _ in items], np.int64)
        cum = np.cumsum([p for _, p in items])
        return types, cum

    def _pattern(self, phase_name: str, role_name: str, weekday: int):
        p = self.pat["patterns"]
        key = f"{phase_name}|{role_name}|{self._day_names[weekday]}"
        if key not in p:
            key = f"{phase_name}|{role_name}|Monday"
        return p.get(key)

    def generate_day(self, phase: np.ndarray) -> Dict[str, np.ndarray]:
        """Stays closed during day ``self.day`` plus the stays opened that day and still open at midnight
        (t_out = +inf).  ``phase``: disease phase of every person at the start of the day."""
        c = self.city
        d = self.day
        t0, t1 = 24.0 * d, 24.0 * (d + 1)
        weekday = d % 7
        out_p, out_l, out_s, out_a, out_b = [], [], [], [], []

        def emit(idx, t_leave):
            """close the current stay of persons idx at t_leave"""
            real = self.cur_loc[idx] >= 0
            idx, t_leave = idx[real], t_leave[real]
            ok = t_leave > self.cur_tin[idx]
            out_p.append(idx[ok]); out_l.append(self.cur_loc[idx[ok]]); out_s.append(self.cur_sub[idx[ok]])
            out_a.append(self.cur_tin[idx[ok]]); out_b.append(t_leave[ok])
            self.open_reported[idx] = False

        roles = self.pat["roles"]
        phases = self.pat["phases"]
        group = phase.astype(np.int64) * 16 + c.role.astype(np.int64)
        for g in np.unique(group):
            ph, role = divmod(int(g), 16)
            idx = np.nonzero(group == g)[0]
            if ph == 6:  # Dead: the person leaves the model
                emit(idx, np.full(len(idx), t0))
                self.cur_loc[idx] = NONE
                continue
            if role < 1 or role > len(roles):
                continue
            acts = self._pattern(phases[ph], roles[role - 1], weekday)
            if not acts:
                continue
            t = np.full(len(idx), t0)
            alive = np.ones(len(idx), bool)  # still before midnight
            for a_i, act in enumerate(acts):
                if not alive.any():
                    break
                x, y, z, w = philox4x32_10(self._pkey(idx), d * 64 + a_i, 0, TAG_SCHEDULE, self.seed, self.seed >> 32)
                u_dur = u01_from_words(x, y)
                u_type = z.astype(np.float64) * 2.0 ** -32
                u_cand = (w >> np.uint64(16)).astype(np.float64) * 2.0 ** -16
                u_sub = (w & np.uint64(0xFFFF)).astype(np.float64) * 2.0 ** -16
                dest_l, dest_s = self._locate(act, idx, u_type, u_cand, u_sub, d * 64 + a_i)
                move = alive & ((dest_l != self.cur_loc[idx]) | (dest_s != self.cur_sub[idx]))
                dur = self._duration(act, u_dur)
                if act["kind"] == 1:  # travel: leave now, arrive after the travel time
                    mv = np.nonzero(move)[0]
                    emit(idx[mv], t[mv])
                    t_arr = t[mv] + dur[mv]
                    arrived = t_arr < t1
                    self.cur_loc[idx[mv]] = np.where(arrived, dest_l[mv], NONE)
                    self.cur_sub[idx[mv]] = np.where(arrived, dest_s[mv], 0)
                    self.cur_tin[idx[mv]] = t_arr
                    t[mv] = t_arr
                    alive[mv] &= arrived
                else:
                    mv = np.nonzero(move)[0]
                    emit(idx[mv], t[mv])
                    self.cur_loc[idx[mv]] = dest_l[mv]
                    self.cur_sub[idx[mv]] = dest_s[mv]
                    self.cur_tin[idx[mv]] = t[mv]
                    if act["kind"] == 0:
                        t = np.where(alive, t + dur, t)
                    else:
                        t = np.where(alive, np.maximum(t, t0 + act["until"]), t)
                    alive &= t < t1
        # stays opened today that are still open at midnight
        opened = np.nonzero((self.cur_loc >= 0) & (self.cur_tin >= t0) & (self.cur_tin < t1) & ~self.open_reported)[0]
        if d == 0:
            opened = np.nonzero((self.cur_loc >= 0) & (self.cur_tin < t1) & ~self.open_reported)[0]
        out_p.append(opened); out_l.append(self.cur_loc[opened]); out_s.append(self.cur_sub[opened])
        out_a.append(self.cur_tin[opened]); out_b.append(np.full(len(opened), np.inf))
        self.open_reported[opened] = True
        self.day += 1
        return dict(person=np.concatenate(out_p).astype(np.int32), loc=np.concatenate(out_l).astype(np.int32),
                    sub=np.concatenate(out_s).astype(np.int16), t_in=np.concatenate(out_a).astype(np.float64),
                    t_out=np.concatenate(out_b).astype(np.float64))

    def _duration(self, act, u):
        if act["dist"] == 2:
            mn, mode, mx = act["min"], act["mode"], act["max"]
            if mx <= mn:
                return np.full(len(u), mn)
            fc = (mode - mn) / (mx - mn)
            lo = mn + np.sqrt((mode - mn) * (mx - mn) * u)
            hi = mx - np.sqrt((mx - mn) * (mx - mode) * (1.0 - u))
            return np.where(u <= fc, lo, hi)
        if act["dist"] == 1:
            return act["min"] + (act["max"] - act["min"]) * u
        return np.full(len(u), act["mode"])

    def _locate(self, act, idx, u_type, u_cand, u_sub, day_act=0):
        c = self.city
        loc = act["locator"] & 0xF
        capped = bool(act["locator"] & LOCATOR_CAP)
        home_l, home_s = c.home_loc[idx].astype(np.int64), c.home_sub[idx].astype(np.int64)
        if loc == 0:
            return home_l, home_s
        if loc == 1:  # Work / School; persons without one stay where they are
            wl = c.work_loc[idx].astype(np.int64)
            has = wl >= 0
            # a closure policy applies to the work / school locations of this city's own table (commuters into another
            # tile, add_commuters, carry ids beyond it)
            own = has & (wl < len(c.loc_type))
            shut = np.zeros(len(idx), bool)
            if ((self.frac_open < 1.0) | (self.frac_act < 1.0)).any():
                shut = own & self._closed(idx, day_act, np.where(own, wl, 0), c.loc_type[np.where(own, wl, 0)])
            dl = np.where(has, wl, self.cur_loc[idx])
            ds = np.where(has, c.work_sub[idx].astype(np.int64), self.cur_sub[idx])
            return np.where(shut, home_l, dl), np.where(shut, home_s, ds)
        if loc == 2 or loc == 5 or act["choice"] < 0:
            return self.cur_loc[idx].copy(), self.cur_sub[idx].copy()
        types, cum = self._choice_tables[act["choice"]]
        pick = np.minimum(np.searchsorted(cum, u_type * cum[-1], side="right"), len(types) - 1)
        ty = types[pick]
        # the anchor of the search: the place the person is in (CurrentLocator); somebody on the way -- in no location --
        # searches from home
        here = self.cur_loc[idx]
        anchor = np.where(here >= 0, here, home_l)
        n = c.n_cand[anchor, ty]
        if loc == 3:  # Nearest(type) to the current place
            dl = c.nearest[anchor, ty].astype(np.int64)
            j = np.zeros(len(idx), np.int64)
        else:  # Random(type) within the maximum distance of the current place
            j = np.minimum((u_cand * n).astype(np.int64), np.maximum(n - 1, 0))
            dl = np.where(n > 0, c.cand[anchor, ty, j], NONE).astype(np.int64)
        divert = getattr(c, "divert", None)
        if capped and divert is not None and divert.any():
            # capacity-constrained form: a keyed share of the visits to an over-subscribed place goes to the next
            # candidate of the search (capacity_diversion)
            u_cap = philox4x32_10(self._pkey(idx), day_act, 4, TAG_SCHEDULE, self.seed, self.seed >> 32)[0].astype(np.float64) * 2.0 ** -32
            shed = (dl >= 0) & (u_cap < divert[np.maximum(dl, 0)]) & (n >= 2)
            j2 = np.where(loc == 3, 1, (j + 1) % np.maximum(n, 1))
            alt = c.cand[anchor, ty, np.minimum(j2, c.cand.shape[2] - 1)].astype(np.int64)
            dl = np.where(shed & (alt >= 0), alt, dl)
        go_home = (ty == TYPE_ID["Accommodation"]) | (dl < 0)
        go_home |= self._closed(idx, day_act, dl, ty)
        nsub = np.maximum(c.loc_nsub[np.maximum(dl, 0)].astype(np.int64), 1)
        ds = np.minimum((u_sub * nsub).astype(np.int64), nsub - 1)
        return np.where(go_home, home_l, dl), np.where(go_home, home_s, ds)
