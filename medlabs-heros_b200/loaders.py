"""Readers for the input files of a reference scenario (``data/<city>/``), producing the tables the engine takes.

The Java host keeps its own loaders (factory/ConstructHerosModel.java); these are the same rules for a host without a
JVM, so that the examples and tools can be pointed at a ``data/<city>`` directory:

  * ``locationtypes.csv``          -> ``properties.load_location_types``            (java :230-269)
  * ``locations.csv[.gz]``         -> ``read_locations``                             (java :300-395)
  * ``people.csv[.gz]``            -> ``read_people``                                (java :450-758)
  * ``activities/*.xlsx``          -> ``read_activity_schedules``                    (java :763-942)
  * both                            -> ``load_city`` = a ``scenario.City`` with real tables (candidate / nearest tables of
                                      the locators are built as for the synthetic city)

Rows the reference rejects are rejected here with the same message on stderr and skipped, not raised (java prints and
``continue``s).  Ids: the engine wants dense indices; location and person ids of the files are kept beside them
(``location_id`` / ``person_id``) for the outputs.
"""
from __future__ import annotations

import csv
import gzip
import sys
from typing import Dict, Optional, Sequence

import numpy as np

from . import scenario
from .xlsx import read_xlsx

STUDENT_SCHOOL = {2: "kindergarten", 3: "primaryschool", 4: "secondaryschool", 5: "college", 6: "university"}
ROLE_LABEL = {7: "Worker", 10: "WeekendWorker", 11: "EssentialWorker", 12: "WorkerSatelliteToCity",
              13: "WorkerCityToSatellite", 14: "WorkerSatelliteToSatellite", 15: "WorkerCountryToCity"}


def _open(path: str):
    return gzip.open(path, "rt", newline="") if path.endswith(".gz") else open(path, newline="")


def _columns(header: Sequence[str], wanted: Sequence[str], what: str) -> Dict[str, int]:
    """case-insensitive header lookup; all columns must be there (java :317-348, :464-493)"""
    low = [h.strip().lower() for h in header]
    missing = [w for w in wanted if w not in low]
    if missing:
        raise ValueError(f"{what} csv-file header row did not contain all column headers\n{list(header)}")
    return {w: low.index(w) for w in wanted}


def read_locations(path: str, type_names: Optional[Sequence[str]] = None) -> Dict[str, np.ndarray]:
    """``locations.csv[.gz]``: columns location_id, nb_sublocations, location_category, lat, lon, area (java :329-348);
    ``area`` is the area of one sub-location, the total surface is ``area * nb_sublocations`` as float (java :381).
    Rows whose category is not a location type are reported and skipped.  Returns the columns ``scenario.City`` takes
    (type, nsub, area_per_sub, lat, lon) plus ``location_id``; rows keep the order of the file."""
    names = list(type_names if type_names is not None else scenario.TYPE_NAMES)
    type_id = {n: i for i, n in enumerate(names)}
    with _open(path) as f:
        rows = list(csv.reader(f))
    col = _columns(rows[0], ("location_id", "nb_sublocations", "location_category", "lat", "lon", "area"), "Location")
    ids, types, nsub, area, lat, lon = [], [], [], [], [], []
    for line, r in enumerate(rows[1:], start=2):
        if not r:
            continue
        cat = r[col["location_category"]]
        if cat not in type_id:
            print(f"Location type {cat} not found on line {line}\n{r}", file=sys.stderr)
            continue
        ids.append(int(r[col["location_id"]]))
        types.append(type_id[cat])
        nsub.append(int(float(r[col["nb_sublocations"]])))
        area.append(float(r[col["area"]]))
        lat.append(float(r[col["lat"]]))
        lon.append(float(r[col["lon"]]))
    return dict(location_id=np.array(ids, np.int64), type=np.array(types, np.uint8), nsub=np.array(nsub, np.int16),
                area_per_sub=np.array(area, np.float32), lat=np.array(lat, np.float32), lon=np.array(lon, np.float32))


def read_people(path: str, locations: Dict[str, np.ndarray], type_names: Optional[Sequence[str]] = None,
                seed: int = 111) -> Dict[str, np.ndarray]:
    """``people.csv[.gz]``: columns person_id, household_id, age, home_id, workplace_id, social_role (java :464-493),
    all parsed through double as the reference does (java :499-504).

    Rules of the reference loader, in its order: an age outside 0..120 is reported (the person is kept, java :506-510);
    a home that is not in the location map drops the person (java :514-520); the home sub-location index is the order
    of first appearance of the household id within its home building, reported when it exceeds the building's number
    of sub-locations (java :531-552); students (roles 2-6) need a school in the map (java :554-562) and of the matching
    type for roles 2, 3, 5, 6 (java :572-640; the check of role 4 is the same with "secondaryschool"); roles 7 and 10-15
    need a work location in the map (java :642-745); unknown roles are reported and dropped.  The gender is drawn
    ``U01 < 0.5`` per person (java :513) -- here a Philox number keyed by (seed, person index) instead of the model's
    Mersenne Twister stream."""
    names = [n.lower() for n in (type_names if type_names is not None else scenario.TYPE_NAMES)]
    index_of = {int(i): k for k, i in enumerate(locations["location_id"])}
    loc_type, loc_nsub = locations["type"], locations["nsub"]
    with _open(path) as f:
        rows = list(csv.reader(f))
    col = _columns(rows[0], ("person_id", "household_id", "age", "home_id", "workplace_id", "social_role"), "Person")
    out = {k: [] for k in ("person_id", "household", "age", "role", "home_loc", "home_sub", "work_loc")}
    household_sub: Dict[int, Dict[int, int]] = {}
    for line, r in enumerate(rows[1:], start=2):
        if not r:
            continue
        pid = int(float(r[col["person_id"]]))
        hh = int(float(r[col["household_id"]]))
        age = int(float(r[col["age"]]))
        home = int(float(r[col["home_id"]]))
        w = r[col["workplace_id"]].strip()
        work = -1 if w == "" else int(float(w))
        role = int(float(r[col["social_role"]]))
        if age < 0 or age > 120:
            print(f"Person {pid} has age {age} on row {line}\n{r}", file=sys.stderr)
        if home not in index_of:
            print(f"homeId {home} not found in the location map on line {line}\n{r}", file=sys.stderr)
            continue
        subs = household_sub.setdefault(home, {})
        if hh in subs:
            home_sub = subs[hh]
        else:
            home_sub = len(subs)
            if home_sub + 1 > int(loc_nsub[index_of[home]]):
                print(f"Person {pid}. The homeId {home} with householdId {hh} has more sublocations ({home_sub + 1}) "
                      f"than defined. Record on row {line}\n{r}", file=sys.stderr)
            subs[hh] = home_sub
        if 2 <= role <= 6:
            if work == -1 or work not in index_of:
                print(f"No school location [{work}] for Student on line {line}\n{r}", file=sys.stderr)
                continue
            if names[int(loc_type[index_of[work]])] != STUDENT_SCHOOL[role]:
                print(f"workSchoolId {work} not a {STUDENT_SCHOOL[role]} in the location map on line {line}\n{r}",
                      file=sys.stderr)
                continue
        elif role in ROLE_LABEL:
            if work == -1 or work not in index_of:
                print(f"No work location [{work}] for {ROLE_LABEL[role]} on line {line}\n{r}", file=sys.stderr)
                continue
        elif role not in (1, 8, 9):
            print(f"social role {role} not recognized on line {line}\n{r}", file=sys.stderr)
            continue
        out["person_id"].append(pid); out["household"].append(hh); out["age"].append(min(max(age, -128), 127))
        out["role"].append(role); out["home_loc"].append(index_of[home]); out["home_sub"].append(home_sub)
        out["work_loc"].append(index_of[work] if work in index_of else -1)
    n = len(out["person_id"])
    x = scenario.philox4x32_10(np.arange(n), 0, 4, scenario.TAG_SCHEDULE, seed & 0xFFFFFFFF, seed >> 32)[0]
    female = (x.astype(np.float64) * 2.0 ** -32 < 0.5).astype(np.uint8)
    return dict(person_id=np.array(out["person_id"], np.int64), household=np.array(out["household"], np.int64),
                age=np.array(out["age"], np.int8), role=np.array(out["role"], np.uint8),
                home_loc=np.array(out["home_loc"], np.int32), home_sub=np.array(out["home_sub"], np.int16),
                work_loc=np.array(out["work_loc"], np.int32), female=female)


def read_infection_rates(path: str) -> Dict[str, np.ndarray]:
    """``epidemiology/infection_rates.csv`` (generic.ProbRatioFilePath; ConstructHerosModel.java:274-295): the locations the
    reference builds as ``LocationProbBased`` -- columns location_id, infection_rate_factor, infection_rate (-1 = not
    given), the rest is descriptive.  Returns the three columns heros_set_rate_locations takes."""
    loc, factor, rate = [], [], []
    with _open(path) as f:
        rows = csv.reader(f)
        next(rows, None)  # header (the reference skips it unread)
        for row in rows:
            if not row:
                continue
            loc.append(int(row[0])); factor.append(float(row[1])); rate.append(float(row[2]))
    return {"location": np.array(loc, np.int32), "rate_factor": np.array(factor, np.float64), "rate": np.array(rate, np.float64)}


def load_city(locations_path: str, people_path: str, type_names: Optional[Sequence[str]] = None, seed: int = 111,
              n_random_candidates: int = 8) -> "scenario.City":
    """A ``scenario.City`` over the real tables of ``data/<city>/locations/locations.csv.gz`` and
    ``data/<city>/people/people.csv.gz``: ``city.install(engine)`` and ``scenario.compile_schedule(city)`` then work as
    for the synthetic city; ``city.location_id`` / ``city.person_id`` map engine indices back to the ids of the files."""
    locations = read_locations(locations_path, type_names)
    people = read_people(people_path, locations, type_names, seed)
    city = scenario.City(len(people["age"]), seed=seed, n_random_candidates=n_random_candidates,
                         location_table=locations, person_table=people)
    city.location_id, city.person_id = locations["location_id"], people["person_id"]
    return city


# role id 1..15 -> role name of the week-pattern keys (java :565-743, model/HerosModel.java:313-330)
ROLES = ["infant", "kindergarten student", "primary school student", "secondary school student", "college student",
         "university student", "worker", "pensioner", "unemployed job-seeker", "weekend worker", "essential worker",
         "worker satellite to city", "worker city to satellite", "worker satellite to satellite",
         "worker country to city"]
PHASES = ["Susceptible", "Exposed", "Infected-Asymptomatic", "Infected-Symptomatic", "Hospitalized", "ICU", "Dead",
          "Recovered"]
ACTIVITY_KIND = {"StochasticDurationActivity": 0, "FixedDurationActivity": 0, "TravelActivityDistanceBased": 1,
                 "TravelActivity": 1, "UntilFixedTimeActivity": 2}
LOCATOR_CAP = 0x10  # flag of the capacity-constrained forms (scenario.capacity_diversion, schedule.cu)
LOCATOR = {"HomeLocator": 0, "WorkLocator": 1, "SchoolLocator": 1, "CurrentLocator": 2, "NearestLocator": 3,
           "NearestLocatorChoice": 3, "NearestLocatorCap": 3 | LOCATOR_CAP, "NearestLocatorChoiceCap": 3 | LOCATOR_CAP,
           "RandomLocator": 4, "RandomLocatorChoice": 4, "RandomLocatorCap": 4 | LOCATOR_CAP,
           "RandomLocatorChoiceCap": 4 | LOCATOR_CAP, "DistanceBasedTravelLocator": 5, "WalkLocator": 5, "BikeLocator": 5,
           "CarLocator": 5}


def read_activity_schedules(path: str, sheet: str = "activityschedules", policy: str = "0") -> dict:
    """The activity-schedule workbook of a scenario (``data/<city>/activities/activityschedules*.xlsx``) as the
    structure ``scenario.load_patterns`` returns and ``scenario.compile_schedule`` / ``ScheduleGenerator`` interpret.

    Row format: columns A-Q of java :763-904 -- policy, disease phase, day, role, activity name, activity class,
    until-time, distribution, mode, min, max, (unused), locator, (unused), travel locator, max distance, location-type
    choice ("LocationType.X:0.6; LocationType.Y:0.4").  Only the rows of ``policy`` are read (policies 1 and 2 of the
    shipped workbook are never referenced, java :400).  A day pattern ends with its first ``UntilFixedTimeActivity 24``:
    10 of the 15 roles store the Susceptible block twice back to back, the second copy is dropped.
    Capacity-constrained locators (``*Cap``) keep the flag ``LOCATOR_CAP``: the schedule generators divert a share of the
    visits to an over-subscribed place to the next candidate (scenario.capacity_diversion)."""
    import json
    rows = read_xlsx(path)[sheet][1:]
    choices, choice_index = [], {}

    def number(x, default=0.0):
        return default if x in (None, "") else float(x)

    def choice_id(text):
        if not text:
            return -1
        items = []
        for part in text.split(";"):
            part = part.strip()
            if not part:
                continue
            if ":" in part:
                name, p = part.split(":")
                items.append([name.strip().replace("LocationType.", ""), float(p)])
            else:
                items.append([part.replace("LocationType.", ""), 1.0])
        key = json.dumps(items)
        if key not in choice_index:
            choice_index[key] = len(choices)
            choices.append(items)
        return choice_index[key]

    patterns, closed = {}, set()
    for r in rows:
        r = list(r) + [None] * (17 - len(r))
        if r[0] != policy:
            continue
        key = f"{r[1]}|{r[3]}|{r[2]}"
        if key in closed:
            continue
        if r[5] not in ACTIVITY_KIND:
            raise ValueError(f"unknown activity class {r[5]!r} in {path}")
        kind = ACTIVITY_KIND[r[5]]
        dist = {"triangular": 2, "uniform": 1}.get((r[7] or "").strip(), 0)
        mode, mn, mx = number(r[8]), number(r[9]), number(r[10])
        locator_name = r[14] if kind == 1 else r[12]
        if locator_name not in LOCATOR:
            raise ValueError(f"unknown locator {locator_name!r} in {path}")
        locator, max_distance, cid = LOCATOR[locator_name], number(r[15]), choice_id(r[16])
        if kind == 1 and dist == 0:
            dist = 1  # travel time ~ U(min, max) hours (the reference derives it from the distance)
        if dist == 0 and kind == 0:
            mn = mx = mode  # FixedDurationActivity
        act = {"name": (r[4] or "").strip(), "kind": kind, "locator": locator, "dist": dist, "min": mn, "mode": mode,
               "max": mx, "until": number(r[6]), "choice": cid, "max_distance": max_distance}
        patterns.setdefault(key, []).append(act)
        if kind == 2 and act["until"] >= 24:
            closed.add(key)
    return {"source": path, "roles": ROLES, "phases": PHASES, "days": ["Monday", "Friday", "Saturday", "Sunday"],
            "choices": choices, "patterns": patterns}
