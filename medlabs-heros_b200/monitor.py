"""Mirror of the medlabs DiseaseMonitor the reference registers with a 0.5 h interval
(factory/ConstructHerosModel.java:145: ``model.setDiseaseMonitor(new DiseaseMonitor(model, progression, 0.5))``).

The reference samples the live person counter of every DiseasePhase (addPerson / removePerson,
disease/Covid19Progression.java:184-186,211) every half hour into the compartment time series whose columns are the
phases in registration order (Covid19Progression.java:130-137) -- the layout of the golden spreadsheets
(``Time(h), Susceptible, Exposed, Infected-Asymptomatic, Infected-Symptomatic, Hospitalized, ICU, Dead, Recovered``).
The GPU engine logs every phase entered (heros_get_transitions); the series is rebuilt from that log exactly, for any
sampling interval, without sampling events on the host.  A transition at exactly a sampling instant is counted in that
sample (phase-change events fire before the monitor's event of the same time).
"""
from __future__ import annotations

import csv
import gzip
import os
from typing import Dict, Tuple

import numpy as np

from ._lib import PHASE_NAMES

HEADER = ("Time(h)",) + PHASE_NAMES


def compartment_series(transitions: Dict[str, np.ndarray], n_persons: int, t_end: float, interval: float = 0.5,
                       person_base: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """(times[k], counts[k, 8]) for t = 0, interval, ... <= t_end from a transition log (person, phase, time); everybody
    starts Susceptible (ConstructHerosModel.java:744-747)."""
    person = np.asarray(transitions["person"], np.int64) - person_base
    phase = np.asarray(transitions["phase"], np.int64)
    time = np.asarray(transitions["time"], np.float64)
    order = np.lexsort((time, person))          # per person in time order: the phase left is the one entered before
    p, ph, t = person[order], phase[order], time[order]
    first = np.r_[True, p[1:] != p[:-1]] if len(p) else np.zeros(0, bool)
    prev = np.where(first, 0, np.r_[0, ph[:-1]]) if len(p) else ph
    by_time = np.argsort(t, kind="stable")
    t, ph, prev = t[by_time], ph[by_time], prev[by_time]
    delta = np.zeros((len(t) + 1, 8), np.int64)
    delta[0, 0] = n_persons
    rows = np.arange(1, len(t) + 1)
    np.add.at(delta, (rows, ph), 1)
    np.add.at(delta, (rows, prev), -1)
    running = np.cumsum(delta, axis=0)
    times = np.arange(0.0, t_end + 1e-9, interval)
    return times, running[np.searchsorted(t, times, side="right")]


class DiseaseMonitor:
    """DiseaseMonitor(model, diseaseProgression, intervalHours): the compartment series of a run of ``model.engine``."""

    def __init__(self, model, diseaseProgression=None, interval: float = 0.5):
        self.model, self.interval = model, float(interval)

    def series(self, t_end: float = None) -> Tuple[np.ndarray, np.ndarray]:
        eng = self.model.engine
        t_end = eng.time if t_end is None else t_end
        return compartment_series(eng.transitions(), eng.n_persons, t_end, self.interval)

    def write_csv(self, path: str, t_end: float = None) -> None:
        times, counts = self.series(t_end)
        with open(path, "w", newline="") as f:
            w = csv.writer(f, quoting=csv.QUOTE_NONNUMERIC)
            w.writerow(HEADER)
            for t, row in zip(times, counts):
                w.writerow([float(t)] + [int(x) for x in row])


class LocationDump:
    """The two hourly occupancy files of the reference (model/HerosModel.java:236-299, ``generic.WriteOutput``):
    ``locationNrPersons.csv.gz`` (``"Time(h)","LocationNr","NrPersons"``) and ``sublocationNrPersons.csv.gz``
    (``"Time(h)","LocationNr","SubLocationNr","NrPersons"``), one block of rows per simulated hour.

    As in the reference, the sub-location file lists the indices 1 .. nSub-1 only (the loop of HerosModel.java:282
    starts at 1), so locations with a single sub-location do not appear in it.  Rows are ordered by location type and
    then by location id (the reference follows the iteration order of a Trove hash map per type, which is not
    reproducible).  Usage: after the stays of a day are in the engine (``submit_visits`` / ``advance_schedule``) and
    before ``step``, call ``dump_hours(t0, t1)`` -- it samples the hours ``t0 <= t < t1`` on the device
    (``heros_sample_occupancy``) and appends them."""

    LOCATION_HEADER = '"Time(h)","LocationNr","NrPersons"\n'
    SUBLOCATION_HEADER = '"Time(h)","LocationNr","SubLocationNr","NrPersons"\n'

    def __init__(self, engine, location_type, output_path: str, interval: float = 1.0):
        self.engine, self.interval = engine, float(interval)
        ty = np.asarray(location_type, np.int64)
        self.order = np.lexsort((np.arange(len(ty)), ty))          # by type, then by id
        n_sub = engine.location_n_sub.astype(np.int64)
        first = engine.location_first_sub
        # (location, sub-location index, position in the per-sub-location counts) of every row of the second file
        loc = np.repeat(self.order, np.maximum(n_sub[self.order] - 1, 0))
        start = np.cumsum(np.r_[0, np.maximum(n_sub[self.order] - 1, 0)])[:-1]
        idx = np.arange(len(loc)) - np.repeat(start, np.maximum(n_sub[self.order] - 1, 0)) + 1
        self.sub_loc, self.sub_index, self.sub_pos = loc, idx, first[loc] + idx
        os.makedirs(output_path, exist_ok=True)
        self.loc_file = gzip.open(os.path.join(output_path, "locationNrPersons.csv.gz"), "wt", encoding="utf-8")
        self.sub_file = gzip.open(os.path.join(output_path, "sublocationNrPersons.csv.gz"), "wt", encoding="utf-8")
        self.loc_file.write(self.LOCATION_HEADER)
        self.sub_file.write(self.SUBLOCATION_HEADER)

    def dump_hours(self, t0: float, t1: float) -> int:
        times = np.arange(t0, t1 - 1e-9, self.interval)
        per_loc, per_sub = self.engine.sample_occupancy(times)
        for j, t in enumerate(times):
            ts = repr(float(t))                                    # Java's Double.toString for these values
            self.loc_file.write("".join(f"{ts},{l},{n}\n" for l, n in zip(self.order.tolist(), per_loc[j, self.order].tolist())))
            self.sub_file.write("".join(f"{ts},{l},{i},{n}\n" for l, i, n in zip(self.sub_loc.tolist(), self.sub_index.tolist(),
                                                                                per_sub[j, self.sub_pos].tolist())))
        return len(times)

    def close(self) -> None:
        self.loc_file.close()
        self.sub_file.close()
