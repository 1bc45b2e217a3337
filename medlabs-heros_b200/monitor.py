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
