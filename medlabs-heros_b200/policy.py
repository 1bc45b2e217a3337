"""Mirror of eu/heros/policy/DiseasePolicy.java (a timed change of a transmission parameter) and
eu/heros/policy/LocationPolicy.java (a timed closure / reopening of a location type), plus the two policy-file readers
of factory/ConstructHerosModel.java:1145-1206."""
from __future__ import annotations

import csv
import sys
from typing import List, Optional, Sequence

from .disease import HerosModel


class DiseasePolicy:
    """DiseasePolicy(model, time, parameterName, value) (java :18-27).

    The reference schedules ``changeParameter`` on the DSOL event list for absolute time ``time``; the GPU engine
    keeps a time-stamped parameter schedule instead, so the change is registered immediately and takes effect for every
    evaluation whose simulator time is >= ``time`` (java :29-43).
    """

    def __init__(self, model: HerosModel, time: float, parameterName: str, value: float):
        self.model = model
        if parameterName not in ("covidT_dist.psi", "covidT_dist.mu"):
            print("Unrecognized variable name " + parameterName, file=sys.stderr)  # java :21-25
            return
        self.changeParameter(parameterName, value, time)

    def changeParameter(self, parameterName: str, value: float, time: float) -> None:
        if parameterName == "covidT_dist.psi":
            self.model.getDiseaseTransmission().setParameter("psi", value, at_time=time)
        elif parameterName == "covidT_dist.mu":
            self.model.getDiseaseTransmission().setParameter("mu", value, at_time=time)
        else:
            print("Unrecognized variable name " + parameterName, file=sys.stderr)


class LocationPolicy:
    """LocationPolicy(model, time, locationTypeName, fractionOpen, fractionActivities, alternativeLocationName,
    reportAsLocationName) (java :19-37).

    The reference schedules ``openCloseLocationType`` on the DSOL event list for absolute time ``time`` (java :35); here
    the policy is queued on the model and ``apply_location_policies(model, t)`` -- called by the day loop before
    ``advance_schedule`` -- hands every policy with ``time <= t`` to ``heros_set_closure_policy``
    (= ``LocationType.setClosurePolicy``, java :45).  Unknown type names are reported on stderr and ignored, as in
    java :24-34."""

    def __init__(self, model: HerosModel, time: float, locationTypeName: str, fractionOpen: float,
                 fractionActivities: float, alternativeLocationName: str, reportAsLocationName: str,
                 type_names: Optional[Sequence[str]] = None):
        from .scenario import TYPE_NAMES
        names = list(type_names if type_names is not None else TYPE_NAMES)
        self.model = model
        self.scheduled = False
        if locationTypeName not in names:
            print("LocationPolicy: Did not recognize locationType " + locationTypeName, file=sys.stderr)
            return
        if alternativeLocationName not in names:
            print("LocationPolicy: Did not recognize alternativeLocationType " + alternativeLocationName, file=sys.stderr)
            return
        self.time = float(time)
        self.locationType = names.index(locationTypeName)
        self.fractionOpen, self.fractionActivities = float(fractionOpen), float(fractionActivities)
        self.alternativeLocationType = names.index(alternativeLocationName)
        self.reportAsLocationName = reportAsLocationName
        if not hasattr(model, "locationPolicies"):
            model.locationPolicies = []
        model.locationPolicies.append(self)
        self.scheduled = True

    def openCloseLocationType(self, target=None) -> None:
        """java :39-46.  ``target``: what receives setClosurePolicy -- the engine of the model (device-side schedule) by
        default, or a ``scenario.ScheduleGenerator`` (host-side schedule)."""
        target = self.model.engine if target is None else target
        target.set_closure_policy(self.locationType, self.fractionOpen, self.fractionActivities,
                                  self.alternativeLocationType)


def apply_location_policies(model: HerosModel, time: float, target=None) -> int:
    """Fires, in file order, the queued LocationPolicy events with event time <= ``time`` (hours); returns how many."""
    due = [p for p in getattr(model, "locationPolicies", []) if p.time <= time]
    for p in due:
        p.openCloseLocationType(target)
        model.locationPolicies.remove(p)
    return len(due)


def schedule_location_policy_rows(model: HerosModel, rows, type_names: Optional[Sequence[str]] = None) -> List[LocationPolicy]:
    """The loop of ConstructHerosModel.scheduleLocationPolicies (java :1157-1171) over rows
    ``Time(d),LocationType,FractionOpen,FractionActivities,AlternativeLocation,ReportAsLocation``; days -> hours."""
    out = []
    for row in rows:
        if len(row) < 6:
            continue
        out.append(LocationPolicy(model, 24.0 * float(row[0]), row[1], float(row[2]), float(row[3]), row[4], row[5],
                                  type_names))
    return out


def schedule_location_policies(model: HerosModel, path: str, type_names: Optional[Sequence[str]] = None) -> List[LocationPolicy]:
    """ConstructHerosModel.scheduleLocationPolicies (java :1145-1174): the CSV file named by
    ``policies.LocationPolicyFile`` (header line skipped)."""
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    return schedule_location_policy_rows(model, rows[1:], type_names)


def schedule_disease_policies(model: HerosModel, path: str) -> List[DiseasePolicy]:
    """ConstructHerosModel.scheduleDiseasePolicies (java :1179-1205): rows ``Time(d),Parameter,Value``."""
    out = []
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    for row in rows[1:]:
        if len(row) < 3:
            continue
        out.append(DiseasePolicy(model, 24.0 * float(row[0]), row[1], float(row[2])))
    return out
