"""Mirror of eu/heros/policy/DiseasePolicy.java: a timed change of a transmission parameter."""
from __future__ import annotations

import sys

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
