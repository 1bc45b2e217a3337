"""Host-side mirror of the reference's disease plug-in API, backed by the GPU engine.

Same class names, method names, argument meaning and error behaviour as
  eu/heros/disease/Covid19TransmissionArea.java, Covid19TransmissionDistance.java, Covid19Progression.java
(the three classes the medlabs engine calls through its abstract DiseaseTransmission / DiseaseProgression API,
SURVEY.md 8b), so parity tests read like tests of the reference.  The Java host would bind the same C entry points
through Panama FFM; this Python layer exists because the container has no JVM.
"""
from __future__ import annotations

import sys
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib, properties
from .engine import HerosEngine


@dataclass
class InfectionRecord:
    """medlabs InfectionRecord(exposurePhase, location) as filled by infectPeople (TArea:135-193)."""
    exposurePhase: str
    location: int
    calculated: bool = False
    infectiousPersons: List[int] = field(default_factory=list)
    infectedPersons: List[int] = field(default_factory=list)
    pInfection: float = float("nan")

    def isCalculated(self) -> bool:
        return self.calculated


class HerosModel:
    """The slice of MedlabsModelInterface the disease plug-ins use: the parameter map (getParameterValue*), the
    simulator clock, and -- instead of person / location object maps -- the GPU engine holding that state."""

    def __init__(self, parameters: Dict[str, str], seed: Optional[int] = None, device: int = 0,
                 trigger: int = _lib.TRIGGER_EXIT_ONLY, window_hours: float = 24.0):
        self.parameters = dict(parameters)
        model_name = self.parameters.get("generic.diseasePropertiesModel", "area").strip().lower()
        if model_name not in ("area", "distance"):
            raise ValueError(f"generic.diseasePropertiesModel must be area or distance, not {model_name!r}")
        self.model_kind = _lib.MODEL_AREA if model_name == "area" else _lib.MODEL_DISTANCE
        self.seed = int(self.parameters.get("generic.Seed", 111)) if seed is None else int(seed)
        self.engine = HerosEngine(self.model_kind, self.seed, trigger=trigger, device=device,
                                  window_hours=window_hours)
        self.diseaseProgression: Optional["Covid19Progression"] = None
        self.diseaseTransmission = None

    # --- MedlabsModelInterface parameter access (TArea:63-68, ConstructHerosModel.java:1111-1113) ---
    def getParameterValue(self, key: str) -> str:
        return self.parameters[key]

    def getParameterValueDouble(self, key: str) -> float:
        return float(self.parameters[key])

    def getParameterValueInt(self, key: str) -> int:
        return int(float(self.parameters[key]))

    def getSimulatorTime(self) -> float:
        return self.engine.time

    def setDiseaseProgression(self, p):  # ConstructHerosModel.java:143
        self.diseaseProgression = p

    def setDiseaseTransmission(self, t):  # ConstructHerosModel.java:144
        self.diseaseTransmission = t

    def getDiseaseTransmission(self):
        return self.diseaseTransmission

    def getDiseaseProgression(self):
        return self.diseaseProgression

    def close(self):
        self.engine.close()


class DiseaseTransmission:
    """medlabs abstract DiseaseTransmission(model, name)."""

    def __init__(self, model: HerosModel, name: str):
        self.model = model
        self.name = name

    def infectPeople(self, location: int, personsInSublocation: Sequence[int], duration: float,
                     allPersonIds: Optional[Sequence[int]] = None, now: Optional[float] = None) -> InfectionRecord:
        """infectPeople(Location, TIntSet personsInSublocation, double duration): one evaluation on the GPU.

        ``allPersonIds`` stands for ``location.getAllPersonIds()`` (only read by the total-location branch) and
        ``now`` for ``model.getSimulator().getSimulatorTime()``.
        """
        t = self.model.getSimulatorTime() if now is None else now
        r = self.model.engine.infect_people(location, personsInSublocation, t, duration, all_persons=allPersonIds)
        return InfectionRecord(Covid19Progression.exposed, location, r["calculated"], r["infectious"].tolist(),
                               r["infected"].tolist(), r["p"])

    def setParameter(self, parameterName: str, value: float, at_time: Optional[float] = None) -> None:
        t = self.model.getSimulatorTime() if at_time is None else at_time
        if self.model.engine.set_parameter(parameterName, value, t) != _lib.OK:
            print("Unrecognized variable name " + parameterName, file=sys.stderr)  # TArea:254 / TDist:263


class Covid19TransmissionArea(DiseaseTransmission):
    """Covid19TransmissionArea(model): reads covidT_area.* (java :63-68)."""

    def __init__(self, model: HerosModel):
        super().__init__(model, "Covid19")
        self.params = properties.area_params(model.parameters)
        model.engine.set_transmission_area(self.params)


class Covid19TransmissionDistance(DiseaseTransmission):
    """Covid19TransmissionDistance(model): reads covidT_dist.* (java :59-68); psi and mu are run-time tunable."""

    def __init__(self, model: HerosModel):
        super().__init__(model, "Covid19")
        self.params = properties.dist_params(model.parameters)
        model.engine.set_transmission_distance(self.params)


class Covid19Progression:
    """Covid19Progression(model): the 8 phases (java :130-137) and covidP.* parameters (java :143-173)."""

    susceptible = "Susceptible"
    exposed = "Exposed"
    infected_asymptomatic = "Infected-Asymptomatic"
    infected_symptomatic = "Infected-Symptomatic"
    hospitalized = "Hospitalized"
    icu = "ICU"
    dead = "Dead"
    recovered = "Recovered"

    def __init__(self, model: HerosModel):
        self.model = model
        self.name = "Covid19"
        self.params = properties.prog_params(model.parameters)  # raises on unparsable values, like MedlabsException
        model.engine.set_progression(self.params)

    def getDiseasePhases(self) -> List[str]:
        return list(_lib.PHASE_NAMES)

    def expose(self, exposedPerson, exposurePhase: str = "Exposed", at_time: Optional[float] = None) -> None:
        """expose(Person, DiseasePhase) for one person id or a sequence of ids (java :182-201)."""
        if exposurePhase != self.exposed:
            raise ValueError("expose() only takes the Exposed phase")
        t = self.model.getSimulatorTime() if at_time is None else at_time
        ids = np.atleast_1d(np.asarray(exposedPerson, np.int32))
        self.model.engine.expose(ids, np.full(len(ids), t, np.float64))


def construct_disease_model(model: HerosModel):
    """ConstructHerosModel.java:136-144: progression first, then the transmission model selected by
    generic.diseasePropertiesModel, and registration on the model."""
    progression = Covid19Progression(model)
    if model.model_kind == _lib.MODEL_AREA:
        transmission = Covid19TransmissionArea(model)
    else:
        transmission = Covid19TransmissionDistance(model)
    model.setDiseaseProgression(progression)
    model.setDiseaseTransmission(transmission)
    return progression, transmission
