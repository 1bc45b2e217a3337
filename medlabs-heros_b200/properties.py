"""Scenario / disease parameter handling on the host side.

Mirrors how the reference reads its inputs for the hot path:
  * ``.properties`` files with dotted keys ``covidT_area.* covidT_dist.* covidP.* generic.* policies.*``
    (src/main/resources/*.properties; loaded by HerosApplication.java:146-150),
  * ``ConditionalProbability`` strings: a number, ``age{0-19: 0.8, 20-55: 0.5}`` or ``gender{M:0.45, F:0.5}``
    (alpha-area.properties:42-47),
  * duration distributions ``Triangular(a,b,c)``, ``Uniform(a,b)``, ``Constant(x)`` parsed by medlabs'
    DistributionParser (Covid19Progression.java:144-173),
  * ``locationtypes.csv`` (ConstructHerosModel.java:230-269) and disease policy CSV files (:1176-1204).
"""
from __future__ import annotations

import csv
import re
from typing import Dict, List, Tuple

import numpy as np

from . import _lib

# covidP.* keys in the order of heros_prog_params (include/heros_gpu.h)
FRACTION_KEYS = ("covidP.FractionAsymptomatic", "covidP.FractionSymptomaticToHospitalized",
                 "covidP.FractionHospitalizedToICU", "covidP.FractionHospitalizedToDead", "covidP.FractionICUToDead")
PERIOD_KEYS = ("covidP.IncubationPeriodAsymptomatic", "covidP.IncubationPeriodSymptomatic",
               "covidP.PeriodAsymptomaticToRecovered", "covidP.PeriodSymptomaticToHospitalized",
               "covidP.PeriodSymptomaticToRecovered", "covidP.PeriodHospitalizedToICU",
               "covidP.PeriodHospitalizedToDead", "covidP.PeriodHospitalizedToRecovered", "covidP.PeriodICUToDead",
               "covidP.PeriodICUToRecovered")
AREA_KEYS = ("contagiousness", "beta", "t_e_min", "t_e_mode", "t_e_max", "calculation_threshold")
DIST_KEYS = ("L", "I", "C", "v_max", "v_0", "r", "psi", "alpha", "mu", "calculation_threshold")


def load_properties(path: str) -> Dict[str, str]:
    """Read a Java ``.properties`` file (``key = value``; ``#`` / ``!`` comments)."""
    out: Dict[str, str] = {}
    with open(path, "r", encoding="utf-8", errors="replace") as f:
        for raw in f:
            line = raw.strip()
            if not line or line[0] in "#!":
                continue
            m = re.match(r"([^=:\s]+)\s*[=:]?\s*(.*)$", line)
            if m:
                out[m.group(1)] = m.group(2).strip()
    return out


def parse_conditional_probability(text: str) -> np.ndarray:
    """``ConditionalProbability.probability(person)`` tabulated as float64 [2 (female)][128 (age)].

    Ages outside every band get probability 0; with overlapping bands the first match wins (the tie rule of medlabs
    is not visible in the reference; the shipped files never overlap).
    """
    s = text.strip()
    table = np.zeros((2, 128), np.float64)
    m = re.match(r"(?i)^(age|gender)\s*\{(.*)\}$", s)
    if not m:
        table[:, :] = float(s)
        return table
    kind, body = m.group(1).lower(), m.group(2)
    entries = [e for e in body.split(",") if e.strip()]
    if kind == "age":
        done = np.zeros(128, bool)
        for e in entries:
            band, val = e.split(":")
            lo, hi = (int(x) for x in band.strip().split("-"))
            v = float(val)
            for a in range(max(lo, 0), min(hi, 127) + 1):
                if not done[a]:
                    table[:, a] = v
                    done[a] = True
    else:
        for e in entries:
            g, val = e.split(":")
            g = g.strip().upper()
            if g.startswith("M"):
                table[0, :] = float(val)
            elif g.startswith("F"):
                table[1, :] = float(val)
            else:
                raise ValueError(f"unknown gender key {g!r} in {text!r}")
    return table


def parse_distribution(text: str) -> Tuple[int, float, float, float]:
    """``DistributionParser.parseDistContinuous``: returns (kind, a, b, c), parameters in days."""
    m = re.match(r"^\s*([A-Za-z]+)\s*\(([^)]*)\)\s*$", text)
    if not m:
        raise ValueError(f"cannot parse distribution {text!r}")
    name = m.group(1).lower()
    args = [float(x) for x in m.group(2).split(",") if x.strip()]
    if name in ("triangular", "tria", "triang") and len(args) == 3:
        return (_lib.DIST_TRIANGULAR, args[0], args[1], args[2])
    if name in ("uniform", "unif") and len(args) == 2:
        return (_lib.DIST_UNIFORM, args[0], args[1], 0.0)
    if name in ("constant", "const") and len(args) == 1:
        return (_lib.DIST_CONSTANT, args[0], 0.0, 0.0)
    raise NotImplementedError(f"distribution {text!r} is not supported by the GPU engine")


def area_params(props: Dict[str, str]) -> _lib.AreaParams:
    """covidT_area.* as read by Covid19TransmissionArea's constructor (java :63-68)."""
    return _lib.AreaParams(*[float(props["covidT_area." + k]) for k in AREA_KEYS])


def dist_params(props: Dict[str, str]) -> _lib.DistParams:
    """covidT_dist.* as read by Covid19TransmissionDistance's constructor (java :59-68)."""
    return _lib.DistParams(*[float(props["covidT_dist." + k]) for k in DIST_KEYS])


def prog_tables(props: Dict[str, str]):
    """covidP.* -> (fractions float64 [5][2][128], periods list of (kind, a, b, c))."""
    fractions = np.stack([parse_conditional_probability(props[k]) for k in FRACTION_KEYS])
    periods = [parse_distribution(props[k]) for k in PERIOD_KEYS]
    return fractions, periods


def prog_params(props: Dict[str, str]) -> _lib.ProgParams:
    """covidP.* as read by Covid19Progression's constructor (java :143-173)."""
    import ctypes as C
    fractions, periods = prog_tables(props)
    pp = _lib.ProgParams()
    f = np.ascontiguousarray(fractions)
    C.memmove(pp.fraction, f.ctypes.data, f.nbytes)
    for i, (kind, a, b, c) in enumerate(periods):
        pp.period[i] = _lib.Duration(kind, 0, a, b, c)
    return pp


def _bool(s: str) -> bool:
    return s.strip().upper() in ("TRUE", "1", "YES", "T")


def load_location_types(path: str) -> Dict[str, object]:
    """``locationtypes.csv`` -- columns 0,1,2,3,5 and 7..9 are the ones the loader reads
    (ConstructHerosModel.java:251-265); type id = row order."""
    names: List[str] = []
    infect_sub, corr, cap, cap_m2, size = [], [], [], [], []
    with open(path, newline="", encoding="utf-8") as f:
        rows = list(csv.reader(f))
    for row in rows[1:]:
        if not row or not row[0].strip():
            continue
        names.append(row[0].strip())
        infect_sub.append(_bool(row[3]))
        corr.append(float(row[5]))
        cap.append(_bool(row[7]) if len(row) > 7 else False)
        cap_m2.append(float(row[8]) if len(row) > 8 else 0.0)
        size.append(float(row[9]) if len(row) > 9 else 1.0)
    return dict(name=names, infect_sub=np.array(infect_sub, np.uint8), correction_factor_area=np.array(corr, np.float64),
                cap_constrained=np.array(cap, np.uint8), cap_persons_per_m2=np.array(cap_m2, np.float64),
                size_factor=np.array(size, np.float64))


def load_disease_policies(path: str) -> List[Tuple[float, str, float]]:
    """Disease policy CSV ``Time(d),Parameter,Value`` -> [(time_hours, parameter, value)]
    (ConstructHerosModel.java:1176-1204: time = 24 * days)."""
    out = []
    with open(path, newline="", encoding="utf-8") as f:
        rows = list(csv.reader(f))
    for row in rows[1:]:
        if len(row) >= 3 and row[0].strip():
            out.append((24.0 * float(row[0]), row[1].strip(), float(row[2])))
    return out
