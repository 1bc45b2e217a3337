"""ctypes binding of the CPU oracle (oracle/heros_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libheros_oracle.so")

NPHASE = 8
S, E, IA, IS, H, ICU, D, R = range(8)
MODEL_AREA, MODEL_DIST = 0, 1
TRIGGER_EXIT_ONLY, TRIGGER_ENTER_AND_EXIT = 0, 1
DIST_CONSTANT, DIST_UNIFORM, DIST_TRIANGULAR = 0, 1, 2


class AreaProps(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("contagiousness", "beta", "t_e_min_days", "t_e_mode_days", "t_e_max_days", "calculation_threshold_s")]


class DistProps(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("L_days", "I_days", "C_days", "v_max", "v_0", "r", "psi", "alpha", "mu", "calculation_threshold_s")]


class Duration(C.Structure):
    _fields_ = [("kind", C.c_int32), ("a", C.c_double), ("b", C.c_double), ("c", C.c_double)]


class ProgProps(C.Structure):
    _fields_ = [("fraction", C.c_double * 128 * 2 * 5), ("period", Duration * 10)]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile).  Returns the library path."""
    src = os.path.join(_HERE, "heros_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        P = C.c_void_p
        L.orc_philox4x32_10.argtypes = [P, P, P]
        L.orc_u01.restype = C.c_double
        L.orc_u01.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32]
        L.orc_exp.restype = C.c_double
        L.orc_exp.argtypes = [C.c_double]
        L.orc_java_max.restype = C.c_double
        L.orc_java_max.argtypes = [C.c_double, C.c_double]
        L.orc_sample_duration_hours.restype = C.c_double
        L.orc_sample_duration_hours.argtypes = [C.POINTER(Duration), C.c_double]
        L.orc_infect_people_area.restype = C.c_int
        L.orc_infect_people_area.argtypes = [C.POINTER(AreaProps), C.c_uint64, C.c_int, C.c_double, C.c_float, C.c_int,
                                             C.c_int, P, P, P, C.c_int, P, P, P, C.c_double, C.c_double,
                                             P, P, P, P, P]
        L.orc_infect_people_dist.restype = C.c_int
        L.orc_infect_people_dist.argtypes = [C.POINTER(DistProps), C.c_uint64, C.c_int, C.c_float, C.c_int,
                                             C.c_int, P, P, P, C.c_int, P, P, P, C.c_double, C.c_double,
                                             P, P, P, P, P]
        L.orc_sim_create.restype = P
        L.orc_sim_create.argtypes = [C.c_int, C.c_int, C.c_uint64]
        L.orc_sim_destroy.argtypes = [P]
        L.orc_sim_set_area.argtypes = [P, C.POINTER(AreaProps)]
        L.orc_sim_set_dist.argtypes = [P, C.POINTER(DistProps)]
        L.orc_sim_set_prog.argtypes = [P, C.POINTER(ProgProps)]
        L.orc_sim_set_location_types.argtypes = [P, C.c_int, P, P]
        L.orc_sim_set_locations.argtypes = [P, C.c_int, P, P, P]
        L.orc_sim_set_persons.argtypes = [P, C.c_int, P, P]
        L.orc_sim_set_parameter.restype = C.c_int
        L.orc_sim_set_parameter.argtypes = [P, C.c_char_p, C.c_double, C.c_double]
        L.orc_sim_expose.argtypes = [P, C.c_int, P, P]
        L.orc_sim_set_roles.argtypes = [P, C.c_int, P]
        L.orc_sim_set_rate_locations.argtypes = [P, C.c_int, P, P, P]
        L.orc_sim_add_visits.argtypes = [P, C.c_int64, P, P, P, P, P]
        L.orc_sim_run.argtypes = [P, C.c_double]
        for name in ("orc_sim_n_infections", "orc_sim_n_transitions", "orc_sim_n_evaluations", "orc_sim_n_draws"):
            getattr(L, name).restype = C.c_int64
            getattr(L, name).argtypes = [P]
        L.orc_sim_get_infections.argtypes = [P, P, P, P, P, P]
        L.orc_sim_get_transitions.argtypes = [P, P, P, P]
        L.orc_sim_get_phase_counts.argtypes = [P, P]
        L.orc_sim_get_agent_state.argtypes = [P, P, P, P, P]
        L.orc_sim_last_calc.restype = C.c_double
        L.orc_sim_last_calc.argtypes = [P, C.c_int32, C.c_int16]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _arr(x, dtype):
    return np.ascontiguousarray(np.asarray(x, dtype=dtype))


def philox4x32_10(ctr, key):
    c = _arr(ctr, np.uint32)
    k = _arr(key, np.uint32)
    out = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_p(c), _p(k), _p(out))
    return out


def u01(seed, person, bits, tag):
    return lib().orc_u01(int(seed) & (2 ** 64 - 1), int(person) & 0xFFFFFFFF, int(bits), int(tag))


def exp(x):
    return lib().orc_exp(float(x))


def sample_duration_hours(kind, a, b, c, u):
    d = Duration(kind, a, b, c)
    return lib().orc_sample_duration_hours(C.byref(d), u)


def area_props(p: dict) -> AreaProps:
    return AreaProps(p["contagiousness"], p["beta"], p["t_e_min"], p["t_e_mode"], p["t_e_max"],
                     p["calculation_threshold"])


def dist_props(p: dict) -> DistProps:
    return DistProps(p["L"], p["I"], p["C"], p["v_max"], p["v_0"], p["r"], p["psi"], p["alpha"], p["mu"],
                     p["calculation_threshold"])


def prog_props(fractions: np.ndarray, periods) -> ProgProps:
    """fractions: float64 [5][2][128]; periods: 10 x (kind, a, b, c)."""
    pp = ProgProps()
    f = _arr(fractions, np.float64).reshape(5, 2, 128)
    C.memmove(pp.fraction, f.ctypes.data, f.nbytes)
    for i, (kind, a, b, c) in enumerate(periods):
        pp.period[i] = Duration(int(kind), float(a), float(b), float(c))
    return pp


def infect_people(model, props, seed, infect_sub, corr, total_area, n_sublocations, sub_ids, sub_phase, sub_exposure,
                  now, duration, all_ids=None, all_phase=None, all_exposure=None):
    """One call of infectPeople.  Returns dict(calculated, infectious, infected, p)."""
    sub_ids = _arr(sub_ids, np.int32)
    sub_phase = _arr(sub_phase, np.uint8)
    sub_exposure = _arr(sub_exposure, np.float32)
    if all_ids is None:
        all_ids, all_phase, all_exposure = sub_ids, sub_phase, sub_exposure
    all_ids = _arr(all_ids, np.int32)
    all_phase = _arr(all_phase, np.uint8)
    all_exposure = _arr(all_exposure, np.float32)
    cap = max(len(sub_ids), len(all_ids), 1)
    infectious = np.zeros(cap, np.int32)
    infected = np.zeros(cap, np.int32)
    n1 = C.c_int32(0)
    n2 = C.c_int32(0)
    p = C.c_double(0.0)
    if model == MODEL_AREA:
        calc = lib().orc_infect_people_area(C.byref(props), seed, int(infect_sub), float(corr), float(total_area),
                                            int(n_sublocations), len(sub_ids), _p(sub_ids), _p(sub_phase),
                                            _p(sub_exposure), len(all_ids), _p(all_ids), _p(all_phase),
                                            _p(all_exposure), float(now), float(duration), _p(infectious),
                                            C.addressof(n1), _p(infected), C.addressof(n2), C.addressof(p))
    else:
        calc = lib().orc_infect_people_dist(C.byref(props), seed, int(infect_sub), float(total_area),
                                            int(n_sublocations), len(sub_ids), _p(sub_ids), _p(sub_phase),
                                            _p(sub_exposure), len(all_ids), _p(all_ids), _p(all_phase),
                                            _p(all_exposure), float(now), float(duration), _p(infectious),
                                            C.addressof(n1), _p(infected), C.addressof(n2), C.addressof(p))
    return dict(calculated=bool(calc), infectious=infectious[:n1.value].copy(), infected=infected[:n2.value].copy(),
                p=p.value)


class Sim:
    """Sequential event-by-event simulation (the medlabs driver contract C1-C8)."""

    def __init__(self, model, trigger, seed):
        self._h = lib().orc_sim_create(model, trigger, seed)
        self.model = model
        self.n_person = 0

    def close(self):
        if self._h:
            lib().orc_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except TypeError:       # interpreter shutdown: the module globals are already gone
            pass

    def set_area(self, props):
        lib().orc_sim_set_area(self._h, C.byref(props))

    def set_dist(self, props):
        lib().orc_sim_set_dist(self._h, C.byref(props))

    def set_prog(self, props):
        lib().orc_sim_set_prog(self._h, C.byref(props))

    def set_location_types(self, infect_sub, corr):
        a = _arr(infect_sub, np.uint8)
        b = _arr(corr, np.float64)
        lib().orc_sim_set_location_types(self._h, len(a), _p(a), _p(b))

    def set_locations(self, type_, n_sub, total_area):
        a = _arr(type_, np.uint8)
        b = _arr(n_sub, np.int16)
        c = _arr(total_area, np.float32)
        lib().orc_sim_set_locations(self._h, len(a), _p(a), _p(b), _p(c))

    def set_persons(self, age, female):
        a = _arr(age, np.int8)
        b = _arr(female, np.uint8)
        self.n_person = len(a)
        lib().orc_sim_set_persons(self._h, len(a), _p(a), _p(b))

    def set_roles(self, role):
        a = _arr(role, np.uint8)
        lib().orc_sim_set_roles(self._h, len(a), _p(a))

    def set_rate_locations(self, location, rate_factor, rate):
        a, b, c = _arr(location, np.int32), _arr(rate_factor, np.float64), _arr(rate, np.float64)
        lib().orc_sim_set_rate_locations(self._h, len(a), _p(a), _p(b), _p(c))

    def set_parameter(self, name, value, at_time):
        return lib().orc_sim_set_parameter(self._h, name.encode(), float(value), float(at_time))

    def expose(self, person, at_time):
        a = _arr(person, np.int32)
        b = _arr(at_time, np.float64)
        lib().orc_sim_expose(self._h, len(a), _p(a), _p(b))

    def add_visits(self, person, loc, sub, t_in, t_out):
        a = _arr(person, np.int32)
        b = _arr(loc, np.int32)
        c = _arr(sub, np.int16)
        d = _arr(t_in, np.float64)
        e = _arr(t_out, np.float64)
        lib().orc_sim_add_visits(self._h, len(a), _p(a), _p(b), _p(c), _p(d), _p(e))

    def run(self, t_until):
        lib().orc_sim_run(self._h, float(t_until))

    def infections(self):
        n = lib().orc_sim_n_infections(self._h)
        person = np.zeros(n, np.int32)
        loc = np.zeros(n, np.int32)
        sub = np.zeros(n, np.int16)
        t = np.zeros(n, np.float64)
        p = np.zeros(n, np.float64)
        lib().orc_sim_get_infections(self._h, _p(person), _p(loc), _p(sub), _p(t), _p(p))
        return dict(person=person, loc=loc, sub=sub, time=t, p=p)

    def transitions(self):
        n = lib().orc_sim_n_transitions(self._h)
        person = np.zeros(n, np.int32)
        phase = np.zeros(n, np.uint8)
        t = np.zeros(n, np.float64)
        lib().orc_sim_get_transitions(self._h, _p(person), _p(phase), _p(t))
        return dict(person=person, phase=phase, time=t)

    def phase_counts(self):
        c = np.zeros(8, np.int64)
        lib().orc_sim_get_phase_counts(self._h, _p(c))
        return c

    def agent_state(self):
        n = self.n_person
        phase = np.zeros(n, np.uint8)
        exposure = np.zeros(n, np.float32)
        next_time = np.zeros(n, np.float64)
        next_phase = np.zeros(n, np.uint8)
        lib().orc_sim_get_agent_state(self._h, _p(phase), _p(exposure), _p(next_time), _p(next_phase))
        return dict(phase=phase, exposure=exposure, next_time=next_time, next_phase=next_phase)

    def n_evaluations(self):
        return lib().orc_sim_n_evaluations(self._h)

    def n_draws(self):
        return lib().orc_sim_n_draws(self._h)

    def last_calc(self, loc, sub):
        return lib().orc_sim_last_calc(self._h, int(loc), int(sub))
