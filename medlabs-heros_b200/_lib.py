"""ctypes binding of libheros_gpu.so (include/heros_gpu.h).

The library is the product: if it is missing or cannot be loaded this module raises -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# HEROS_GPU_LIB: another build of the same library (csrc/Makefile: `make debug` = device-side bounds checks)
LIB_PATH = os.environ.get("HEROS_GPU_LIB") or os.path.join(PKG_DIR, "libheros_gpu.so")

ABI_VERSION = 1
OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_CAPACITY, ERR_UNKNOWN_PARAMETER = 0, -1, -2, -3, -4, -5
MODEL_AREA, MODEL_DISTANCE = 0, 1
TRIGGER_EXIT_ONLY, TRIGGER_ENTER_AND_EXIT = 0, 1
FLAG_EXACT_STATS = 1
FLAG_PHASE_CLOCKS = 2
FLAG_RESIDENT_INPUTS = 4
DIST_CONSTANT, DIST_UNIFORM, DIST_TRIANGULAR = 0, 1, 2
PHASE_NAMES = ("Susceptible", "Exposed", "Infected-Asymptomatic", "Infected-Symptomatic", "Hospitalized", "ICU",
               "Dead", "Recovered")  # Covid19Progression.java:130-137


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("model", C.c_int32), ("trigger", C.c_int32), ("device", C.c_int32),
                ("seed", C.c_uint64), ("window_hours", C.c_double), ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class AreaParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("contagiousness", "beta", "t_e_min", "t_e_mode", "t_e_max", "calculation_threshold")]


class DistParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("L", "I", "C", "v_max", "v_0", "r", "psi", "alpha", "mu", "calculation_threshold")]


class Duration(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("a", C.c_double), ("b", C.c_double), ("c", C.c_double)]


class ProgParams(C.Structure):
    _fields_ = [("fraction", C.c_double * 128 * 2 * 5), ("period", Duration * 10)]


class Activity(C.Structure):
    _fields_ = [("kind", C.c_int32), ("locator", C.c_int32), ("choice", C.c_int32), ("dist", C.c_int32),
                ("min", C.c_double), ("mode", C.c_double), ("max", C.c_double), ("until", C.c_double)]


class StepStats(C.Structure):
    _fields_ = [("windows", C.c_int64), ("visits", C.c_int64), ("sublocations", C.c_int64),
                ("evaluations", C.c_int64), ("draws", C.c_int64), ("infections", C.c_int64),
                ("transitions", C.c_int64), ("gpu_launches", C.c_int64), ("gpu_ms", C.c_double),
                ("transmit_ms", C.c_double), ("near_ties", C.c_int64), ("progress_ms", C.c_double),
                ("bucket_ms", C.c_double), ("transmit_small_ms", C.c_double), ("transmit_big_ms", C.c_double),
                ("resolve_ms", C.c_double), ("retire_ms", C.c_double), ("big_sublocations", C.c_int64),
                ("big_stays", C.c_int64), ("bucket_count_ms", C.c_double), ("bucket_scan_ms", C.c_double),
                ("bucket_scatter_ms", C.c_double), ("group_done_ms", C.c_double * 4), ("sub32_done_ms", C.c_double)]

    def as_dict(self):
        out = {n: getattr(self, n) for n, _ in self._fields_}
        out["group_done_ms"] = list(self.group_done_ms)
        return out


# every symbol include/heros_gpu.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "heros_abi_version": (C.c_int32, []),
    "heros_create": (C.c_int32, [C.POINTER(Config), C.POINTER(_P)]),
    "heros_destroy": (C.c_int32, [_P]),
    "heros_last_error": (C.c_char_p, [_P]),
    "heros_set_location_types": (C.c_int32, [_P, C.c_int32, _P, _P, _P, _P, _P]),
    "heros_set_locations": (C.c_int32, [_P, C.c_int32, _P, _P, _P, _P, _P]),
    "heros_set_persons": (C.c_int32, [_P, C.c_int32, _P, _P, _P, _P, _P, _P]),
    "heros_set_capacity_diversion": (C.c_int32, [_P, C.c_int32, _P]),
    "heros_set_rate_locations": (C.c_int32, [_P, C.c_int32, _P, _P, _P]),
    "heros_set_transmission_area": (C.c_int32, [_P, C.POINTER(AreaParams)]),
    "heros_set_transmission_distance": (C.c_int32, [_P, C.POINTER(DistParams)]),
    "heros_set_progression": (C.c_int32, [_P, C.POINTER(ProgParams)]),
    "heros_set_parameter": (C.c_int32, [_P, C.c_char_p, C.c_double, C.c_double]),
    "heros_expose": (C.c_int32, [_P, C.c_int32, _P, _P]),
    "heros_seed_infections": (C.c_int32, [_P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "heros_submit_visits": (C.c_int32, [_P, C.c_int64, _P, _P, _P, _P, _P]),
    "heros_submit_visits_begin": (C.c_int32, [_P, C.c_int64, _P, _P, _P, _P, _P]),
    "heros_submit_visits_end": (C.c_int32, [_P]),
    "heros_submit_visits_device": (C.c_int32, [_P, C.c_int64, _P, _P, _P, _P, _P]),
    "heros_set_schedule": (C.c_int32, [_P, C.c_int32, _P, _P, _P, C.c_int32, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32,
                                       _P, _P, _P, _P, C.c_uint64]),
    "heros_advance_schedule": (C.c_int32, [_P, C.c_int32]),
    "heros_validate_schedule_tables": (C.c_int32, [C.c_int32, _P, C.c_int32, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P,
                                                   C.c_int32, _P, _P, C.c_int32, C.c_char_p, C.c_int32]),
    "heros_snapshot_size": (C.c_int32, [_P, C.POINTER(C.c_int64)]),
    "heros_snapshot_save": (C.c_int32, [_P, C.c_int64, _P, C.POINTER(C.c_int64)]),
    "heros_snapshot_load": (C.c_int32, [_P, C.c_int64, _P]),
    "heros_sample_occupancy": (C.c_int32, [_P, C.c_int32, _P, C.c_int64, _P, C.c_int64, _P]),
    "heros_set_closure_policy": (C.c_int32, [_P, C.c_int32, C.c_double, C.c_double, C.c_int32]),
    "heros_debug_timeline": (C.c_int32, [_P, C.c_int32, _P, _P, C.POINTER(C.c_int64)]),
    "heros_debug_get_stays": (C.c_int32, [_P, C.c_int64, _P, _P, _P, _P, _P, C.POINTER(C.c_int64)]),
    "heros_step": (C.c_int32, [_P, C.c_double, C.POINTER(StepStats)]),
    "heros_step_async": (C.c_int32, [_P, C.c_double]),
    "heros_wait": (C.c_int32, [_P, C.POINTER(StepStats)]),
    "heros_window_finish_async": (C.c_int32, [_P]),
    "heros_window_abort": (C.c_int32, [_P]),
    "heros_set_id_bases": (C.c_int32, [_P, C.c_int32, C.c_int32]),
    "heros_window_begin": (C.c_int32, [_P, C.c_double]),
    "heros_export_agent_words": (C.c_int32, [_P, C.c_int64, _P, _P, _P]),
    "heros_submit_guest_visits_device": (C.c_int32, [_P, C.c_int64, _P, _P, _P, _P, _P, _P, _P]),
    "heros_window_transmit": (C.c_int32, [_P]),
    "heros_export_guest_candidates": (C.c_int32, [_P, C.c_int64, _P, _P, _P, _P, _P, C.POINTER(C.c_int64)]),
    "heros_import_candidates_device": (C.c_int32, [_P, C.c_int64, _P, _P, _P, _P, _P]),
    "heros_window_finish": (C.c_int32, [_P, C.POINTER(StepStats)]),
    "heros_guest_block_bytes": (C.c_int64, [C.c_int64]),
    "heros_candidate_block_bytes": (C.c_int64, [C.c_int64]),
    "heros_pack_guest_blocks_device": (C.c_int32, [_P, C.c_int32, _P, _P, _P, _P, _P, _P, C.c_int64, _P]),
    "heros_submit_guest_blocks_device": (C.c_int32, [_P, C.c_int32, C.c_int64, _P]),
    "heros_export_candidate_blocks_device": (C.c_int32, [_P, C.c_int32, _P, C.c_int64, _P]),
    "heros_import_candidate_blocks_device": (C.c_int32, [_P, C.c_int32, C.c_int64, _P]),
    "heros_exchange_create": (C.c_int32, [_P, C.c_int32, C.c_int32, C.c_int64, C.c_int64, _P, C.POINTER(_P)]),
    "heros_exchange_connect": (C.c_int32, [_P, _P, _P]),
    "heros_exchange_set_sources": (C.c_int32, [_P, C.c_uint64]),
    "heros_window_run_peers": (C.c_int32, [_P, C.c_double, _P, _P, _P, _P, _P, _P, _P, C.POINTER(StepStats)]),
    "heros_window_run_peers_async": (C.c_int32, [_P, C.c_double, _P, _P, _P, _P, _P, _P, _P]),
    "heros_window_send_guests": (C.c_int32, [_P, _P, _P, _P, _P, _P, _P]),
    "heros_window_recv_guests": (C.c_int32, [_P]),
    "heros_window_send_candidates": (C.c_int32, [_P, _P]),
    "heros_window_recv_candidates": (C.c_int32, [_P]),
    "heros_get_infections": (C.c_int32, [_P, C.c_int64, C.c_int64, _P, _P, _P, _P, _P, C.POINTER(C.c_int64)]),
    "heros_window_results": (C.c_int32, [_P, C.c_int64, _P, C.POINTER(C.c_int64)]),
    "heros_get_infections_range": (C.c_int32, [_P, C.c_int64, C.c_int64, _P, _P, _P, _P, _P]),
    "heros_get_infections_unordered": (C.c_int32, [_P, C.c_int64, C.c_int64, _P, _P, _P, _P, _P, C.POINTER(C.c_int64)]),
    "heros_get_transitions": (C.c_int32, [_P, C.c_int64, C.c_int64, _P, _P, _P, C.POINTER(C.c_int64)]),
    "heros_get_phase_counts": (C.c_int32, [_P, _P]),
    "heros_get_count_series": (C.c_int32, [_P, C.c_int64, _P, _P, C.POINTER(C.c_int64)]),
    "heros_get_agent_state": (C.c_int32, [_P, _P, _P, _P, _P]),
    "heros_set_agent_state": (C.c_int32, [_P, C.c_int32, _P, _P, _P, C.c_double]),
    "heros_get_last_calculation": (C.c_int32, [_P, C.c_int32, C.c_int16, C.POINTER(C.c_double)]),
    "heros_get_time": (C.c_int32, [_P, C.POINTER(C.c_double)]),
    "heros_get_stream": (C.c_int32, [_P, C.POINTER(_P)]),
    "heros_get_phase_clocks": (C.c_int32, [_P, _P]),
    "heros_infect_people": (C.c_int32, [_P, C.c_int32, C.c_int32, _P, C.c_int32, _P, C.c_double, C.c_double,
                                        C.POINTER(C.c_int32), _P, C.POINTER(C.c_int32), _P, C.POINTER(C.c_int32),
                                        C.POINTER(C.c_double)]),
}

_lib = None


class HerosError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libheros_gpu status {code}: {message}")
        self.code = code


def load() -> C.CDLL:
    """Load libheros_gpu.so and bind every exported symbol.  Raises if the CUDA library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  The HEROS engine has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if lib.heros_abi_version() != ABI_VERSION:
            raise ImportError("libheros_gpu.so ABI version mismatch; rebuild")
        _lib = lib
    return _lib
