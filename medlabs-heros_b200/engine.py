"""numpy-facing wrapper of the C ABI (include/heros_gpu.h): one ``HerosEngine`` = one ``heros_handle``."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _lib
from ._lib import HerosError


def _arr(x, dtype):
    return np.ascontiguousarray(np.asarray(x, dtype=dtype))


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class HerosEngine:
    """The GPU engine for the transmission + progression hot path.  All compute happens in libheros_gpu.so."""

    def __init__(self, model: int, seed: int, trigger: int = _lib.TRIGGER_EXIT_ONLY, device: int = 0,
                 window_hours: float = 24.0, flags: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        cfg = _lib.Config(_lib.ABI_VERSION, int(model), int(trigger), int(device), int(seed) & (2 ** 64 - 1),
                          float(window_hours), int(flags), 0)
        rc = self._lib.heros_create(C.byref(cfg), C.byref(self._h))
        if rc != _lib.OK:
            msg = self._lib.heros_last_error(None).decode()
            self._h = C.c_void_p()
            raise HerosError(rc, msg)
        self.model = int(model)
        self.n_persons = 0
        self.n_locations = 0

    # ---- plumbing ----
    def _check(self, rc: int):
        if rc != _lib.OK:
            raise HerosError(rc, self._lib.heros_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.heros_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- static inputs ----
    def set_location_types(self, infect_sub, correction_factor_area, cap_constrained=None, cap_persons_per_m2=None,
                           size_factor=None):
        a = _arr(infect_sub, np.uint8)
        b = _arr(correction_factor_area, np.float64)
        c = None if cap_constrained is None else _arr(cap_constrained, np.uint8)
        d = None if cap_persons_per_m2 is None else _arr(cap_persons_per_m2, np.float64)
        e = None if size_factor is None else _arr(size_factor, np.float64)
        self._check(self._lib.heros_set_location_types(self._h, len(a), _p(a), _p(b), _p(c), _p(d), _p(e)))

    def set_locations(self, type_, n_sub, total_area, lat=None, lon=None):
        a = _arr(type_, np.uint8)
        b = _arr(n_sub, np.int16)
        c = _arr(total_area, np.float32)
        d = None if lat is None else _arr(lat, np.float32)
        e = None if lon is None else _arr(lon, np.float32)
        self._check(self._lib.heros_set_locations(self._h, len(a), _p(a), _p(b), _p(c), _p(d), _p(e)))
        self.n_locations = len(a)
        self.location_n_sub = b.copy()
        # first entry of every location in the location-major sub-location order (a location without sub-locations
        # still owns one entry)
        self.location_first_sub = np.concatenate([[0], np.cumsum(np.maximum(b.astype(np.int64), 1))])

    def set_persons(self, age, female, role=None, home_loc=None, home_sub=None, work_loc=None):
        a = _arr(age, np.int8)
        b = _arr(female, np.uint8)
        c = None if role is None else _arr(role, np.uint8)
        d = None if home_loc is None else _arr(home_loc, np.int32)
        e = None if home_sub is None else _arr(home_sub, np.int16)
        f = None if work_loc is None else _arr(work_loc, np.int32)
        self._check(self._lib.heros_set_persons(self._h, len(a), _p(a), _p(b), _p(c), _p(d), _p(e), _p(f)))
        self.n_persons = len(a)

    # ---- parameters ----
    def set_transmission_area(self, params: _lib.AreaParams):
        self._check(self._lib.heros_set_transmission_area(self._h, C.byref(params)))

    def set_transmission_distance(self, params: _lib.DistParams):
        self._check(self._lib.heros_set_transmission_distance(self._h, C.byref(params)))

    def set_progression(self, params: _lib.ProgParams):
        self._check(self._lib.heros_set_progression(self._h, C.byref(params)))

    def set_parameter(self, name: str, value: float, at_time: float) -> int:
        """Returns the status instead of raising for unknown names (the reference only prints a message)."""
        rc = self._lib.heros_set_parameter(self._h, name.encode(), float(value), float(at_time))
        if rc not in (_lib.OK, _lib.ERR_UNKNOWN_PARAMETER):
            self._check(rc)
        return rc

    # ---- dynamic inputs ----
    def expose(self, person, at_time):
        a = _arr(person, np.int32)
        b = np.broadcast_to(_arr(at_time, np.float64), a.shape).copy()
        self._check(self._lib.heros_expose(self._h, len(a), _p(a), _p(b)))

    def seed_infections(self, n_infected: int, min_age: int = 0, max_age: int = 127) -> np.ndarray:
        out = np.zeros(n_infected, np.int32)
        self._check(self._lib.heros_seed_infections(self._h, n_infected, min_age, max_age, _p(out)))
        return out

    def set_rate_locations(self, location, rate_factor, rate):
        """heros_set_rate_locations: the LocationProbBased rows of infection_rates.csv (location id, infection_rate_factor,
        infection_rate; -1 = not given)."""
        a, b, c = _arr(location, np.int32), _arr(rate_factor, np.float64), _arr(rate, np.float64)
        if not (len(a) == len(b) == len(c)):
            raise ValueError("rate-location columns differ in length")
        self._check(self._lib.heros_set_rate_locations(self._h, len(a), _p(a), _p(b), _p(c)))

    def submit_visits(self, person, loc, sub, t_in, t_out):
        a = _arr(person, np.int32)
        b = _arr(loc, np.int32)
        c = _arr(sub, np.int16)
        d = _arr(t_in, np.float64)
        e = _arr(t_out, np.float64)
        if not (len(a) == len(b) == len(c) == len(d) == len(e)):
            raise ValueError("visit columns differ in length")
        self._check(self._lib.heros_submit_visits(self._h, len(a), _p(a), _p(b), _p(c), _p(d), _p(e)))

    def submit_visits_ptr(self, n: int, person: int, loc: int, sub: int, t_in: int, t_out: int, device: bool):
        """Raw-pointer submission (pinned host buffers or device buffers, e.g. torch tensors' data_ptr())."""
        fn = self._lib.heros_submit_visits_device if device else self._lib.heros_submit_visits
        self._check(fn(self._h, n, C.c_void_p(person), C.c_void_p(loc), C.c_void_p(sub), C.c_void_p(t_in),
                       C.c_void_p(t_out)))

    def submit_visits_begin_ptr(self, n: int, person: int, loc: int, sub: int, t_in: int, t_out: int):
        """heros_submit_visits_begin: the host -> device copies start, the call returns at once; the (pinned) buffers
        must stay untouched until ``submit_visits_end``."""
        self._check(self._lib.heros_submit_visits_begin(self._h, n, C.c_void_p(person), C.c_void_p(loc), C.c_void_p(sub),
                                                        C.c_void_p(t_in), C.c_void_p(t_out)))

    def submit_visits_end(self):
        """heros_submit_visits_end: the stays are taken in behind their copies; returns when the buffers have been read."""
        self._check(self._lib.heros_submit_visits_end(self._h))

    # ---- device-side activity schedule ----
    def set_schedule(self, tables: Dict[str, np.ndarray], seed: int):
        """``tables``: the arrays of ``scenario.compile_schedule`` (heros_set_schedule)."""
        t = tables
        acts = np.ascontiguousarray(t["activities"])
        assert acts.dtype.itemsize == C.sizeof(_lib.Activity)
        ps, pc = _arr(t["pattern_start"], np.int32), _arr(t["pattern_count"], np.int32)
        cs, ct, cc = _arr(t["choice_start"], np.int32), _arr(t["choice_type"], np.int32), _arr(t["choice_cum"], np.float64)
        near, ncand, cand = _arr(t["nearest"], np.int32), _arr(t["n_cand"], np.int32), _arr(t["cand"], np.int32)
        wsub = _arr(t["work_sub"], np.int16)
        self._check(self._lib.heros_set_schedule(self._h, len(acts), _p(acts), _p(ps), _p(pc), len(cs) - 1, _p(cs), _p(ct),
                                                 _p(cc), int(t["n_types"]), int(t["accommodation_type"]),
                                                 int(t["n_candidates"]), _p(near), _p(ncand), _p(cand), _p(wsub),
                                                 int(seed) & (2 ** 64 - 1)))
        divert = t.get("divert")
        if divert is not None:
            self.set_capacity_diversion(np.asarray(divert)[:np.asarray(t["nearest"]).shape[0]])  # a shard: its own locations

    def set_capacity_diversion(self, divert):
        """heros_set_capacity_diversion: per location, the share of the visits of a *Cap locator that goes to the next
        candidate (``scenario.capacity_diversion``); None removes the table."""
        if divert is None:
            self._check(self._lib.heros_set_capacity_diversion(self._h, 0, None))
            return
        d = _arr(divert, np.float64)
        self._check(self._lib.heros_set_capacity_diversion(self._h, len(d), _p(d)))

    def advance_schedule(self, day: int):
        self._check(self._lib.heros_advance_schedule(self._h, int(day)))

    def set_closure_policy(self, location_type: int, fraction_open: float, fraction_activities: float,
                           alternative_type: int):
        """LocationType.setClosurePolicy of the reference (LocationPolicy.java:45) for the device-side schedule."""
        self._check(self._lib.heros_set_closure_policy(self._h, int(location_type), float(fraction_open),
                                                       float(fraction_activities), int(alternative_type)))

    def snapshot(self) -> np.ndarray:
        """The dynamic state of the idle engine as one uint8 array (heros_snapshot_save)."""
        n = C.c_int64(0)
        self._check(self._lib.heros_snapshot_size(self._h, C.byref(n)))
        blob = np.zeros(n.value, np.uint8)
        self._check(self._lib.heros_snapshot_save(self._h, n.value, _p(blob), C.byref(n)))
        return blob[:n.value]

    def restore(self, blob) -> None:
        """heros_snapshot_load into an engine that has been given the same static inputs."""
        b = np.ascontiguousarray(np.frombuffer(blob, np.uint8) if isinstance(blob, (bytes, bytearray)) else blob, np.uint8)
        self._check(self._lib.heros_snapshot_load(self._h, len(b), _p(b)))

    def sample_occupancy(self, times, locations: bool = True, sublocations: bool = True):
        """Persons per location ``[len(times), n_locations]`` and per sub-location ``[len(times), n_sublocations]``
        (location-major, see ``location_first_sub``) at the given times, from the stays held for the coming window
        (heros_sample_occupancy = the hourly samples of HerosModel.locationDump, HerosModel.java:266-299)."""
        t = _arr(times, np.float64)
        n_sub_total = int(self.location_first_sub[-1])
        per_loc = np.zeros((len(t), self.n_locations), np.int32) if locations else None
        per_sub = np.zeros((len(t), n_sub_total), np.int32) if sublocations else None
        self._check(self._lib.heros_sample_occupancy(self._h, len(t), _p(t), self.n_locations, _p(per_loc), n_sub_total,
                                                     _p(per_sub)))
        return per_loc, per_sub

    def debug_stays(self) -> Dict[str, np.ndarray]:
        n = C.c_int64(0)
        self._check(self._lib.heros_debug_get_stays(self._h, 0, None, None, None, None, None, C.byref(n)))
        m = n.value
        person, loc, sub = np.zeros(m, np.int32), np.zeros(m, np.int32), np.zeros(m, np.int16)
        t_in, t_out = np.zeros(m, np.float64), np.zeros(m, np.float64)
        if m:
            self._check(self._lib.heros_debug_get_stays(self._h, m, _p(person), _p(loc), _p(sub), _p(t_in), _p(t_out),
                                                        C.byref(n)))
        return dict(person=person, loc=loc, sub=sub, t_in=t_in, t_out=t_out)

    # ---- hot path ----
    def step(self, t_until: float) -> Dict[str, float]:
        st = _lib.StepStats()
        self._check(self._lib.heros_step(self._h, float(t_until), C.byref(st)))
        return st.as_dict()

    def step_async(self, t_until: float) -> None:
        """heros_step_async: the windows up to ``t_until`` are enqueued; collect them with ``wait``."""
        self._check(self._lib.heros_step_async(self._h, float(t_until)))

    def wait(self) -> Dict[str, float]:
        """heros_wait: everything enqueued has finished; statistics of the windows since the last wait / step."""
        st = _lib.StepStats()
        self._check(self._lib.heros_wait(self._h, C.byref(st)))
        return st.as_dict()

    TIMELINE_IDS = ("k_window_open", "k_scan_lookback", "k_bucket_scatter", "k_window_roll", "k_transmit_stream",
                    "k_transmit_sub<8>", "k_transmit_sub<32>", "k_hot_chain<0>", "k_hot_chain<1>", "k_hot_chain<2>",
                    "k_hot_eval", "k_transmit_total", "k_resolve", "k_submit", "k_advance_day", "k_rate_locations")

    def timeline(self, enable: bool = True):
        """heros_debug_timeline: {kernel: (start_us, end_us)} averaged over the windows since the collection was started
        (relative to the start of the window's first kernel), and the number of windows; (re)starts or stops it."""
        start = (C.c_double * 32)()
        end = (C.c_double * 32)()
        n = C.c_int64(0)
        self._check(self._lib.heros_debug_timeline(self._h, 1 if enable else 0, start, end, C.byref(n)))
        rows = {name: (start[k], end[k]) for k, name in enumerate(self.TIMELINE_IDS) if end[k] >= 0 and n.value > 0}
        return rows, int(n.value)

    def window_abort(self) -> None:
        self._check(self._lib.heros_window_abort(self._h))

    def window_finish_async(self) -> None:
        self._check(self._lib.heros_window_finish_async(self._h))

    # ---- multi-GPU window phases (device pointers, e.g. torch tensors' data_ptr()) ----
    def set_id_bases(self, person_base: int, location_base: int):
        self._check(self._lib.heros_set_id_bases(self._h, int(person_base), int(location_base)))

    def window_begin(self, t_until: float):
        self._check(self._lib.heros_window_begin(self._h, float(t_until)))

    def export_agent_words(self, n: int, person_ptr: int, word_out_ptr: int, ill_end_out_ptr: int):
        self._check(self._lib.heros_export_agent_words(self._h, n, C.c_void_p(person_ptr), C.c_void_p(word_out_ptr),
                                                       C.c_void_p(ill_end_out_ptr)))

    def submit_guest_visits(self, n: int, person: int, word: int, ill_end: int, loc: int, sub: int, t_in: int,
                            t_out: int):
        self._check(self._lib.heros_submit_guest_visits_device(self._h, n, *(C.c_void_p(x) for x in
                                                               (person, word, ill_end, loc, sub, t_in, t_out))))

    def window_transmit(self):
        self._check(self._lib.heros_window_transmit(self._h))

    def export_guest_candidates(self, capacity: int, person: int, loc: int, sub: int, t: int, p: int) -> int:
        n = C.c_int64(0)
        self._check(self._lib.heros_export_guest_candidates(self._h, capacity, *(C.c_void_p(x) for x in
                                                            (person, loc, sub, t, p)), C.byref(n)))
        return n.value

    def import_candidates(self, n: int, person: int, loc: int, sub: int, t: int, p: int):
        self._check(self._lib.heros_import_candidates_device(self._h, n, *(C.c_void_p(x) for x in
                                                             (person, loc, sub, t, p))))

    # ---- the same exchange as fixed-capacity blocks (no host synchronisation inside a window) ----
    def guest_block_bytes(self, capacity: int) -> int:
        return int(self._lib.heros_guest_block_bytes(int(capacity)))

    def candidate_block_bytes(self, capacity: int) -> int:
        return int(self._lib.heros_candidate_block_bytes(int(capacity)))

    def pack_guest_blocks(self, row_offsets, person: int, loc: int, sub: int, t_in: int, t_out: int, capacity: int,
                          blocks_out: int):
        off = _arr(row_offsets, np.int64)
        self._check(self._lib.heros_pack_guest_blocks_device(self._h, len(off) - 1, _p(off), *(C.c_void_p(x) for x in
                                                             (person, loc, sub, t_in, t_out)), int(capacity),
                                                             C.c_void_p(blocks_out)))

    def submit_guest_blocks(self, n_blocks: int, capacity: int, blocks: int):
        self._check(self._lib.heros_submit_guest_blocks_device(self._h, int(n_blocks), int(capacity), C.c_void_p(blocks)))

    def export_candidate_blocks(self, person_bases, capacity: int, blocks_out: int):
        b = _arr(person_bases, np.int32)
        self._check(self._lib.heros_export_candidate_blocks_device(self._h, len(b) - 1, _p(b), int(capacity),
                                                                   C.c_void_p(blocks_out)))

    def import_candidate_blocks(self, n_blocks: int, capacity: int, blocks: int):
        self._check(self._lib.heros_import_candidate_blocks_device(self._h, int(n_blocks), int(capacity),
                                                                   C.c_void_p(blocks)))

    # ---- peer-to-peer exchange over NVLink (CUDA IPC): no collective call ----
    def exchange_create(self, rank: int, n_peers: int, guest_capacity: int, candidate_capacity: int):
        """Returns (64-byte IPC handle as a numpy uint8 array, device pointer of the region)."""
        handle = np.zeros(64, np.uint8)
        region = C.c_void_p()
        self._check(self._lib.heros_exchange_create(self._h, int(rank), int(n_peers), int(guest_capacity),
                                                    int(candidate_capacity), _p(handle), C.byref(region)))
        return handle, int(region.value or 0)

    def exchange_connect(self, handles: Optional[np.ndarray], local_regions=None):
        hs = None if handles is None else np.ascontiguousarray(handles, np.uint8)
        loc = None
        if local_regions is not None:
            loc = (C.c_void_p * len(local_regions))(*[C.c_void_p(int(x)) if x else None for x in local_regions])
        self._check(self._lib.heros_exchange_connect(self._h, _p(hs), loc))

    def exchange_set_sources(self, ranks):
        mask = 0
        for r in ranks:
            mask |= 1 << int(r)
        self._check(self._lib.heros_exchange_set_sources(self._h, mask))

    def window_run_peers(self, t_until: float, row_offsets: np.ndarray, person: int, loc: int, sub: int, t_in: int,
                         t_out: int, person_bases: np.ndarray) -> Dict[str, float]:
        st = _lib.StepStats()
        self._check(self._lib.heros_window_run_peers(self._h, float(t_until), _p(row_offsets),
                                                     *(C.c_void_p(x) for x in (person, loc, sub, t_in, t_out)),
                                                     _p(person_bases), C.byref(st)))
        return st.as_dict()

    def window_run_peers_async(self, t_until: float, row_offsets: np.ndarray, person: int, loc: int, sub: int, t_in: int,
                               t_out: int, person_bases: np.ndarray) -> None:
        """heros_window_run_peers_async: the whole window enqueued, nothing waited for; collect with ``wait``."""
        self._check(self._lib.heros_window_run_peers_async(self._h, float(t_until), _p(row_offsets),
                                                           *(C.c_void_p(x) for x in (person, loc, sub, t_in, t_out)),
                                                           _p(person_bases)))

    def window_send_guests(self, row_offsets, person: int, loc: int, sub: int, t_in: int, t_out: int):
        off = _arr(row_offsets, np.int64)
        self._check(self._lib.heros_window_send_guests(self._h, _p(off), *(C.c_void_p(x) for x in (person, loc, sub, t_in, t_out))))

    def window_recv_guests(self):
        self._check(self._lib.heros_window_recv_guests(self._h))

    def window_send_candidates(self, person_bases):
        b = _arr(person_bases, np.int32)
        self._check(self._lib.heros_window_send_candidates(self._h, _p(b)))

    def window_recv_candidates(self):
        self._check(self._lib.heros_window_recv_candidates(self._h))

    def window_finish(self) -> Dict[str, float]:
        st = _lib.StepStats()
        self._check(self._lib.heros_window_finish(self._h, C.byref(st)))
        return st.as_dict()

    # ---- outputs ----
    def infections(self, first: int = 0, ordered: bool = True) -> Dict[str, np.ndarray]:
        """Infection events from number ``first`` on: ordered by (time, person) (heros_get_infections), or -- cheaper -- in
        the order the device logged them (heros_get_infections_unordered: windows in order, arbitrary inside a window)."""
        fn = self._lib.heros_get_infections if ordered else self._lib.heros_get_infections_unordered
        n = C.c_int64(0)
        self._check(fn(self._h, first, 0, None, None, None, None, None, C.byref(n)))
        m = n.value
        person = np.zeros(m, np.int32)
        loc = np.zeros(m, np.int32)
        sub = np.zeros(m, np.int16)
        t = np.zeros(m, np.float64)
        p = np.zeros(m, np.float64)
        if m:
            self._check(fn(self._h, first, m, _p(person), _p(loc), _p(sub), _p(t), _p(p), C.byref(n)))
        return dict(person=person, loc=loc, sub=sub, time=t, p=p)

    def window_results(self, window: int):
        """heros_window_results: (compartment counts at the end of window ``window``, infection events logged by then);
        waits for that window only."""
        c = np.zeros(8, np.int64)
        n = C.c_int64(0)
        self._check(self._lib.heros_window_results(self._h, int(window), _p(c), C.byref(n)))
        return c, int(n.value)

    def infections_range(self, first: int, n: int) -> Dict[str, np.ndarray]:
        """heros_get_infections_range: log rows [first, first + n) of finished windows, nothing waited for."""
        person, loc, sub = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int16)
        t, p = np.zeros(n, np.float64), np.zeros(n, np.float64)
        self._check(self._lib.heros_get_infections_range(self._h, int(first), int(n), _p(person), _p(loc), _p(sub), _p(t), _p(p)))
        return dict(person=person, loc=loc, sub=sub, time=t, p=p)

    def transitions(self, first: int = 0) -> Dict[str, np.ndarray]:
        n = C.c_int64(0)
        self._check(self._lib.heros_get_transitions(self._h, first, 0, None, None, None, C.byref(n)))
        m = n.value
        person = np.zeros(m, np.int32)
        phase = np.zeros(m, np.uint8)
        t = np.zeros(m, np.float64)
        if m:
            self._check(self._lib.heros_get_transitions(self._h, first, m, _p(person), _p(phase), _p(t), C.byref(n)))
        return dict(person=person, phase=phase, time=t)

    def phase_counts(self) -> np.ndarray:
        c = np.zeros(8, np.int64)
        self._check(self._lib.heros_get_phase_counts(self._h, _p(c)))
        return c

    def count_series(self):
        n = C.c_int64(0)
        self._check(self._lib.heros_get_count_series(self._h, 0, None, None, C.byref(n)))
        m = n.value
        t = np.zeros(m, np.float64)
        c = np.zeros((m, 8), np.int64)
        if m:
            self._check(self._lib.heros_get_count_series(self._h, m, _p(t), _p(c), C.byref(n)))
        return t, c

    def agent_state(self) -> Dict[str, np.ndarray]:
        n = self.n_persons
        phase = np.zeros(n, np.uint8)
        exposure = np.zeros(n, np.float32)
        next_time = np.zeros(n, np.float64)
        next_phase = np.zeros(n, np.uint8)
        self._check(self._lib.heros_get_agent_state(self._h, _p(phase), _p(exposure), _p(next_time), _p(next_phase)))
        return dict(phase=phase, exposure=exposure, next_time=next_time, next_phase=next_phase)

    def set_agent_state(self, person, phase, exposure, now: float = 0.0):
        a = _arr(person, np.int32)
        b = _arr(phase, np.uint8)
        c = _arr(exposure, np.float32)
        self._check(self._lib.heros_set_agent_state(self._h, len(a), _p(a), _p(b), _p(c), float(now)))

    def last_calculation(self, loc: int, sub: int) -> float:
        out = C.c_double(0.0)
        self._check(self._lib.heros_get_last_calculation(self._h, int(loc), int(sub), C.byref(out)))
        return out.value

    def phase_clocks(self) -> np.ndarray:
        out = np.zeros((4, 16), np.uint64)
        self._check(self._lib.heros_get_phase_clocks(self._h, _p(out)))
        return out

    @property
    def stream(self) -> int:
        """cudaStream_t of the engine as an integer (e.g. for ``torch.cuda.ExternalStream``)."""
        out = C.c_void_p()
        self._check(self._lib.heros_get_stream(self._h, C.byref(out)))
        return int(out.value or 0)

    @property
    def time(self) -> float:
        out = C.c_double(0.0)
        self._check(self._lib.heros_get_time(self._h, C.byref(out)))
        return out.value

    def infect_people(self, location: int, persons_in_sublocation, now: float, duration: float, all_persons=None):
        """1:1 shim of DiseaseTransmission.infectPeople evaluated on the GPU; returns a dict."""
        a = _arr(persons_in_sublocation, np.int32)
        b = None if all_persons is None else _arr(all_persons, np.int32)
        cap = max(len(a), 0 if b is None else len(b), 1)
        infectious = np.zeros(cap, np.int32)
        infected = np.zeros(cap, np.int32)
        calc, n1, n2, p = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_double(0.0)
        self._check(self._lib.heros_infect_people(self._h, int(location), len(a), _p(a), 0 if b is None else len(b),
                                                  _p(b), float(now), float(duration), C.byref(calc), _p(infectious),
                                                  C.byref(n1), _p(infected), C.byref(n2), C.byref(p)))
        return dict(calculated=bool(calc.value), infectious=infectious[:n1.value].copy(),
                    infected=infected[:n2.value].copy(), p=p.value)
