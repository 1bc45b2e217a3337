"""heros_infect_people: the 1:1 GPU shim of DiseaseTransmission.infectPeople against the oracle's restatement,
through the mirror of the reference's plug-in classes."""
import math

import numpy as np
import pytest

import heros_b200 as hb
from common import load_props
from oracle import binding as ob
from test_oracle import area as area_props_o, dist_props as dist_props_o

pytestmark = pytest.mark.gpu
S, E, IA, IS, H, ICU, D, R = range(8)


def build_model(model, n_persons=64, extra=None):
    props = load_props(model)
    props.update(extra or {})
    m = hb.HerosModel(props, seed=111)
    # types: 0 infectSub, 1 not; locations: 0 (1 sub, 10 m2) 1 (4 subs, 120 m2 total) 2 (type 1, 1 sub) 3 degenerate
    m.engine.set_location_types([1, 0], [1.0, 0.7])
    m.engine.set_locations([0, 0, 1, 0, 0], [1, 4, 1, 0, 1], [10.0, 120.0, 300000.0, 0.0, -1.0])
    m.engine.set_persons(np.full(n_persons, 40, np.int8), np.zeros(n_persons, np.uint8))
    prog, trans = hb.construct_disease_model(m)
    return m, prog, trans


def oracle_call(model, props_o, loc_cfg, ids, phase, exposure, now, duration, **kw):
    infect_sub, corr, area, nsub = loc_cfg
    return ob.infect_people(model, props_o, 111, infect_sub, corr, area, nsub, ids, phase, exposure, now, duration, **kw)


@pytest.mark.parametrize("model", ["area", "distance"])
def test_shim_matches_oracle_on_random_sets(model):
    m, prog, trans = build_model(model, 64, {"covidT_area.contagiousness": "3.0", "covidT_dist.alpha": "0.9"})
    rng = np.random.default_rng(4)
    phase = rng.choice([S, S, S, E, IA, IS, H, R, D], 64).astype(np.uint8)
    now = 500.0
    exposure = (now - rng.uniform(0, 260, 64)).astype(np.float32)
    m.engine.set_agent_state(np.arange(64), phase, exposure, now=now)
    props_o = area_props_o(pB=3.0) if model == "area" else dist_props_o(alpha=0.9)
    kind = ob.MODEL_AREA if model == "area" else ob.MODEL_DIST
    cfgs = {0: (1, 1.0, 10.0, 1), 1: (1, 1.0, 120.0, 4), 2: (0, 0.7, 300000.0, 1), 3: (1, 1.0, 0.0, 0), 4: (1, 1.0, -1.0, 1)}
    n_pos = 0
    for trial in range(60):
        loc = int(rng.integers(0, 5))
        ids = rng.choice(64, int(rng.integers(1, 30)), replace=False).astype(np.int32)
        duration = float(rng.choice([0.001, 1 / 60, 0.5, 2.0, 9.0]))
        rec = trans.infectPeople(loc, ids, duration, now=now)
        want = oracle_call(kind, props_o, cfgs[loc], ids, phase[ids], exposure[ids], now, duration)
        assert rec.isCalculated() == want["calculated"]
        assert rec.infectiousPersons == want["infectious"].tolist()     # same order as the set iteration
        assert rec.infectedPersons == want["infected"].tolist()
        # the shim sums in set order like the Java loop, so p is bit-identical
        assert (math.isnan(rec.pInfection) and math.isnan(want["p"])) or rec.pInfection == want["p"]
        n_pos += rec.pInfection > 0
    assert n_pos > 10
    m.close()


@pytest.mark.parametrize("model", ["area", "distance"])
def test_total_location_branch_is_accepted_by_the_windowed_engine(model):
    """infectSub = false and nSub >= 2 (the total-location branch): heros_set_locations used to refuse it, the windowed
    engine now evaluates such locations per location (tests/test_gpu_parity.py::test_total_location_branch_in_windows);
    the 1:1 shim on the same tables is covered by test_shim_matches_oracle_call_by_call."""
    props = load_props(model)
    props["covidT_area.contagiousness"] = "30.0"
    m = hb.HerosModel(props, seed=111)
    m.engine.set_location_types([1, 0], [1.0, 1.0])
    m.engine.set_locations([1], [3], [60.0])
    m.close()


def test_mirror_classes_and_expose():
    m, prog, trans = build_model("area", 16)
    assert prog.getDiseasePhases() == list(hb.PHASE_NAMES)
    assert isinstance(trans, hb.Covid19TransmissionArea) and m.getDiseaseTransmission() is trans
    prog.expose([3, 5], hb.Covid19Progression.exposed, at_time=0.0)
    prog.expose(3, at_time=1.0)  # already exposed: ignored, like the isSusceptible() check of the seeding loop
    c = m.engine.phase_counts()
    assert c[0] == 14 and c[1] == 2
    rec = trans.infectPeople(0, [3, 5, 7], 2.0, now=70.0)
    assert rec.calculated and rec.infectiousPersons == [3, 5] and rec.exposurePhase == "Exposed" and rec.location == 0
    trans.setParameter("psi", 2.0)  # prints "Unrecognized variable name psi" (TArea:254), no exception
    with pytest.raises(ValueError):
        prog.expose([1], "Recovered")
    m.close()
