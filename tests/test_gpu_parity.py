"""GPU parity: the windowed CUDA engine (through the C ABI) against the sequential CPU oracle on the same stays."""
import numpy as np
import pytest

from common import Scenario, assert_same_run, load_props, run_gpu, run_oracle
from heros_b200 import _lib

pytestmark = pytest.mark.gpu

MODELS = {"area": _lib.MODEL_AREA, "distance": _lib.MODEL_DISTANCE}
TRIGGERS = {"exit": _lib.TRIGGER_EXIT_ONLY, "both": _lib.TRIGGER_ENTER_AND_EXIT}


def hot_props(model):
    """Reference alpha parameters with the base contagiousness raised so that a few-hundred-person city produces
    hundreds of infections within two weeks."""
    p = load_props(model)
    if model == "area":
        p["covidT_area.contagiousness"] = "4.0"
    else:
        p["covidT_dist.alpha"] = "1.5"
    return p


@pytest.mark.parametrize("model", ["area", "distance"])
@pytest.mark.parametrize("trigger", ["exit", "both"])
def test_small_city_matches_oracle(model, trigger):
    scn = Scenario(seed=11, n_persons=600, n_households=220, n_workplaces=25, n_shops=6, days=14)
    props = hot_props(model)
    ora = run_oracle(scn, props, MODELS[model], TRIGGERS[trigger], seed=111)
    gpu = run_gpu(scn, props, MODELS[model], TRIGGERS[trigger], seed=111)
    assert len(ora["infections"]["person"]) > 50
    assert_same_run(gpu, ora)
    assert sum(s["evaluations"] for s in gpu["stats"]) == ora["evaluations"]
    assert sum(s["near_ties"] for s in gpu["stats"]) == 0
