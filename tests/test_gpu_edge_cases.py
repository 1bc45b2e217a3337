"""Edge cases of the hot path against the oracle: empty inputs, everybody leaving at the same instant, whole groups
with identical times, degenerate and invalid stays (rejected loudly, the rest of the run unchanged)."""
import numpy as np
import pytest

import heros_b200 as hb
from common import assert_same_run, load_props, run_gpu, run_oracle

pytestmark = pytest.mark.gpu


class Hand:
    """A hand-made scenario with the attributes common.run_oracle / run_gpu read."""

    def __init__(self, n_persons, loc_nsub, stays, seeds, days, per_sub_area=40.0, seed=5):
        rng = np.random.default_rng(seed)
        self.days = days
        self.type_infect_sub = np.array([1], np.uint8)
        self.type_corr = np.array([1.0], np.float64)
        self.loc_nsub = np.asarray(loc_nsub, np.int16)
        self.loc_type = np.zeros(len(self.loc_nsub), np.uint8)
        self.loc_area = (per_sub_area * np.maximum(self.loc_nsub, 1)).astype(np.float32)
        self.n_loc = len(self.loc_nsub)
        self.n_persons = n_persons
        self.age = rng.integers(20, 60, n_persons).astype(np.int8)
        self.female = (rng.random(n_persons) < 0.5).astype(np.uint8)
        cols = list(zip(*stays)) if len(stays) else [[], [], [], [], []]
        self.person = np.array(cols[0], np.int32)
        self.loc = np.array(cols[1], np.int32)
        self.sub = np.array(cols[2], np.int16)
        self.t_in = np.array(cols[3], np.float64)
        self.t_out = np.array(cols[4], np.float64)
        self.seeds = np.asarray(seeds, np.int32)


def hot(model):
    """parameters under which the seeds are contagious from the start of day 3 and infect quickly"""
    props = load_props(model)
    if model == "area":
        props["covidT_area.contagiousness"] = "40.0"
    else:
        props["covidT_dist.alpha"] = "8.0"
    return props, hb.MODEL_AREA if model == "area" else hb.MODEL_DISTANCE


@pytest.mark.parametrize("model", ["area", "distance"])
def test_no_stays_at_all(model):
    props, kind = hot(model)
    scn = Hand(50, [3, 1, 2], [], seeds=[1, 7, 30], days=6)
    ora = run_oracle(scn, props, kind, hb.TRIGGER_EXIT_ONLY, 42)
    gpu = run_gpu(scn, props, kind, hb.TRIGGER_EXIT_ONLY, 42)
    assert_same_run(gpu, ora)
    assert len(gpu["infections"]["person"]) == 0 and len(gpu["transitions"]["person"]) >= 3
    assert sum(s["visits"] for s in gpu["stats"]) == 0
    gpu["engine"].close()


@pytest.mark.parametrize("model,trigger", [("area", 0), ("distance", 0), ("area", 1), ("distance", 1)])
def test_everybody_leaves_at_the_same_instant(model, trigger):
    """Open stays from t = 0 (nobody moves: no evaluation), closed for everybody at exactly t = 80 h: one calculated
    evaluation per sub-location with duration 80 h, every later event of the instant has duration 0."""
    props, kind = hot(model)
    n, per_home = 240, 6
    home = [(p, p // (per_home * 4), (p // per_home) % 4, 0.0, 80.0) for p in range(n)]
    back = [(p, p // (per_home * 4), (p // per_home) % 4, 80.0 + 0.5, np.inf) for p in range(n)]
    scn = Hand(n, [4] * (n // (per_home * 4)), home + back, seeds=range(0, n, 9), days=5)
    ora = run_oracle(scn, props, kind, trigger, 7)
    for submit in ("all", "daily"):
        gpu = run_gpu(scn, props, kind, trigger, 7, submit=submit)
        assert_same_run(gpu, ora)
        gpu["engine"].close()
    assert len(ora["infections"]["person"]) > 10
    assert (ora["infections"]["time"] == 80.0).all()


@pytest.mark.parametrize("model", ["area", "distance"])
def test_whole_groups_with_identical_times(model):
    """60 / 300 / 1500 persons share one sub-location with identical entry and exit times, hour after hour: every
    movement event of an instant but the first has duration 0 (sub-warp, warp and block kernels)."""
    props, kind = hot(model)
    stays, p0 = [], 0
    sizes = (60, 300, 1500, 9)
    for loc, size in enumerate(sizes):
        for p in range(p0, p0 + size):
            stays.append((p, loc, 0, 0.0, 50.0))
            for d in (2, 3):
                for h in range(8, 14):
                    stays.append((p, loc, 0, 24.0 * d + h + 0.25, 24.0 * d + h + 1.0))
        p0 += size
    scn = Hand(p0, [1] * len(sizes), stays, seeds=[0, 1, 70, 71, 72, 400, 401, 402, 403, 1861], days=4, per_sub_area=400.0)
    ora = run_oracle(scn, props, kind, hb.TRIGGER_EXIT_ONLY, 3)
    gpu = run_gpu(scn, props, kind, hb.TRIGGER_EXIT_ONLY, 3)
    assert_same_run(gpu, ora)
    assert len(ora["infections"]["person"]) > 100
    gpu["engine"].close()


def test_degenerate_and_invalid_stays():
    props, kind = hot("area")
    n = 120
    good = []
    for p in range(n):
        good.append((p, p % 5, p % 3, 0.0, 60.0 + (p % 7)))
        good.append((p, (p + 1) % 5, (p + 2) % 3, 61.0 + (p % 7) + 0.5, 90.0 + 0.01 * p))
    degenerate = [(3, 1, 0, 70.0, 70.0), (4, 1, 0, 71.0, 70.5)]                       # ignored (t_out <= t_in)
    bad = [(-1, 0, 0, 1.0, 2.0), (n, 0, 0, 1.0, 2.0), (5, 5, 0, 1.0, 2.0), (5, -2, 0, 1.0, 2.0), (5, 0, 3, 1.0, 2.0),
           (5, 0, -1, 1.0, 2.0), (5, 0, 0, np.nan, 2.0), (5, 0, 0, 1.0, np.nan), (5, 0, 0, -1.0, 2.0)]
    clean = Hand(n, [3] * 5, good, seeds=range(0, n, 11), days=4)
    ora = run_oracle(clean, props, kind, hb.TRIGGER_EXIT_ONLY, 9)
    dirty = Hand(n, [3] * 5, good + degenerate + bad, seeds=range(0, n, 11), days=4)
    from common import make_engine
    eng = make_engine(dirty, props, kind, hb.TRIGGER_EXIT_ONLY, 9)
    eng.expose(dirty.seeds, np.zeros(len(dirty.seeds)))
    eng.submit_visits(dirty.person, dirty.loc, dirty.sub, dirty.t_in, dirty.t_out)
    with pytest.raises(hb.HerosError) as err:          # loud, once: the window ran without the 9 rejected stays
        eng.step(24.0)
    assert "9 submitted stays were rejected" in str(err.value)
    assert eng.time == 24.0
    counts = [eng.phase_counts()]
    for d in range(1, 4):
        eng.step(24.0 * (d + 1))
        counts.append(eng.phase_counts())
    gpu = dict(infections=eng.infections(), transitions=eng.transitions(), counts=np.array(counts), state=eng.agent_state())
    assert_same_run(gpu, ora)
    assert len(ora["infections"]["person"]) > 5
    eng.submit_visits(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int16), np.zeros(0), np.zeros(0))
    eng.step(24.0 * 5)                                   # an empty submission is fine
    eng.close()
