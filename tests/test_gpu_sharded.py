"""Sharding the population over several engines must not change a single bit of the result: two / three shards living
on one GPU (the all-to-alls replaced by in-process list shuffles) against the same city on one engine and against the
sequential oracle."""
import numpy as np
import pytest
import torch

import heros_b200 as hb
from common import Scenario, load_props, run_oracle, sort_events
from heros_b200 import _lib, properties, sharded

pytestmark = pytest.mark.gpu


def build_engine(scn, props, model, trigger, seed, p0, p1, l0, l1):
    eng = hb.HerosEngine(model, seed, trigger=trigger, flags=_lib.FLAG_EXACT_STATS)
    eng.set_location_types(scn.type_infect_sub, scn.type_corr)
    eng.set_locations(scn.loc_type[l0:l1], scn.loc_nsub[l0:l1], scn.loc_area[l0:l1])
    eng.set_persons(scn.age[p0:p1], scn.female[p0:p1])
    if model == _lib.MODEL_AREA:
        eng.set_transmission_area(properties.area_params(props))
    else:
        eng.set_transmission_distance(properties.dist_params(props))
    eng.set_progression(properties.prog_params(props))
    return eng


@pytest.mark.parametrize("model,trigger,world,exchange", [("area", _lib.TRIGGER_ENTER_AND_EXIT, 2, "rows"),
                                                          ("distance", _lib.TRIGGER_EXIT_ONLY, 3, "rows"),
                                                          ("area", _lib.TRIGGER_EXIT_ONLY, 3, "blocks"),
                                                          ("distance", _lib.TRIGGER_ENTER_AND_EXIT, 2, "blocks"),
                                                          ("distance", _lib.TRIGGER_EXIT_ONLY, 3, "peers"),
                                                          ("area", _lib.TRIGGER_ENTER_AND_EXIT, 2, "peers")])
def test_sharded_equals_unsharded_and_oracle(model, trigger, world, exchange):
    scn = Scenario(seed=21, n_persons=900, n_households=330, n_workplaces=30, n_shops=4, days=9,
                   shop_visits_per_day=2.0, tie_fraction=0.2, short_gap_fraction=0.2)
    props = load_props(model)
    props["covidT_area.contagiousness" if model == "area" else "covidT_dist.alpha"] = "4.0" if model == "area" else "1.5"
    kind = _lib.MODEL_AREA if model == "area" else _lib.MODEL_DISTANCE
    ora = run_oracle(scn, props, kind, trigger, seed=77)

    pb = np.linspace(0, scn.n_persons, world + 1).astype(np.int64)
    lb = np.linspace(0, scn.n_loc, world + 1).astype(np.int64)
    dev = torch.device("cuda")
    shards = []
    for r in range(world):
        eng = build_engine(scn, props, kind, trigger, 77, pb[r], pb[r + 1], lb[r], lb[r + 1])
        shards.append(sharded.ShardedEngine(eng, r, world, pb, lb, device=dev))
    p_owner = np.searchsorted(pb[1:], scn.person, side="right")
    l_owner = np.searchsorted(lb[1:], scn.loc, side="right")
    for r, sh in enumerate(shards):
        seeds = scn.seeds[(scn.seeds >= pb[r]) & (scn.seeds < pb[r + 1])]
        sh.engine.expose(seeds, np.zeros(len(seeds)))                       # global ids in
        own = (p_owner == r) & (l_owner == r)
        sh.engine.submit_visits(scn.person[own], scn.loc[own], scn.sub[own], scn.t_in[own], scn.t_out[own])
    counts = []
    n_remote = 0
    if exchange == "peers":  # kernels write straight into the other shards' regions (same process: device pointers)
        exchanges = [sharded.PeerExchange(sh, guest_capacity=2048, candidate_capacity=512) for sh in shards]
        for x in exchanges:
            x.connect(local_peers=exchanges)
    else:
        exchanges = [sharded.BlockExchange(sh, guest_capacity=2048, candidate_capacity=512) for sh in shards]
    for d in range(scn.days):
        t0, t1 = 24.0 * d, 24.0 * (d + 1)
        remotes, offsets = [], []
        for r in range(world):
            m = (p_owner == r) & (l_owner != r) & (scn.t_in < t1) & (scn.t_out >= t0)   # active in the window
            n_remote += int(m.sum())
            cols = {"person": scn.person[m], "loc": scn.loc[m], "sub": scn.sub[m], "t_in": scn.t_in[m], "t_out": scn.t_out[m]}
            if exchange != "rows":  # the producer of the stays groups them by destination shard
                cols, off = sharded.sort_by_destination(cols, lb, world)
                offsets.append(off)
            remotes.append({k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in cols.items()})
        if exchange == "blocks":
            sharded.step_window_blocks_in_process(exchanges, t1, remotes, offsets)
        elif exchange == "peers":
            sharded.step_window_peers_in_process(exchanges, t1, remotes, offsets)
        else:
            sharded.step_window_in_process(shards, t1, remotes)
        counts.append(sum(sh.engine.phase_counts() for sh in shards))
    assert n_remote > 1000 and (exchange != "rows" or sum(sh.sent_candidates for sh in shards) > 10)

    inf = [sh.engine.infections() for sh in shards]
    tr = [sh.engine.transitions() for sh in shards]
    gi = sort_events({k: np.concatenate([x[k] for x in inf]) for k in inf[0]})
    gt = sort_events({k: np.concatenate([x[k] for x in tr]) for k in tr[0]})
    oi, ot = sort_events(ora["infections"]), sort_events(ora["transitions"])
    assert len(gi["person"]) > 100
    for k in ("person", "loc", "sub", "time"):
        np.testing.assert_array_equal(gi[k], oi[k])
    np.testing.assert_allclose(gi["p"], oi["p"], rtol=1e-12, atol=4.5e-16)
    for k in ("person", "phase", "time"):
        np.testing.assert_array_equal(gt[k], ot[k])
    np.testing.assert_array_equal(np.array(counts), ora["counts"])

    # ... and bit-identical, probabilities included, to the unsharded engine
    one = build_engine(scn, props, kind, trigger, 77, 0, scn.n_persons, 0, scn.n_loc)
    one.expose(scn.seeds, np.zeros(len(scn.seeds)))
    one.submit_visits(scn.person, scn.loc, scn.sub, scn.t_in, scn.t_out)
    for d in range(scn.days):
        one.step(24.0 * (d + 1))
    ui = sort_events(one.infections())
    for k in ("person", "loc", "sub", "time", "p"):
        np.testing.assert_array_equal(gi[k], ui[k])
    state = one.agent_state()
    for r, sh in enumerate(shards):
        st = sh.engine.agent_state()
        for k in ("phase", "exposure", "next_phase"):
            np.testing.assert_array_equal(st[k], state[k][pb[r]:pb[r + 1]])
    for sh in shards:
        sh.engine.close()
    one.close()


def test_submission_order_does_not_change_probabilities():
    """Exposure sums are formed in canonical (person, t_in) order: shuffling the submitted stays changes nothing."""
    scn = Scenario(seed=22, n_persons=700, n_households=250, n_workplaces=20, n_shops=3, days=7, shop_visits_per_day=2.0)
    props = load_props("distance")
    props["covidT_dist.alpha"] = "1.5"
    results = []
    for perm_seed in (0, 1):
        order = np.random.default_rng(perm_seed).permutation(scn.n_visits)
        eng = build_engine(scn, props, _lib.MODEL_DISTANCE, _lib.TRIGGER_EXIT_ONLY, 5, 0, scn.n_persons, 0, scn.n_loc)
        eng.expose(scn.seeds, np.zeros(len(scn.seeds)))
        eng.submit_visits(scn.person[order], scn.loc[order], scn.sub[order], scn.t_in[order], scn.t_out[order])
        for d in range(scn.days):
            eng.step(24.0 * (d + 1))
        results.append(sort_events(eng.infections()))
        eng.close()
    assert len(results[0]["person"]) > 50
    for k in ("person", "loc", "sub", "time", "p"):
        np.testing.assert_array_equal(results[0][k], results[1][k])
