"""One city cut into geographic shards (sharded.partition_city: Morton order, persons owned by their home shard), every
shard with its own engine and its own schedule generator, stays in other shards' locations exchanged peer to peer between
ALL shards -- against the sequential oracle on the whole city: infection events, transitions and daily counts identical."""
import numpy as np
import pytest
import torch

import heros_b200 as hb
from common import load_props, oracle_props, sort_events
from heros_b200 import _lib, properties, scenario, sharded
from oracle import binding as ob

pytestmark = pytest.mark.gpu


def keyed_seeds(n_persons, seed, n):
    x, y, _, _ = scenario.philox4x32_10(np.arange(n_persons), 0, 0, 0x200, seed, seed >> 32)
    h = (x << np.uint64(32)) | y
    return np.sort(np.lexsort((np.arange(n_persons), h))[:n]).astype(np.int32)


@pytest.mark.parametrize("model,world,exchange", [("distance", 3, "peers"), ("area", 4, "peers"), ("distance", 2, "blocks")])
def test_partitioned_city_equals_the_oracle_on_the_whole_city(model, world, exchange):
    days, seed = 8, 111
    city = scenario.City(9000, seed=3, work_anywhere_fraction=0.25)
    part = sharded.partition_city(city, world)
    whole = part.global_city()
    props = load_props(model)
    props["covidT_area.contagiousness" if model == "area" else "covidT_dist.alpha"] = "4.0" if model == "area" else "1.5"
    kind = _lib.MODEL_AREA if model == "area" else _lib.MODEL_DISTANCE
    seeds = keyed_seeds(whole.n_persons, seed, 60)

    # the oracle on the whole city, driven by the whole city's schedule generator
    area, dist, prog = oracle_props(props)
    sim = ob.Sim(ob.MODEL_AREA if model == "area" else ob.MODEL_DIST, ob.TRIGGER_EXIT_ONLY, seed)
    sim.set_area(area) if model == "area" else sim.set_dist(dist)
    sim.set_prog(prog)
    sim.set_location_types(whole.type_infect_sub, whole.type_correction)
    sim.set_locations(whole.loc_type, whole.loc_nsub, whole.loc_area)
    sim.set_persons(whole.age, whole.female)
    sim.expose(seeds, np.zeros(len(seeds)))
    gen = scenario.ScheduleGenerator(whole, seed)
    ora_counts = []
    for d in range(days):
        st = gen.generate_day(sim.agent_state()["phase"])
        sim.add_visits(st["person"], st["loc"], st["sub"], st["t_in"], st["t_out"])
        sim.run(24.0 * (d + 1))
        ora_counts.append(sim.phase_counts())

    dev = torch.device("cuda")
    shards, gens, cities = [], [], []
    for r in range(world):
        c = part.shard_city(r)
        eng = hb.HerosEngine(kind, seed, flags=_lib.FLAG_EXACT_STATS)
        c.install(eng)
        (eng.set_transmission_area(properties.area_params(props)) if model == "area"
         else eng.set_transmission_distance(properties.dist_params(props)))
        eng.set_progression(properties.prog_params(props))
        sh = sharded.ShardedEngine(eng, r, world, part.person_bases, part.loc_bases, device=dev)
        mine = seeds[(seeds >= part.person_bases[r]) & (seeds < part.person_bases[r + 1])]
        eng.expose(mine, np.zeros(len(mine)))
        shards.append(sh); gens.append(scenario.ScheduleGenerator(c, seed)); cities.append(c)
    if exchange == "peers":
        xs = [sharded.PeerExchange(sh, guest_capacity=8192, candidate_capacity=1024) for sh in shards]
        for x in xs:
            x.connect(local_peers=xs)   # every shard listens to every other one: no source mask
    else:
        xs = [sharded.BlockExchange(sh, guest_capacity=8192, candidate_capacity=1024) for sh in shards]
    open_remote = [None] * world
    counts, n_remote = [], 0
    for d in range(days):
        remotes, offsets = [], []
        for r, sh in enumerate(shards):
            st = gens[r].generate_day(sh.engine.agent_state()["phase"])
            own, rem, off, open_remote[r] = sharded.split_stays(cities[r], st, r, part.person_bases, part.loc_bases, open_remote[r])
            sh.engine.submit_visits(own["person"], own["loc"], own["sub"], own["t_in"], own["t_out"])
            n_remote += len(rem["person"])
            remotes.append({k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in rem.items()})
            offsets.append(off)
        if exchange == "peers":
            sharded.step_window_peers_in_process(xs, 24.0 * (d + 1), remotes, offsets)
        else:
            sharded.step_window_blocks_in_process(xs, 24.0 * (d + 1), remotes, offsets)
        counts.append(sum(sh.engine.phase_counts() for sh in shards))
    assert n_remote > 2000

    inf = [sh.engine.infections() for sh in shards]
    tr = [sh.engine.transitions() for sh in shards]
    gi = sort_events({k: np.concatenate([x[k] for x in inf]) for k in inf[0]})
    gt = sort_events({k: np.concatenate([x[k] for x in tr]) for k in tr[0]})
    oi, ot = sort_events(sim.infections()), sort_events(sim.transitions())
    assert len(oi["person"]) > 150
    for k in ("person", "loc", "sub", "time"):
        np.testing.assert_array_equal(gi[k], oi[k])
    np.testing.assert_allclose(gi["p"], oi["p"], rtol=1e-12, atol=4.5e-16)
    for k in ("person", "phase", "time"):
        np.testing.assert_array_equal(gt[k], ot[k])
    np.testing.assert_array_equal(np.array(counts), np.array(ora_counts))
    for sh in shards:
        sh.engine.close()
