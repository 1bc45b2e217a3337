"""Host-side logic of the sharded (multi-GPU) path on CPU: ownership lookup, routing, and the variable-size
all-to-all over torch.distributed with the gloo backend, world_size 2."""
import os
import socket

import torch
import torch.multiprocessing as mp

from heros_b200 import sharded


def test_owner_of_and_route():
    bases = torch.tensor([0, 10, 25, 40])  # three ranks: [0,10) [10,25) [25,40)
    ids = torch.tensor([0, 9, 10, 24, 25, 39, 12], dtype=torch.int32)
    assert sharded.owner_of(ids, bases).tolist() == [0, 0, 1, 1, 2, 2, 1]
    cols = {"person": ids, "x": torch.arange(7, dtype=torch.float64)}
    routed, counts = sharded.route(cols, sharded.owner_of(ids, bases), 3)
    assert counts.tolist() == [2, 3, 2]
    assert routed["person"].tolist() == [0, 9, 10, 24, 12, 25, 39]      # stable inside a destination
    assert routed["x"].tolist() == [0.0, 1.0, 2.0, 3.0, 6.0, 4.0, 5.0]
    parts = sharded.split_by_counts(routed, counts.tolist())
    assert [p["person"].tolist() for p in parts] == [[0, 9], [10, 24, 12], [25, 39]]
    empty = sharded.concat_columns([], sharded.CANDIDATE_COLUMNS, "cpu")
    assert all(v.numel() == 0 for v in empty.values()) and empty["sub"].dtype == torch.int16


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank owns persons [100 r, 100 r + 100) and sends rows about persons of both ranks
        g = torch.Generator().manual_seed(1234 + rank)
        n = 37 + 11 * rank
        person = torch.randint(0, 100 * world, (n,), generator=g, dtype=torch.int32)
        cols = {"person": person, "time": person.to(torch.float64) * 0.5 + rank,
                "sub": (person % 7).to(torch.int16)}
        bases = torch.arange(0, 100 * (world + 1), 100)
        routed, counts = sharded.route(cols, sharded.owner_of(person, bases), world)
        got, recv = sharded.all_to_all_rows(routed, counts)
        # everything received belongs to me, and sender r's rows arrive in r's order
        assert (sharded.owner_of(got["person"], bases) == rank).all()
        assert sum(recv) == got["person"].numel()
        start = 0
        for src in range(world):
            gs = torch.Generator().manual_seed(1234 + src)
            p_src = torch.randint(0, 100 * world, (37 + 11 * src,), generator=gs, dtype=torch.int32)
            mine = p_src[sharded.owner_of(p_src, bases) == rank]
            assert got["person"][start:start + recv[src]].tolist() == mine.tolist()
            assert torch.equal(got["time"][start:start + recv[src]], mine.to(torch.float64) * 0.5 + src)
            start += recv[src]
        assert got["sub"].dtype == torch.int16
        # an empty exchange must work too
        none = {k: v[:0] for k, v in cols.items()}
        got, recv = sharded.all_to_all_rows(none, torch.zeros(world, dtype=torch.int64))
        assert got["person"].numel() == 0 and recv == [0] * world
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_all_to_all_rows_gloo_world_size_2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as m:
        out = m.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert dict(out) == {0: 1, 1: 1}


def test_sort_by_destination_groups_rows_by_owner():
    import numpy as np
    bases = [0, 10, 25, 40]
    cols = {"person": np.arange(7, dtype=np.int32), "loc": np.array([30, 3, 12, 39, 0, 24, 10], np.int32),
            "sub": np.zeros(7, np.int16), "t_in": np.arange(7.0), "t_out": np.arange(7.0) + 1}
    out, off = sharded.sort_by_destination(cols, bases, 3)
    assert off.tolist() == [0, 2, 5, 7]
    assert out["loc"].tolist() == [3, 0, 12, 24, 10, 30, 39]          # stable inside a destination
    assert out["person"].tolist() == [1, 4, 2, 5, 6, 0, 3]
    empty, off = sharded.sort_by_destination({k: v[:0] for k, v in cols.items()}, bases, 3)
    assert off.tolist() == [0, 0, 0, 0] and len(empty["loc"]) == 0


def _partition_worker(rank, world, port, out):
    """One process per shard of ONE partitioned city: every rank builds the city, keeps its shard, generates its persons'
    day, splits it (own / remote by destination) and sends the remote stays to their owners over gloo; what a rank holds
    afterwards must be exactly the stays the WHOLE city's generator puts into the rank's locations."""
    import numpy as np
    import torch.distributed as dist
    from heros_b200 import scenario
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        whole0 = scenario.City(8000, seed=12, work_anywhere_fraction=0.3)
        part = sharded.partition_city(whole0, world)
        city = part.shard_city(rank)
        gen = scenario.ScheduleGenerator(city, 111)
        ref_gen = scenario.ScheduleGenerator(part.global_city(), 111)
        open_remote = None
        for d in range(3):
            phase_all = np.zeros(whole0.n_persons, np.uint8)
            if d == 1:
                phase_all[::11] = 3
            st = gen.generate_day(phase_all[part.person_bases[rank]:part.person_bases[rank + 1]])
            own, rem, off, open_remote = sharded.split_stays(city, st, rank, part.person_bases, part.loc_bases, open_remote)
            cols = {k: torch.from_numpy(np.ascontiguousarray(rem[k])) for k in sharded.VISIT_COLUMNS}
            counts = torch.from_numpy(np.diff(off).astype(np.int64))
            got, recv = sharded.all_to_all_rows(cols, counts)
            assert recv[rank] == 0
            mine = {k: np.concatenate([own[k], got[k].numpy()]) for k in sharded.VISIT_COLUMNS}
            ref = ref_gen.generate_day(phase_all)
            here = (ref["loc"] >= part.loc_bases[rank]) & (ref["loc"] < part.loc_bases[rank + 1])
            # a guest stay that is still open from an earlier day is sent again every day: leave those copies out
            fresh = ~((mine["t_in"] < 24.0 * d) & np.isinf(mine["t_out"]) &
                      ~((mine["person"] >= part.person_bases[rank]) & (mine["person"] < part.person_bases[rank + 1])))
            a = np.c_[mine["person"], mine["loc"], mine["sub"], mine["t_in"], mine["t_out"]][fresh]
            b = np.c_[ref["person"], ref["loc"], ref["sub"], ref["t_in"], ref["t_out"]][here]
            a, b = a[np.lexsort(a.T[::-1])], b[np.lexsort(b.T[::-1])]
            assert a.shape == b.shape and (a == b).all(), (rank, d, a.shape, b.shape)
            assert sum(recv) > 50
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_partitioned_city_exchange_gloo_world_size_2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as m:
        out = m.dict()
        procs = [ctx.Process(target=_partition_worker, args=(r, world, port, out)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(300)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert dict(out) == {0: 1, 1: 1}
