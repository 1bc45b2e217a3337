"""Sharded multi-GPU runs: one engine per GPU, geography-partitioned persons and locations.

Per window every shard
  1. runs the disease-phase state machine of its own persons                     (heros_window_begin),
  2. sends the stays its persons spend in OTHER shards' locations, with the packed agent word of the person, to
     the owner of the location                                                    (all-to-all #1, 40 B per stay),
  3. evaluates every sub-location it owns, guests included                        (heros_window_transmit),
  4. sends the successful draws on guests back to the shard that owns the person  (all-to-all #2, 26 B per candidate),
  5. resolves earliest-draw-wins over local and returned candidates               (heros_window_finish).
The only collectives on the data path are these two all-to-alls (NCCL over NVLink when the tensors live on GPUs);
nothing else is exchanged, and because draws are keyed by global person id and exposure sums are formed in canonical
order the sharded run is bit-identical to the same population on a single engine.

The routing helpers are plain torch and run on CPU tensors too (gloo), which is how tests/test_sharded_host.py
exercises them without a GPU.  On a GPU every torch operation of a shard (routing, packing, the NCCL collectives) is
issued on the engine's own CUDA stream (heros_get_stream), so it is ordered with the engine's kernels without any
extra synchronisation.
"""
from __future__ import annotations

import contextlib
from typing import Dict, List, Optional, Sequence

import torch

GUEST_COLUMNS = (("person", torch.int32), ("word", torch.int64), ("ill_end", torch.float64), ("loc", torch.int32),
                 ("sub", torch.int16), ("t_in", torch.float64), ("t_out", torch.float64))
CANDIDATE_COLUMNS = (("person", torch.int32), ("loc", torch.int32), ("sub", torch.int16), ("time", torch.float64),
                     ("p", torch.float64))


def owner_of(ids: torch.Tensor, bases: torch.Tensor) -> torch.Tensor:
    """Rank that owns each global id; ``bases`` = first id of every rank followed by the total (ascending)."""
    return torch.bucketize(ids.to(torch.int64), bases[1:].to(ids.device), right=True)


def route(columns: Dict[str, torch.Tensor], dest: torch.Tensor, world: int):
    """Stable grouping of the rows by destination rank: (columns in destination order, rows per destination)."""
    order = torch.sort(dest, stable=True).indices
    counts = torch.bincount(dest, minlength=world)[:world]
    return {k: v[order].contiguous() for k, v in columns.items()}, counts


def split_by_counts(columns: Dict[str, torch.Tensor], counts: Sequence[int]) -> List[Dict[str, torch.Tensor]]:
    out, start = [], 0
    for c in counts:
        out.append({k: v[start:start + int(c)] for k, v in columns.items()})
        start += int(c)
    return out


def concat_columns(parts: List[Dict[str, torch.Tensor]], schema, device) -> Dict[str, torch.Tensor]:
    return {k: (torch.cat([p[k] for p in parts]) if parts else torch.empty(0, dtype=dt, device=device))
            for k, dt in schema}


def all_to_all_rows(columns: Dict[str, torch.Tensor], send_counts: torch.Tensor, group=None):
    """Variable-size all-to-all of row-aligned columns (torch.distributed: NCCL on GPU tensors, gloo on CPU tensors).
    Two collectives per exchange: the row counts, then ONE payload in which the rows for a destination travel as a
    block of column slices (widest columns first, the block padded to 8 bytes, so every slice is aligned for its
    dtype on arrival).  Returns (received columns in source-rank order, rows received per source rank)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    device = next(iter(columns.values())).device
    send = send_counts.to(device=device, dtype=torch.int64)
    recv = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(recv, send, group=group)
    send_l, recv_l = send.tolist(), recv.tolist()
    names = sorted(columns, key=lambda k: -columns[k].element_size())
    widths = [columns[k].element_size() for k in names]
    row_bytes = sum(widths)

    def block(rows):
        return (rows * row_bytes + 7) // 8 * 8

    pieces, start = [], 0
    for d in range(world):
        c = send_l[d]
        used = 0
        for k, w in zip(names, widths):
            pieces.append(columns[k][start:start + c].contiguous().view(torch.uint8))
            used += c * w
        if block(c) > used:
            pieces.append(torch.zeros(block(c) - used, dtype=torch.uint8, device=device))
        start += c
    sbuf = torch.cat(pieces) if pieces else torch.empty(0, dtype=torch.uint8, device=device)
    rbuf = torch.empty(sum(block(r) for r in recv_l), dtype=torch.uint8, device=device)
    dist.all_to_all_single(rbuf, sbuf, output_split_sizes=[block(r) for r in recv_l],
                           input_split_sizes=[block(c) for c in send_l], group=group)
    parts = {k: [] for k in names}
    pos = 0
    for r in recv_l:
        off = pos
        for k, w in zip(names, widths):
            parts[k].append(rbuf[off:off + r * w].view(columns[k].dtype))
            off += r * w
        pos += block(r)
    return {k: torch.cat(parts[k]) for k in columns}, recv_l


class ShardedEngine:
    """One shard: a HerosEngine plus the routing tables.  ``person_bases`` / ``location_bases``: first global id of every
    rank followed by the total."""

    def __init__(self, engine, rank: int, world: int, person_bases, location_bases, device="cuda"):
        self.engine, self.rank, self.world, self.device = engine, rank, world, torch.device(device)
        self.person_bases = torch.as_tensor(person_bases, dtype=torch.int64)
        self.location_bases = torch.as_tensor(location_bases, dtype=torch.int64)
        engine.set_id_bases(int(self.person_bases[rank]), int(self.location_bases[rank]))
        self.stream = torch.cuda.ExternalStream(engine.stream, device=self.device) if self.device.type == "cuda" else None
        self.sent_stays = 0
        self.sent_candidates = 0

    def on_stream(self):
        """Context in which torch work is issued on the engine's CUDA stream."""
        return torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()

    # ---- phase 1: state machine, then the stays that leave the shard ----
    def begin_and_route(self, t_until: float, remote: Optional[Dict[str, torch.Tensor]]):
        """remote: this shard's persons' stays in other shards' locations that are active in the window (device tensors
        person i32, loc i32, sub i16, t_in f64, t_out f64, global ids).  Returns (columns by destination, counts)."""
        eng = self.engine
        eng.window_begin(t_until)
        if remote is not None and remote["person"].device != self.device:  # pinned host buffers: H2D on the engine stream
            remote = {k: v.to(self.device, non_blocking=True) for k, v in remote.items()}
        n = 0 if remote is None else int(remote["person"].numel())
        word = torch.empty(n, dtype=torch.int64, device=self.device)
        ill_end = torch.empty(n, dtype=torch.float64, device=self.device)
        if n:
            eng.export_agent_words(n, remote["person"].data_ptr(), word.data_ptr(), ill_end.data_ptr())
            cols = {"person": remote["person"], "word": word, "ill_end": ill_end, "loc": remote["loc"],
                    "sub": remote["sub"], "t_in": remote["t_in"], "t_out": remote["t_out"]}
            dest = owner_of(remote["loc"], self.location_bases)
        else:
            cols = concat_columns([], GUEST_COLUMNS, self.device)
            dest = torch.empty(0, dtype=torch.int64, device=self.device)
        self.sent_stays += n
        return route(cols, dest, self.world)

    # ---- phase 2: guests in, every evaluation, candidates out ----
    def transmit_and_route(self, guests: Dict[str, torch.Tensor]):
        eng = self.engine
        n = int(guests["person"].numel())
        if n:
            g = {k: guests[k].contiguous() for k, _ in GUEST_COLUMNS}
            eng.submit_guest_visits(n, *(g[k].data_ptr() for k, _ in GUEST_COLUMNS))
        eng.window_transmit()
        cap = max(4096, 4 * n)
        while True:
            out = {k: torch.empty(cap, dtype=dt, device=self.device) for k, dt in CANDIDATE_COLUMNS}
            try:
                m = eng.export_guest_candidates(cap, *(out[k].data_ptr() for k, _ in CANDIDATE_COLUMNS))
                break
            except Exception as e:  # HEROS_ERR_CAPACITY: retry with the count the engine reported
                if getattr(e, "code", None) != -4:
                    raise
                cap *= 4
        cols = {k: v[:m] for k, v in out.items()}
        dest = owner_of(cols["person"], self.person_bases)
        self.sent_candidates += m
        return route(cols, dest, self.world)

    # ---- phase 3: candidates in, resolve ----
    def finish(self, candidates: Dict[str, torch.Tensor]):
        n = int(candidates["person"].numel())
        if n:
            c = {k: candidates[k].contiguous() for k, _ in CANDIDATE_COLUMNS}
            self.engine.import_candidates(n, *(c[k].data_ptr() for k, _ in CANDIDATE_COLUMNS))
        return self.engine.window_finish()

    # ---- one window over torch.distributed ----
    def step_window(self, t_until: float, remote: Optional[Dict[str, torch.Tensor]], group=None):
        with self.on_stream():
            cols, counts = self.begin_and_route(t_until, remote)
            guests, _ = all_to_all_rows(cols, counts, group)
            cols, counts = self.transmit_and_route(guests)
            cands, _ = all_to_all_rows(cols, counts, group)
            return self.finish(cands)


def step_window_in_process(shards: List[ShardedEngine], t_until: float, remotes: List[Optional[Dict[str, torch.Tensor]]]):
    """The same window for several shards living in ONE process (e.g. all on one GPU): the all-to-alls become list
    shuffles.  Used to check that a sharded run equals the unsharded one without needing several GPUs."""
    world = len(shards)
    dev = shards[0].device

    def run(fn, args):
        out = []
        for s, a in zip(shards, args):
            with s.on_stream():
                out.append(fn(s, a))
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)  # the shards' streams meet here, as they would in the collective
        return out

    out1 = run(lambda s, r: s.begin_and_route(t_until, r), remotes)
    parts = [split_by_counts(c, n.tolist()) for c, n in out1]
    inbox = [concat_columns([parts[src][dst] for src in range(world)], GUEST_COLUMNS, dev) for dst in range(world)]
    out2 = run(lambda s, g: s.transmit_and_route(g), inbox)
    parts = [split_by_counts(c, n.tolist()) for c, n in out2]
    inbox = [concat_columns([parts[src][dst] for src in range(world)], CANDIDATE_COLUMNS, dev) for dst in range(world)]
    return run(lambda s, c: s.finish(c), inbox)


# ---------------------------------------------------------------------------------------------------------------------
# the same window with fixed-capacity blocks: two equal-split all-to-alls, no host synchronisation inside the window
# ---------------------------------------------------------------------------------------------------------------------
def sort_by_destination(columns: Dict[str, "np.ndarray"], location_bases, world: int):
    """Host-side (numpy) preparation of a shard's remote stays: rows grouped by the rank that owns the location, plus the
    world + 1 row offsets heros_pack_guest_blocks_device takes.  Done by whoever produces the stays (the schedule)."""
    import numpy as np
    bases = np.asarray(location_bases, np.int64)
    dest = np.searchsorted(bases[1:], columns["loc"].astype(np.int64), side="right")
    order = np.argsort(dest, kind="stable")
    offsets = np.concatenate([[0], np.cumsum(np.bincount(dest, minlength=world)[:world])]).astype(np.int64)
    return {k: v[order] for k, v in columns.items()}, offsets


class BlockExchange:
    """Per-shard buffers and the window driver of the block exchange (heros_pack_guest_blocks_device and friends).
    ``guest_capacity`` / ``candidate_capacity``: rows per peer block, the same on every rank."""

    VISIT_COLUMNS = ("person", "loc", "sub", "t_in", "t_out")

    def __init__(self, shard: ShardedEngine, guest_capacity: int, candidate_capacity: int = 16384):
        self.shard = shard
        eng, world, dev = shard.engine, shard.world, shard.device
        self.gc = max(4, (int(guest_capacity) + 3) // 4 * 4)
        self.cc = max(4, (int(candidate_capacity) + 3) // 4 * 4)
        self.g_bytes, self.c_bytes = eng.guest_block_bytes(self.gc), eng.candidate_block_bytes(self.cc)
        self.g_send = torch.zeros(world * self.g_bytes, dtype=torch.uint8, device=dev)
        self.g_recv = torch.zeros_like(self.g_send)
        self.c_send = torch.zeros(world * self.c_bytes, dtype=torch.uint8, device=dev)
        self.c_recv = torch.zeros_like(self.c_send)
        self.person_bases = shard.person_bases.to(torch.int32).numpy()

    # the three phases, separated by the two exchanges
    def begin_and_pack(self, t_until: float, remote: Optional[Dict[str, torch.Tensor]], row_offsets):
        sh, eng = self.shard, self.shard.engine
        eng.window_begin(t_until)
        if remote is not None and remote["person"].device != sh.device:  # pinned host columns: H2D on the engine stream
            remote = {k: v.to(sh.device, non_blocking=True) for k, v in remote.items()}
        self._keep = remote  # read asynchronously by the pack kernel
        ptrs = [0] * 5 if remote is None or remote["person"].numel() == 0 else [remote[k].data_ptr() for k in self.VISIT_COLUMNS]
        sh.sent_stays += int(row_offsets[-1])
        eng.pack_guest_blocks(row_offsets, *ptrs, self.gc, self.g_send.data_ptr())

    def guests_in_and_transmit(self):
        sh, eng = self.shard, self.shard.engine
        eng.submit_guest_blocks(sh.world, self.gc, self.g_recv.data_ptr())
        eng.window_transmit()
        eng.export_candidate_blocks(self.person_bases, self.cc, self.c_send.data_ptr())

    def candidates_in_and_finish(self):
        sh, eng = self.shard, self.shard.engine
        eng.import_candidate_blocks(sh.world, self.cc, self.c_recv.data_ptr())
        return eng.window_finish()

    def step_window(self, t_until: float, remote: Optional[Dict[str, torch.Tensor]], row_offsets, group=None):
        """One window over torch.distributed; everything is enqueued on the engine stream, the only host
        synchronisation is the one at the end of heros_window_finish."""
        import torch.distributed as dist
        with self.shard.on_stream():
            self.begin_and_pack(t_until, remote, row_offsets)
            dist.all_to_all_single(self.g_recv, self.g_send, group=group)
            self.guests_in_and_transmit()
            dist.all_to_all_single(self.c_recv, self.c_send, group=group)
            return self.candidates_in_and_finish()


def step_window_blocks_in_process(exchanges: List[BlockExchange], t_until: float, remotes, offsets):
    """The block exchange for several shards living in ONE process: the all-to-alls become block transposes."""
    world = len(exchanges)
    dev = exchanges[0].shard.device

    def run(fn):
        out = []
        for x in exchanges:
            with x.shard.on_stream():
                out.append(fn(x))
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        return out

    def transpose(send_of, recv_of, nbytes):
        for dst in range(world):
            for src in range(world):
                recv_of(exchanges[dst])[src * nbytes:(src + 1) * nbytes] = send_of(exchanges[src])[dst * nbytes:(dst + 1) * nbytes]
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)

    for x, r, o in zip(exchanges, remotes, offsets):
        with x.shard.on_stream():
            x.begin_and_pack(t_until, r, o)
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    transpose(lambda x: x.g_send, lambda x: x.g_recv, exchanges[0].g_bytes)
    run(lambda x: x.guests_in_and_transmit())
    transpose(lambda x: x.c_send, lambda x: x.c_recv, exchanges[0].c_bytes)
    return run(lambda x: x.candidates_in_and_finish())


# ---------------------------------------------------------------------------------------------------------------------
# the block exchange with no collective at all: the kernels write into the peers' memory over NVLink (CUDA IPC)
# ---------------------------------------------------------------------------------------------------------------------
class PeerExchange:
    """Window driver of the peer-to-peer exchange (heros_exchange_create / heros_window_send_guests and friends): the
    shards of one node map each other's receive regions once (the 64-byte IPC handles are gathered with
    torch.distributed); afterwards a window needs no collective call and no host synchronisation before its end.
    ``local_peers``: the PeerExchange objects of shards living in THIS process (tests), connected by device pointer."""

    VISIT_COLUMNS = BlockExchange.VISIT_COLUMNS

    def __init__(self, shard: ShardedEngine, guest_capacity: int, candidate_capacity: int = 4096):
        self.shard = shard
        self.gc = max(4, (int(guest_capacity) + 3) // 4 * 4)
        self.cc = max(4, (int(candidate_capacity) + 3) // 4 * 4)
        self.handle, self.region = shard.engine.exchange_create(shard.rank, shard.world, self.gc, self.cc)
        self.person_bases = shard.person_bases.to(torch.int32).numpy()
        self._keep = None

    def connect(self, group=None, local_peers: Optional[List["PeerExchange"]] = None):
        import numpy as np
        if local_peers is not None:
            self.shard.engine.exchange_connect(None, [0 if p is self else p.region for p in local_peers])
            return
        import torch.distributed as dist
        mine = torch.from_numpy(self.handle.copy()).to(self.shard.device)
        all_handles = [torch.empty_like(mine) for _ in range(self.shard.world)]
        dist.all_gather(all_handles, mine, group=group)
        self.shard.engine.exchange_connect(np.stack([h.cpu().numpy() for h in all_handles]))

    def begin_and_send(self, t_until: float, remote: Optional[Dict[str, torch.Tensor]], row_offsets):
        sh, eng = self.shard, self.shard.engine
        eng.window_begin(t_until)
        if remote is not None and remote["person"].device != sh.device:  # pinned host columns: H2D on the engine stream
            remote = {k: v.to(sh.device, non_blocking=True) for k, v in remote.items()}
        self._keep = remote  # read asynchronously by the pack kernel
        ptrs = [0] * 5 if remote is None or remote["person"].numel() == 0 else [remote[k].data_ptr() for k in self.VISIT_COLUMNS]
        sh.sent_stays += int(row_offsets[-1])
        eng.window_send_guests(row_offsets, *ptrs)

    def receive_and_transmit(self):
        eng = self.shard.engine
        eng.window_recv_guests()
        eng.window_transmit()
        eng.window_send_candidates(self.person_bases)

    def receive_and_finish(self):
        eng = self.shard.engine
        eng.window_recv_candidates()
        return eng.window_finish()

    def set_sources(self, ranks):
        """Only these shards ever send guests to this one: only their flags are awaited (default: all shards)."""
        self.shard.engine.exchange_set_sources(ranks)

    def step_window(self, t_until: float, remote: Optional[Dict[str, torch.Tensor]], row_offsets, group=None, wait: bool = True):
        """One window, one call into the library: begin, send / receive guests, transmit, send / receive candidates,
        finish -- all enqueued on the engine stream, one host synchronisation at the end (``wait=False``: none; the
        windows are then collected with ``engine.wait()``)."""
        import numpy as np
        sh = self.shard
        with sh.on_stream():
            if remote is not None and remote["person"].device != sh.device:  # pinned host columns: H2D on the engine stream
                remote = {k: v.to(sh.device, non_blocking=True) for k, v in remote.items()}
            self._keep_ring = getattr(self, "_keep_ring", [])
            self._keep_ring.append(remote)  # read asynchronously by the pack kernel: kept for a few windows
            del self._keep_ring[:-4]
            ptrs = [0] * 5 if remote is None or remote["person"].numel() == 0 else [remote[k].data_ptr() for k in self.VISIT_COLUMNS]
            off = np.ascontiguousarray(row_offsets, np.int64)
            sh.sent_stays += int(off[-1])
            if wait:
                return sh.engine.window_run_peers(t_until, off, *ptrs, self.person_bases)
            return sh.engine.window_run_peers_async(t_until, off, *ptrs, self.person_bases)


def step_window_peers_in_process(exchanges: List[PeerExchange], t_until: float, remotes, offsets):
    """The peer-to-peer window for several shards living in ONE process on one GPU: every phase is issued for all shards
    before the next one starts, so no kernel ever waits for a flag that has not been raised yet."""
    dev = exchanges[0].shard.device
    for x, r, o in zip(exchanges, remotes, offsets):
        with x.shard.on_stream():
            x.begin_and_send(t_until, r, o)
    torch.cuda.synchronize(dev)
    for x in exchanges:
        with x.shard.on_stream():
            x.receive_and_transmit()
    torch.cuda.synchronize(dev)
    out = []
    for x in exchanges:
        with x.shard.on_stream():
            out.append(x.receive_and_finish())
    return out


# ---------------------------------------------------------------------------------------------------------------------
# geographic partition of ONE city: locations in Morton order of (x, y), cut into contiguous ranges of
# equal expected load; a person belongs to the shard of its home
# ---------------------------------------------------------------------------------------------------------------------
def morton_codes(x, y, bits: int = 16):
    """Z-order key of every point: x and y quantised to ``bits`` bits over their bounding box, bits interleaved."""
    import numpy as np

    def quantise(v):
        v = np.asarray(v, np.float64)
        lo, hi = float(v.min()), float(v.max())
        q = np.zeros(len(v), np.uint64) if hi <= lo else np.minimum(((v - lo) / (hi - lo) * (1 << bits)).astype(np.uint64),
                                                                    np.uint64((1 << bits) - 1))
        return q

    def spread(q):  # abcd -> 0a0b0c0d
        q = q & np.uint64(0xFFFFFFFF)
        for shift, mask in ((16, 0x0000FFFF0000FFFF), (8, 0x00FF00FF00FF00FF), (4, 0x0F0F0F0F0F0F0F0F),
                            (2, 0x3333333333333333), (1, 0x5555555555555555)):
            q = (q | (q << np.uint64(shift))) & np.uint64(mask)
        return q

    return spread(quantise(x)) | (spread(quantise(y)) << np.uint64(1))


class CityPartition:
    """Result of ``partition_city``.  The global numbering of a sharded run makes every shard's persons and locations
    contiguous: ``loc_new[l]`` / ``person_new[p]`` = new global id of location l / person p of the original city,
    ``loc_bases`` / ``person_bases`` = first new id of every shard followed by the total.  ``shard_city(r)`` is the
    city of shard r (own persons and locations in local numbering, foreign locations its persons can reach appended
    behind them with their global ids in ``ext_global``); ``global_city()`` the whole city in the new numbering -- the
    single-engine / oracle run a sharded run must equal."""

    def __init__(self, city, shard_of_location, n_shards: int):
        import numpy as np
        self.city, self.n_shards = city, int(n_shards)
        self.shard_of_location = np.asarray(shard_of_location, np.int32)
        self.shard_of_person = self.shard_of_location[city.home_loc]
        # new ids: stable within a shard (original order kept)
        l_order = np.argsort(self.shard_of_location, kind="stable")
        p_order = np.argsort(self.shard_of_person, kind="stable")
        self.loc_old, self.person_old = l_order, p_order            # new id -> original id
        self.loc_new = np.empty(len(l_order), np.int64); self.loc_new[l_order] = np.arange(len(l_order))
        self.person_new = np.empty(len(p_order), np.int64); self.person_new[p_order] = np.arange(len(p_order))
        self.loc_bases = np.concatenate([[0], np.cumsum(np.bincount(self.shard_of_location, minlength=n_shards))]).astype(np.int64)
        self.person_bases = np.concatenate([[0], np.cumsum(np.bincount(self.shard_of_person, minlength=n_shards))]).astype(np.int64)

    def _make(self, persons_old, locs_old_own, locs_old_ext, person_gid, loc_gid, ext_global):
        """A City over the given original persons / locations (own first, then foreign), references renumbered."""
        import numpy as np
        from .scenario import City, NONE
        c = self.city
        locs = np.concatenate([locs_old_own, locs_old_ext]).astype(np.int64)
        local_of = np.full(c.n_locations, NONE, np.int64)
        local_of[locs] = np.arange(len(locs))

        def remap(v):
            v = np.asarray(v, np.int64)
            return np.where(v >= 0, local_of[np.maximum(v, 0)], NONE)

        out = object.__new__(City)
        out.n_persons = len(persons_old)
        out.n_locations = len(locs_old_own)
        out.type_infect_sub, out.type_correction = c.type_infect_sub.copy(), c.type_correction.copy()
        out.n_random_candidates = c.n_random_candidates
        for name in ("loc_type", "loc_nsub", "loc_area", "loc_x", "loc_y"):
            setattr(out, name, getattr(c, name)[locs])
        for name in ("age", "female", "role", "home_sub", "work_sub", "household"):
            setattr(out, name, getattr(c, name)[persons_old])
        out.home_loc = remap(c.home_loc[persons_old]).astype(np.int32)
        out.work_loc = remap(c.work_loc[persons_old]).astype(np.int32)
        # locator tables: a row per location a person of the shard can be in (own and foreign), entries likewise
        out.nearest = remap(c.nearest[locs]).astype(np.int32)
        out.cand = remap(c.cand[locs]).astype(np.int32)
        out.n_cand = c.n_cand[locs].copy()
        if getattr(c, "divert", None) is not None:
            out.divert = c.divert[locs]
        out.person_gid = np.asarray(person_gid, np.int64)
        out.loc_gid = np.asarray(loc_gid, np.int64)
        if ext_global is not None:
            out.ext_global = np.asarray(ext_global, np.int64)
        return out

    def global_city(self):
        import numpy as np
        n_l, n_p = len(self.loc_old), len(self.person_old)
        return self._make(self.person_old, self.loc_old, np.empty(0, np.int64), np.arange(n_p), np.arange(n_l), None)

    def shard_city(self, r: int):
        import numpy as np
        c = self.city
        persons = self.person_old[self.person_bases[r]:self.person_bases[r + 1]]
        own = self.loc_old[self.loc_bases[r]:self.loc_bases[r + 1]]
        # Every other location of the city follows as a foreign one: a person searches the nearest / a random place of a
        # type from wherever it IS (CurrentLocator), e.g. from a workplace in another shard, so the rows of the locator
        # tables of all places it can get to are needed -- transitively, the whole neighbourhood.  (Tables only: the
        # shard's engine holds its own locations alone.)
        ext = self.loc_old[np.r_[0:self.loc_bases[r], self.loc_bases[r + 1]:len(self.loc_old)]]
        return self._make(persons, own, ext, self.person_new[persons], np.concatenate([self.loc_new[own], self.loc_new[ext]]),
                          self.loc_new[ext])


def expected_load(city):
    """Expected stays per day of every location, the weight the partition balances: residents are at home about three
    times a day, workers / pupils at their place twice, and every person visits the public places the locator tables of
    its home point to about once."""
    import numpy as np
    n = city.n_locations
    w = 3.0 * np.bincount(city.home_loc, minlength=n)[:n].astype(np.float64)
    has = city.work_loc >= 0
    w += 2.0 * np.bincount(city.work_loc[has], minlength=n)[:n]
    residents = np.bincount(city.home_loc, minlength=n)[:n].astype(np.float64)
    near = city.nearest
    n_types = max(1, (near >= 0).any(axis=0).sum())
    for t in range(near.shape[1]):
        col = near[:, t]
        ok = col >= 0
        if ok.any():
            w += np.bincount(col[ok], weights=residents[ok] / n_types, minlength=n)[:n]
    return w + 1e-3


def partition_city(city, n_shards: int, weights=None) -> CityPartition:
    """Cuts one city into ``n_shards`` geographic shards: the locations are ordered along the Morton (Z-order) curve of
    their coordinates and the curve is cut into contiguous pieces of equal expected load (``expected_load`` unless
    ``weights`` is given); a person is owned by the shard of its home (the district columns WK_CODE /
    BU_CODE of the reference's locations.csv would serve the same purpose, metadata/DataDictionary.xlsx sheet
    ``locations`` rows 12-13)."""
    import numpy as np
    if n_shards < 1:
        raise ValueError("n_shards must be positive")
    w = expected_load(city) if weights is None else np.asarray(weights, np.float64)
    code = morton_codes(city.loc_x, city.loc_y)
    order = np.argsort(code, kind="stable")
    cum = np.cumsum(w[order])
    shard_sorted = np.minimum((cum - 0.5 * w[order]) / cum[-1] * n_shards, n_shards - 1).astype(np.int32)
    shard = np.empty(city.n_locations, np.int32)
    shard[order] = shard_sorted
    return CityPartition(city, shard, n_shards)


VISIT_COLUMNS = ("person", "loc", "sub", "t_in", "t_out")


def split_stays(city, st, rank: int, person_bases, location_bases, open_remote=None, other_base=None):
    """The stays a shard's schedule generator produced for one day (local ids) -> what the shard hands to its engine:
    (own columns with global ids, remote columns grouped by destination shard, the world + 1 row offsets of the groups,
    the remote stays still open at the end of the day).  A stay in a location of another shard is a *remote* stay: it
    travels to the owner of the location every window it is active in; while it is open (t_out = +inf) it is kept in
    ``open_remote`` and sent again the next day, until its closed version arrives.  ``other_base``: see
    ``City.global_location_ids`` (tiles joined by ``add_commuters``)."""
    import numpy as np
    world = len(person_bases) - 1
    person = (st["person"].astype(np.int64) + int(person_bases[rank])).astype(np.int32)
    loc = city.global_location_ids(st["loc"], int(location_bases[rank]), other_base)
    is_remote = st["loc"] >= city.n_locations
    cols = {"person": person, "loc": loc, "sub": st["sub"], "t_in": st["t_in"], "t_out": st["t_out"]}
    own = {k: v[~is_remote] for k, v in cols.items()}
    rem = {k: v[is_remote] for k, v in cols.items()}
    if open_remote is not None and len(open_remote["person"]):
        # an open remote stay of an earlier day stays a guest until its closed version arrives
        key_now = (rem["person"].astype(np.int64) << 32) ^ rem["t_in"].view(np.int64)
        key_old = (open_remote["person"].astype(np.int64) << 32) ^ open_remote["t_in"].view(np.int64)
        keep = ~np.isin(key_old, key_now)
        rem = {k: np.concatenate([open_remote[k][keep], rem[k]]) for k in rem}
    still_open = np.isinf(rem["t_out"])
    open_next = {k: v[still_open] for k, v in rem.items()}
    rem, offsets = sort_by_destination(rem, location_bases, world)
    return own, rem, offsets, open_next
