"""The one exchange step of the path: after DisparityToCloud every rank holds the points of its own
camera pair(s); the sink (CloudOptimization) wants all of them, in pair order.  Pairs shard one per rank with no
other data-path communication, so this is a count all-gather followed by one all-gather of the payload padded to the
largest count.  Works on whatever device the tensors live on: NCCL over NVLink for CUDA tensors, gloo for the CPU tests.

Since round 2 the data path does NOT go through this module: the exchange lives behind the C ABI (csrc/comm.cu:
sb200_comm_* / sb200_exchange_*; bench.py and the C++ CLI call that).  This file stays as the CPU-testable statement of the
same ticket / staging-slot logic (tests/test_exchange_gloo.py, world size 2 over gloo).
"""
from __future__ import annotations

import threading

import torch
import torch.distributed as dist

POINT_BYTES = 24 + 3 + 4  # xyz f64[3] + bgr u8[3] + source pixel index i32


def pair_to_rank(pair: int, world: int) -> int:
    """pair p -> rank p mod G (the C++ host mirror uses the same map for its per-device workers)."""
    return pair % world


def pairs_of_rank(rank: int, world: int, n_pairs: int):
    return list(range(rank, n_pairs, world))


def allgather_counts(n_local: int, device, group=None) -> torch.Tensor:
    world = dist.get_world_size(group)
    cnt = torch.tensor([n_local], dtype=torch.int64, device=device)
    cnts = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(cnts, cnt, group=group)
    return cnts


def allgather_points(xyz: torch.Tensor, bgr: torch.Tensor, pix: torch.Tensor, n_local: int, group=None, out=None):
    """xyz [cap,3] f64, bgr [cap,3] u8, pix [cap] i32 with the first n_local rows valid (cap >= the largest count of any
    rank).  Returns (counts [world] on the host, xyz_all, bgr_all, pix_all) where *_all are [world, nmax, ...] buffers;
    rank r's points are [r, :counts[r]].  `out` may carry reusable gather buffers (dict)."""
    world = dist.get_world_size(group)
    cnts = allgather_counts(n_local, xyz.device, group).cpu()
    nmax = max(int(cnts.max()), 1)
    if xyz.shape[0] < nmax:
        raise ValueError(f"point buffers hold {xyz.shape[0]} rows but another rank produced {nmax}")
    res = []
    for name, t in (("xyz", xyz), ("bgr", bgr), ("pix", pix)):
        src = t[:nmax].contiguous()
        shape = (world,) + tuple(src.shape)
        buf = None if out is None else out.get(name)
        if buf is None or tuple(buf.shape) != shape:
            buf = torch.empty(shape, dtype=src.dtype, device=src.device)
            if out is not None:
                out[name] = buf
        dist.all_gather_into_tensor(buf.view(-1), src.view(-1), group=group)
        res.append(buf)
    return cnts, res[0], res[1], res[2]


class PointExchanger:
    """The exchange, overlapped with the next pair's matching (SURVEY.md 8e "Overlap"): `submit` snapshots the local points
    into a staging buffer on the current stream and starts the all-gathers asynchronously; the matcher may then overwrite
    its point buffers while NCCL is still sending.  `finish` waits for the gathers of the last submit (the next `submit`
    does so implicitly before it reuses the staging buffers) and returns the gathered buffers."""

    def __init__(self, group=None):
        self.group = group
        self.stage = {}
        self.out = {}
        self.pending = []
        self.cnts = None

    def submit(self, xyz, bgr, pix, n_local):
        self.finish()
        world = dist.get_world_size(self.group)
        self.cnts = allgather_counts(n_local, xyz.device, self.group).cpu()
        nmax = max(int(self.cnts.max()), 1)
        if xyz.shape[0] < nmax:
            raise ValueError(f"point buffers hold {xyz.shape[0]} rows but another rank produced {nmax}")
        for name, t in (("xyz", xyz), ("bgr", bgr), ("pix", pix)):
            st = self.stage.get(name)
            if st is None or st.shape[0] < nmax:
                st = torch.empty((nmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
                self.stage[name] = st
            src = st[:nmax]
            src.copy_(t[:nmax], non_blocking=True)
            shape = (world,) + tuple(src.shape)
            buf = self.out.get(name)
            if buf is None or tuple(buf.shape) != shape:
                buf = torch.empty(shape, dtype=src.dtype, device=src.device)
                self.out[name] = buf
            self.pending.append(dist.all_gather_into_tensor(buf.view(-1), src.reshape(-1), group=self.group, async_op=True))
        return nmax * world * POINT_BYTES

    def finish(self):
        for w in self.pending:
            w.wait()
        self.pending = []
        if self.cnts is None:
            return None
        return self.cnts, self.out["xyz"], self.out["bgr"], self.out["pix"]


class OrderedPointExchange:
    """The exchange when a rank has several camera pairs in flight (one producer thread per context).

    Collectives must be issued in the same order on every rank, and a producer must never wait for another rank.  So the
    producers only SNAPSHOT: `submit(k, seq, ...)` copies the pair's points into one of this producer's staging slots on the
    producer's own stream and returns; one exchange thread per rank issues the gathers in ticket order
    (ticket = seq * n_producers + k, identical on every rank): it makes its stream wait for the snapshot, all-gathers the
    counts (the only host-blocking step, and it blocks this thread alone), then all-gathers the payload padded to the largest
    count.  A staging slot is reused once the gather that read it has been enqueued (host handshake) and has completed
    (stream-level event wait), so a producer runs up to `slots` pairs ahead of the exchange.  `on_gathered(ticket, counts,
    xyz_all, bgr_all, pix_all)` (optional) runs on the exchange thread after each gather has been enqueued."""

    def __init__(self, n_producers, n_tickets, group=None, slots=2, on_gathered=None):
        self.np_, self.n_tickets, self.group, self.slots, self.on_gathered = n_producers, n_tickets, group, slots, on_gathered
        self.stage = {}
        self.slot_free = {(k, s): threading.Event() for k in range(n_producers) for s in range(slots)}
        for e in self.slot_free.values():
            e.set()
        self.slot_event = {}
        self.queue = {}
        self.cv = threading.Condition()
        self.out = {}
        self.err = None
        self.last = None
        self.xstream = None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.started = False

    def submit(self, k, seq, xyz, bgr, pix, n_local):
        """Producer side; call with the producer's stream current.  Returns as soon as the snapshot copy is enqueued."""
        slot = seq % self.slots
        while not self.slot_free[(k, slot)].wait(timeout=1.0):
            if self.err is not None:
                raise RuntimeError("point exchange failed") from self.err
        self.slot_free[(k, slot)].clear()
        cuda = xyz.is_cuda
        if cuda and (k, slot) in self.slot_event:
            torch.cuda.current_stream(xyz.device).wait_event(self.slot_event[(k, slot)])  # the gather that read this slot is done
        st = self.stage.get((k, slot))
        if st is None:
            st = {"xyz": torch.empty_like(xyz), "bgr": torch.empty_like(bgr), "pix": torch.empty_like(pix)}
            self.stage[(k, slot)] = st
        for name, t in (("xyz", xyz), ("bgr", bgr), ("pix", pix)):
            st[name][:n_local].copy_(t[:n_local], non_blocking=True)
        ev = None
        if cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(xyz.device))
        with self.cv:
            self.queue[seq * self.np_ + k] = (k, slot, int(n_local), ev, xyz.device)
            if not self.started:
                self.started = True
                self.thread.start()
            self.cv.notify_all()

    def _gather_one(self, ticket, k, slot, n_local, ev, device):
        world = dist.get_world_size(self.group)
        st = self.stage[(k, slot)]
        if ev is not None:
            torch.cuda.current_stream(device).wait_event(ev)
        cnts = allgather_counts(n_local, device, self.group).cpu()
        nmax = max(int(cnts.max()), 1)
        if st["xyz"].shape[0] < nmax:
            raise ValueError(f"point buffers hold {st['xyz'].shape[0]} rows but another rank produced {nmax}")
        res = {}
        for name in ("xyz", "bgr", "pix"):
            src = st[name][:nmax]
            row = src[0].numel() if src.dim() > 1 else 1
            buf = self.out.get(name)
            need = world * st[name].shape[0] * row
            if buf is None or buf.numel() < need:
                buf = torch.empty(need, dtype=src.dtype, device=src.device)  # sized for the capacity once: no reallocation later
                self.out[name] = buf
            dst = buf[: world * nmax * row]
            dist.all_gather_into_tensor(dst, src.reshape(-1), group=self.group)
            res[name] = dst.view((world, nmax) + tuple(src.shape[1:]))
        if ev is not None:
            e2 = torch.cuda.Event()
            e2.record(torch.cuda.current_stream(device))
            self.slot_event[(k, slot)] = e2
        self.last = (cnts, res["xyz"], res["bgr"], res["pix"])
        if self.on_gathered is not None:
            self.on_gathered(ticket, *self.last)
        self.slot_free[(k, slot)].set()

    def _run(self):
        try:
            for ticket in range(self.n_tickets):
                with self.cv:
                    while ticket not in self.queue:
                        self.cv.wait(timeout=1.0)
                        if self.err is not None:
                            return
                    k, slot, n_local, ev, device = self.queue.pop(ticket)
                if device.type == "cuda":
                    if self.xstream is None:
                        self.xstream = torch.cuda.Stream(device=device)
                    with torch.cuda.stream(self.xstream):
                        self._gather_one(ticket, k, slot, n_local, ev, device)
                else:
                    self._gather_one(ticket, k, slot, n_local, ev, device)
        except BaseException as e:  # noqa: BLE001
            self.err = e
            for ev in self.slot_free.values():
                ev.set()

    def abort(self, exc):
        self.err = exc
        with self.cv:
            self.cv.notify_all()

    def finish(self):
        """Wait until every ticket has been gathered (the current CUDA stream then waits for the last gather); returns the last
        gathered buffers."""
        if self.started:
            while self.thread.is_alive():
                self.thread.join(timeout=1.0)
        if self.err is not None:
            raise RuntimeError("point exchange failed") from self.err
        if self.xstream is not None:
            torch.cuda.current_stream(self.xstream.device).wait_stream(self.xstream)
        return self.last


def concat_in_pair_order(cnts, xyz_all, bgr_all, pix_all):
    """Flatten the gathered buffers to the reference's InsertPoint order: pair 0's points, then pair 1's, ..."""
    xs, bs, ps = [], [], []
    for r, n in enumerate(cnts.tolist()):
        xs.append(xyz_all[r, :n])
        bs.append(bgr_all[r, :n])
        ps.append(pix_all[r, :n])
    return torch.cat(xs), torch.cat(bs), torch.cat(ps)
