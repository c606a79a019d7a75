"""N > 1 host logic on CPU: world_size-2 gloo run of the point exchange (reconstruction_b200/exchange.py), the step
bench.py performs over NCCL after DisparityToCloud."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from reconstruction_b200 import exchange


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _points(rank, n):
    rng = np.random.default_rng(100 + rank)
    return rng.standard_normal((n, 3)), rng.integers(0, 256, (n, 3), dtype=np.uint8), np.sort(rng.choice(10_000, n, replace=False)).astype(np.int32)


def _worker(rank, world, port, counts, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cap = 64
        n = counts[rank]
        x, b, p = _points(rank, n)
        xyz = torch.zeros((cap, 3), dtype=torch.float64); xyz[:n] = torch.from_numpy(x)
        bgr = torch.zeros((cap, 3), dtype=torch.uint8); bgr[:n] = torch.from_numpy(b)
        pix = torch.zeros(cap, dtype=torch.int32); pix[:n] = torch.from_numpy(p)
        bufs = {}
        for _ in range(2):  # second call reuses the gather buffers
            cnts, xa, ba, pa = exchange.allgather_points(xyz, bgr, pix, n, out=bufs)
        fx, fb, fp = exchange.concat_in_pair_order(cnts, xa, ba, pa)
        # the overlapped form: submit, clobber the source buffers (as the next pair's matching would), then finish
        ex = exchange.PointExchanger()
        for _ in range(2):
            ex.submit(xyz, bgr, pix, n)
            keep = (xyz.clone(), bgr.clone(), pix.clone())
            xyz.fill_(-1.0); bgr.fill_(7); pix.fill_(-5)
            c2, xa2, ba2, pa2 = ex.finish()
            xyz.copy_(keep[0]); bgr.copy_(keep[1]); pix.copy_(keep[2])
        gx, gb, gp = exchange.concat_in_pair_order(c2, xa2, ba2, pa2)
        assert torch.equal(gx, fx) and torch.equal(gb, fb) and torch.equal(gp, fp) and c2.tolist() == cnts.tolist()
        q.put((rank, cnts.tolist(), fx.numpy().copy(), fb.numpy().copy(), fp.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_allgather_points_world2():
    world, counts = 2, [17, 40]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, counts, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ex = np.concatenate([_points(r, counts[r])[0] for r in range(world)])
    eb = np.concatenate([_points(r, counts[r])[1] for r in range(world)])
    ep = np.concatenate([_points(r, counts[r])[2] for r in range(world)])
    for rank, cnts, fx, fb, fp in got:
        assert cnts == counts
        assert np.array_equal(fx.view(np.int64), ex.view(np.int64)), f"rank {rank}: xyz not in pair order"
        assert np.array_equal(fb, eb) and np.array_equal(fp, ep)


def test_pair_sharding_map():
    for world in (1, 2, 4, 8):
        seen = sorted(p for r in range(world) for p in exchange.pairs_of_rank(r, world, 10))
        assert seen == list(range(10))
        assert all(exchange.pair_to_rank(p, world) == p % world for p in range(10))
    assert exchange.pairs_of_rank(3, 8, 10) == [3]
    assert exchange.pairs_of_rank(1, 8, 10) == [1, 9]


def _ordered_worker(rank, world, port, q):
    """Two producer threads (contexts) per rank, three pairs each, submitted in whatever order the threads get there."""
    import threading
    import time

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cap, NP, STEPS = 48, 2, 3
        seen = []

        def on_gathered(ticket, cnts, xa, ba, pa):
            fx, fb, fp = exchange.concat_in_pair_order(cnts, xa, ba, pa)
            seen.append((ticket, cnts.tolist(), fx.numpy().copy(), fb.numpy().copy(), fp.numpy().copy()))

        ex = exchange.OrderedPointExchange(NP, NP * STEPS, on_gathered=on_gathered)

        def producer(k):
            xyz = torch.zeros((cap, 3), dtype=torch.float64)
            bgr = torch.zeros((cap, 3), dtype=torch.uint8)
            pix = torch.zeros(cap, dtype=torch.int32)
            for seq in range(STEPS):
                time.sleep(0.01 * ((rank + k + seq) % 3))  # ranks and contexts drift apart
                n = 5 + 7 * rank + 3 * k + seq
                x, b, p = _points(1000 * rank + 10 * k + seq, n)
                xyz[:n] = torch.from_numpy(x); bgr[:n] = torch.from_numpy(b); pix[:n] = torch.from_numpy(p)
                ex.submit(k, seq, xyz, bgr, pix, n)
                xyz.fill_(-1.0); bgr.fill_(9); pix.fill_(-3)  # the next pair's matching overwrites the buffers

        th = [threading.Thread(target=producer, args=(k,)) for k in range(NP)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        ex.finish()
        q.put((rank, seen))
    finally:
        dist.destroy_process_group()


def test_ordered_exchange_several_pairs_in_flight():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ordered_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in range(world):
        seen = got[rank]
        assert [t for t, *_ in seen] == list(range(6))  # gathers were issued in ticket order on every rank
        for ticket, cnts, fx, fb, fp in seen:
            seq, k = divmod(ticket, 2)
            exp_counts = [5 + 7 * r + 3 * k + seq for r in range(world)]
            assert cnts == exp_counts
            ex = np.concatenate([_points(1000 * r + 10 * k + seq, exp_counts[r])[0] for r in range(world)])
            eb = np.concatenate([_points(1000 * r + 10 * k + seq, exp_counts[r])[1] for r in range(world)])
            ep = np.concatenate([_points(1000 * r + 10 * k + seq, exp_counts[r])[2] for r in range(world)])
            assert np.array_equal(fx.view(np.int64), ex.view(np.int64)) and np.array_equal(fb, eb) and np.array_equal(fp, ep), (rank, ticket)
