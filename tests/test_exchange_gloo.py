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
