"""The exchange step through the C ABI (sb200_comm_*, reconstruction_b200/csrc/comm.cu): every
rank contributes the points of its camera pair and receives all pairs' points in pair order
(CloudOptimization/CCloudOptimization.cpp:123 appends the clouds pair by pair).

  * world 1 (any GPU box): the synchronous call and the overlapped submit / wait form with two producers whose tickets arrive
    out of order, against the points read back directly;
  * world 2 (needs two GPUs, skipped otherwise): two processes, one pair each, NCCL over NVLink; both ranks must end up with
    pair 0's points followed by pair 1's, bit for bit."""
import multiprocessing as mp
import os
import threading

import numpy as np
import pytest

from reconstruction_b200 import capi, synth

pytestmark = pytest.mark.gpu
L, W0, H0 = 2, 96, 72


@pytest.fixture(scope="module")
def lib():
    import torch

    assert torch.cuda.is_available(), "these tests need the B200"
    capi.build()
    return capi.load()


def _matched(pair_id, device=0):
    sp = synth.make_pair(W0, H0, L, pair_id=pair_id)
    g = capi.StereoB200(L, W0, H0, device=device)
    g.set_pair(*sp.image, *sp.mask)
    g.set_calib(sp.Q, sp.R_final, sp.T_final)
    n = g.match_pair()
    return g, n


def test_world1_synchronous(lib):
    g, n = _matched(31)
    xyz, bgr, pix = g.get_points(n)
    comm = capi.PointComm(0, 0, 1, capi.comm_unique_id())
    for _ in range(3):  # slots are reused
        counts, gx, gb, gp = comm.allgather_points(g, capacity=n + 5)
        assert counts.tolist() == [n]
        assert np.array_equal(gx.view(np.int64), xyz.view(np.int64)) and np.array_equal(gb, bgr) and np.array_equal(gp, pix)
    st = comm.stats()
    assert st["exchanges"] == 3 and st["bytes_received"] == 3 * n * 31
    comm.close()
    g.close()


def test_world1_two_producers_out_of_order(lib):
    """tickets = seq * producers + producer; producer 1 submits before producer 0: the exchange thread must still gather in ticket order"""
    ctxs = [_matched(32), _matched(33)]
    pts = [g.get_points(n) for g, n in ctxs]
    comm = capi.PointComm(0, 0, 1, capi.comm_unique_id(), producers=2, slots=2)
    n_seq = 3

    def producer(k, delay):
        import time

        for seq in range(n_seq):
            time.sleep(delay)
            comm.submit(ctxs[k][0], k, seq)

    th = [threading.Thread(target=producer, args=(1, 0.0)), threading.Thread(target=producer, args=(0, 0.05))]
    for t in th:
        t.start()
    for ticket in range(2 * n_seq):
        k = ticket % 2
        n = ctxs[k][1]
        counts, gx, gb, gp = comm.wait(ticket, capacity=n, want_host=True)
        assert counts.tolist() == [n], ticket
        assert np.array_equal(gx.view(np.int64), pts[k][0].view(np.int64)) and np.array_equal(gp, pts[k][2]), ticket
    for t in th:
        t.join()
    comm.drain(2 * n_seq)
    comm.close()
    for g, _ in ctxs:
        g.close()


def _rank_main(rank, world, uid, q):
    try:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        capi.load()
        g, n = _matched(40 + rank, device=rank)
        mine = g.get_points(n)
        comm = capi.PointComm(rank, rank, world, uid)
        counts, gx, gb, gp = comm.allgather_points(g, capacity=4 * W0 * H0 * 4)
        comm.close()
        g.close()
        q.put((rank, counts.tolist(), gx, gb, gp, mine, None))
    except BaseException as e:  # noqa: BLE001
        q.put((rank, None, None, None, None, None, repr(e)))


def test_world2_two_processes(lib):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    uid = capi.comm_unique_id()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, uid, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        r = q.get(timeout=300)
        assert r[6] is None, r[6]
        res[r[0]] = r
    for p in procs:
        p.join(timeout=60)
    n0, n1 = len(res[0][5][0]), len(res[1][5][0])
    exp_xyz = np.concatenate([res[0][5][0], res[1][5][0]])
    exp_pix = np.concatenate([res[0][5][2], res[1][5][2]])
    for r in (0, 1):
        assert res[r][1] == [n0, n1]
        assert np.array_equal(res[r][2].view(np.int64), exp_xyz.view(np.int64)), f"rank {r}: gathered xyz"
        assert np.array_equal(res[r][4], exp_pix), f"rank {r}: gathered pixel indices"
