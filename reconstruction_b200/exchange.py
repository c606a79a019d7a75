"""The one exchange step of the path: after DisparityToCloud every rank holds the points of its own
camera pair(s); the sink (CloudOptimization) wants all of them, in pair order.  Pairs shard one per rank with no
other data-path communication, so this is a count all-gather followed by one all-gather of the payload padded to the
largest count.  Works on whatever device the tensors live on: NCCL over NVLink for CUDA tensors (bench.py, one
process per GPU), gloo for the CPU tests.
"""
from __future__ import annotations

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


def concat_in_pair_order(cnts, xyz_all, bgr_all, pix_all):
    """Flatten the gathered buffers to the reference's InsertPoint order: pair 0's points, then pair 1's, ..."""
    xs, bs, ps = [], [], []
    for r, n in enumerate(cnts.tolist()):
        xs.append(xyz_all[r, :n])
        bs.append(bgr_all[r, :n])
        ps.append(pix_all[r, :n])
    return torch.cat(xs), torch.cat(bs), torch.cat(ps)
