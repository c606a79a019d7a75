#!/usr/bin/env python
"""bench.py — disparity Mpix/s of the stereo-matching hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C|B|small]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one camera pair per GPU: ConstructPyrm + MatchOneLayer for every
pyramid level + DisparityToCloud (CStereoMatching.cpp:21-29), and for N > 1 the all-gather of the per-pair point
buffers over NCCL (the one exchange step of the path).  Pairs are independent, so N GPUs process N pairs per step
("scaling": "weak").

  value   whole-job Mpix/s (N * W*H * K / time) with the rectified top-level images already in HBM
  e2e     the same through sb200_match_pair_host (the entry the C++ CStereoMatching mirror calls): pinned host
          images in, points out, H2D/D2H inside the timed region
  roofline  the dominant kernel (DisparityRefine, k_refine_fused: one launch = several sweeps of both directions):
          algorithmic 22 B per pixel-iteration x the pixel-iterations one launch covers, divided by its
          average device time, measured with CUDA events on the launching stream inside the timed region
  cpu_baseline  the reference's own CStereoMatching.cpp compiled unmodified (oracle/_ref) — or the oracle port if that
          library is absent — on the host cores, on a bounded sample, rank 0 / N=1 only

--impl reference times the reference's CPU implementation alone (same JSON shape, "impl": "reference").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "disparity Mpix/s (and 3D pts/s) at 10-view 4096x3072, 1/2/4/8 GPU vs CPU ref"
CONFIGS = {  # name -> (pyrm_num, lowest_w, lowest_h, description)
    "C": (5, 256, 192, "10-view 4096x3072 synthetic rig, 5-level pyramid (BASELINE.json configs[2]); one adjacent pair per GPU per step"),
    "B": (3, 512, 384, "10-view 2048x1536 synthetic rig, 3-level pyramid (BASELINE.json configs[1]); one adjacent pair per GPU per step"),
    "E": (5, 375, 250, "20-view 6000x4000 synthetic rig, 5-level pyramid (BASELINE.json configs[4] frame size); one adjacent pair per GPU per step"),
    "small": (3, 128, 96, "512x384 3-level smoke configuration"),
}
ALGO_BYTES_PER_PX_ITER = 22  # f64 in 8 + f64 out 8 + BGR 3 + 3


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_REAL_STDOUT = None


def protect_stdout():
    """Everything any library prints to fd 1 (NCCL's version banner, make's chatter) goes to stderr: stdout carries exactly one
    JSON line, written by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, line)
    else:
        os.write(_REAL_STDOUT, line)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_first(self, timeout=10.0):
        """Block until nvidia-smi has delivered its first sample (its start-up is then over)."""
        t0 = time.time()
        while self.proc is not None and not self.lines and time.time() - t0 < timeout:
            time.sleep(0.05)

    def mark(self):
        """Samples taken from now on belong to the timed region."""
        self.first = len(self.lines)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines[getattr(self, "first", 0):]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_run(kind, L, w0, h0, steps, warmup, threads=None, keep_outputs=False):
    """Time `steps` whole-pair runs of the CPU checker (match_pair: pyramid + all levels + cloud)."""
    from oracle import pyoracle
    from reconstruction_b200 import synth

    sp = synth.make_pair(w0, h0, L, pair_id=0)
    o = pyoracle.CpuStereo(kind, L, w0, h0)
    if threads:
        o.set_threads(threads)
    cores = o.max_threads()
    o.set_pair(*sp.image, *sp.mask)
    o.set_calib(sp.Q, sp.R_final, sp.T_final)
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)  # the reference printf()s its progress
    try:
        for _ in range(warmup):
            o.match_pair()
        t0 = time.perf_counter()
        n = 0
        for _ in range(steps):
            n = o.match_pair()
        dt = time.perf_counter() - t0
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)
    W, H = sp.top_size
    res = {"mpix_s": W * H * steps / dt / 1e6, "pts_s": n * steps / dt, "sec_per_step": dt / steps, "cores": cores,
           "sample": f"{steps} x one synthetic pair {W}x{H}, {L}-level pyramid, all stages + DisparityToCloud"}
    if keep_outputs:  # for the parity block: the checker's final maps and points of this very pair
        import numpy as np

        xyz = np.empty((n, 3))
        if n:
            import ctypes

            o._f("get_points", None, [ctypes.c_void_p] * 2)(o.h, xyz.ctypes.data_as(ctypes.c_void_p))
        res["outputs"] = {"pair": sp, "d": [o.get_disparity(k, L - 1) for k in (0, 1)], "xyz": xyz}
    return res


def pick_cpu_kind():
    from oracle import pyoracle

    try:
        pyoracle.build()
    except Exception as e:  # noqa: BLE001
        log("oracle build:", e)
    if pyoracle.available("ref"):
        return "ref", "reference"
    return "port", "port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind, label = pick_cpu_kind()
    L, w0f, h0f, _ = CONFIGS[args.config]
    total = args.steps + args.warmup
    ncpu = os.cpu_count() or 8
    # The reference needs minutes per 4096x3072 pair (OrderConstraint is O(W^2 H), CStereoMatching.cpp:337-364), so each step is a
    # bounded sample: the same pyramid depth (same sweeps per level), a smaller frame, sized so that the whole run stays within a
    # few minutes (0.24 Mpix/s measured on the GPU box's 16 host threads; half of that assumed).  Config B is small enough to be
    # timed at its full frame when the run has few steps.  A second, smaller frame is always timed once as well: the two
    # per-pixel rates show which way the frame-size bias goes (the smaller frame is FASTER per pixel, so the sampled rate
    # over-states the reference at the full frame and the GPU / CPU ratio is conservative).
    budget = 150.0 / max(total, 1)  # seconds per step
    est_rate = 0.06e6 * max(ncpu, 8) / 8
    cands = [(w0f, h0f)] if args.config in ("B", "small") else []
    cands += [(96, 72), (80, 60), (64, 48), (48, 36), (40, 30), (32, 24)]
    w0, h0 = 32, 24
    for cand in cands:
        px = (cand[0] << (L - 1)) * (cand[1] << (L - 1))
        if px / est_rate <= budget:
            w0, h0 = cand
            break
    r = cpu_run(kind, L, w0, h0, args.steps, args.warmup, threads=ncpu)  # explicit: torchrun exports OMP_NUM_THREADS=1
    small = (max(w0 * 2 // 3 // 8 * 8, 24), max(h0 * 2 // 3 // 6 * 6, 18))
    r2 = cpu_run(kind, L, small[0], small[1], 1, 0, threads=ncpu)
    same = (w0, h0) == (w0f, h0f)
    out = {
        "impl": "reference", "metric": METRIC, "value": r["mpix_s"], "unit": "Mpix/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * r["sec_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": CONFIGS[args.config][3], "pyrm_num": L, "cpu_sample": r["sample"], "same_config": same,
                   "why": None if same else (f"the reference needs minutes per {w0f << (L - 1)}x{h0f << (L - 1)} pair; a step times one pair of the same "
                                            f"{L}-level pyramid at {w0 << (L - 1)}x{h0 << (L - 1)} so that the run ends within minutes"),
                   "frame_size_bias": {"sample_Mpix_s": r["mpix_s"], "sample_top": [w0 << (L - 1), h0 << (L - 1)],
                                       "smaller_Mpix_s": r2["mpix_s"], "smaller_top": [small[0] << (L - 1), small[1] << (L - 1)],
                                       "direction": "per-pixel CPU cost grows with the frame (OrderConstraint O(W^2 H)): the reference is slower per "
                                                    "pixel at the full frame than at the sampled one, so value over-states it"}},
        "pts_per_s": r["pts_s"], "gpu_launches": 0,
        "cpu_baseline": {"value": r["mpix_s"], "unit": "Mpix/s", "cores": r["cores"], "kind": label, "sample": r["sample"]},
        "e2e": {"value": r["mpix_s"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(out)
    return 0


# ------------------------------------------------------------------------------------------ GPU arm
class _DevArr:
    """CUDA array interface view of a raw device pointer (zero copy into torch)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 3}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from reconstruction_b200 import capi, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log(f"note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the stereo path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # torch.distributed is only the launcher-side bootstrap here (unique id, barriers, max / sum of a few scalars); the data
        # path's collective is NCCL inside the library (sb200_comm_*).  SB200_BENCH_TORCH_BACKEND=gloo keeps torch off the GPUs.
        backend = os.environ.get("SB200_BENCH_TORCH_BACKEND", "nccl")
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
    red_dev = "cuda" if (world > 1 and os.environ.get("SB200_BENCH_TORCH_BACKEND", "nccl") == "nccl") else "cpu"
    capi.load()  # raises if libstereo_b200.so was not built

    L, w0, h0, desc = CONFIGS[args.config]
    W, H = w0 << (L - 1), h0 << (L - 1)
    npx = W * H
    NC = max(1, args.pairs_in_flight)  # camera pairs in flight per GPU: one context + one stream + one host thread each
    t0 = time.time()
    pairs = [synth.make_pair(w0, h0, L, pair_id=2 * rank + k) for k in range(2)]  # two inputs, alternated between steps
    log(f"[rank {rank}] synthetic pairs {W}x{H} built in {time.time() - t0:.1f}s")
    gs = [capi.StereoB200(L, w0, h0, device=local) for _ in range(NC)]
    streams = [torch.cuda.ExternalStream(g.stream(), device=local) for g in gs]

    # device-resident inputs (for `value`) and pinned host inputs / outputs (for `e2e`); inputs are read-only and shared
    dev_in, pin_in = [], []
    for sp in pairs:
        host = [torch.from_numpy(a) for a in (*sp.image, *sp.mask)]
        pin_in.append([h.pin_memory() for h in host])
        dev_in.append([h.cuda(local) for h in host])
    pin_xyz = [torch.empty((npx, 3), dtype=torch.float64).pin_memory() for _ in range(NC)]
    torch.cuda.synchronize()

    # N > 1: the exchange step goes through the C ABI (sb200_comm_* / sb200_exchange_*, csrc/comm.cu): the rank's contexts only
    # snapshot their points; the communicator's exchange thread issues the NCCL collectives (counts, then one grouped un-padded
    # broadcast per rank) in ticket order on its own low-priority stream, so no context ever waits for another rank.
    # torch.distributed only carries the unique id, the barriers and the max over ranks of the timings.
    comm = None
    if world > 1 and not args.no_exchange:
        box = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        comm = capi.PointComm(local, rank, world, box[0], producers=NC, slots=int(os.environ.get("SB200_BENCH_SLOTS", "2")))
    xch = [None]      # the communicator while a timed region with the exchange is running
    seq_base = [0]    # tickets are numbered over the communicator's lifetime

    def exchange(k, seq, n_local):
        """Snapshot of the pair's points for the all-gather, overlapped with the following pairs' matching; the last gathers are
        waited for inside the timed region."""
        if xch[0] is None:
            return
        xch[0].submit(gs[k], k, seq_base[0] + seq)

    def step_resident(k, i):
        sp = pairs[(i + k) % 2]
        gs[k].set_calib(sp.Q, sp.R_final, sp.T_final)
        gs[k].stage_device(*dev_in[(i + k) % 2])
        n = gs[k].match_pair()
        exchange(k, i, n)
        return n

    def step_e2e(k, i):
        sp = pairs[(i + k) % 2]
        # exactly the call the C++ CStereoMatching::MatchAllLayer mirror makes (isoutput = 0): the InsertPoint payload (xyz f64) out
        n = gs[k].match_pair_host(*pin_in[(i + k) % 2], sp.Q, sp.R_final, sp.T_final, pin_xyz[k], None, None, npx)
        exchange(k, i, n)
        return n

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, ctxs, profile=False):
        """`steps` steps on every context in `ctxs` (one host thread per context when there are several), device time from the
        first enqueue to the completion of every context's last step, max over ranks."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for k in ctxs:
            gs[k].set_profiling(profile)
            if profile:
                gs[k].stage_ms(reset=True)
                gs[k].refine_profile(reset=True)
        launches0 = sum(gs[k].launch_count() for k in ctxs)
        xch[0] = comm if (comm is not None and len(ctxs) == NC) else None
        npts = [0] * NC
        errs = []

        trace = os.environ.get("SB200_BENCH_TRACE") == "1"

        def work(k):
            # No torch CUDA state is touched in the worker threads (the C ABI brings its own streams).  A `with torch.cuda.stream(...)`
            # here restored the thread's previous device - device 0 in a fresh thread - when the first worker finished, which made
            # every rank but 0 create a CUDA context on GPU 0 inside the timed region: 0.3 - 2 s during which the process's other
            # CUDA calls stalled (found with SB200_BENCH_TRACE=1 on two GPUs: only rank 1, only the last step, all contexts at once).
            try:
                for i in range(steps):
                    t_a = time.perf_counter()
                    npts[k] = step_fn(k, i)
                    if trace:
                        log(f"[trace rank {rank} ctx {k} {step_fn.__name__} step {i}] {1e3 * (time.perf_counter() - t_a):.1f} ms "
                            f"(graph launches / instantiations {gs[k].graph_info()})")
            except BaseException as e:  # noqa: BLE001
                errs.append(e)

        barrier()
        lead = streams[ctxs[0]]
        ev0.record(lead)  # the device is idle here (barrier above), so this is the start of every context's work
        if len(ctxs) == 1:
            work(ctxs[0])
        else:
            th = [threading.Thread(target=work, args=(k,)) for k in ctxs]
            for t in th:
                t.start()
            for t in th:
                t.join()
        if errs:
            raise errs[0]
        with torch.cuda.stream(lead):
            if xch[0] is not None:
                seq_base[0] += steps
                xch[0].drain(seq_base[0] * NC)  # the last steps' gathers complete inside the timed region
            for k in ctxs[1:]:
                lead.wait_stream(streams[k])
            ev1.record(lead)
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=red_dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, npts[ctxs[0]], sum(gs[k].launch_count() for k in ctxs) - launches0

    # The clock sampler (nvidia-smi in loop mode) is started here, before the warm-up, so that its start-up (NVML attaches to every
    # GPU of the box) is over long before the timed region; only the samples taken inside the timed region are used.
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    every = list(range(NC))
    for k in every:
        with torch.cuda.stream(streams[k]):
            for i in range(args.warmup):
                gs[k].set_calib(pairs[i % 2].Q, pairs[i % 2].R_final, pairs[i % 2].T_final)
                gs[k].stage_device(*dev_in[i % 2])
                gs[k].match_pair()

    # (1) one pair at a time on context 0, with the per-stage / per-launch timers on: the roofline and stage figures
    g = gs[0]
    single_ms, n_pts, single_launches = None, 0, 0
    if NC > 1:
        single_ms, n_pts, single_launches = timed(step_resident, args.steps, [0], profile=True)  # the matcher alone, no exchange
    if comm is not None:
        timed(step_resident, 1, every)  # untimed: the exchange's staging and gather buffers get allocated here
        comm.stats(reset=True)
    if rank == 0:
        sampler.wait_first()
        sampler.mark()
    # (2) the headline: NC pairs in flight per GPU
    ms, n_pts2, launches = timed(step_resident, args.steps, every, profile=(NC == 1))
    clocks = sampler.stop() if rank == 0 else None
    xstats = comm.stats(reset=True) if comm is not None else None
    n_pts = n_pts or n_pts2
    stage_ms = g.stage_ms(reset=True)
    ncc_ms = g.stage_level_ms(2, L - 1)  # HighLevelInitialMatch at the top level, both directions
    ncc_px = int(sum(int(m[4]) * int(m[5]) for m in g.get_margins(L - 1)))
    counters = g.refine_counters(reset=True)
    top_ms, top_n, top_px = g.refine_profile(level=L - 1, reset=False)
    all_ms, all_n, all_px = g.refine_profile(level=-1, reset=True)
    g.set_profiling(False)

    for k in every:
        with torch.cuda.stream(streams[k]):
            for i in range(min(args.warmup, 2)):
                gs[k].match_pair_host(*pin_in[i % 2], pairs[i % 2].Q, pairs[i % 2].R_final, pairs[i % 2].T_final, pin_xyz[k], None, None, npx)
    e2e_ms, n_pts_e2e, _ = timed(step_e2e, args.steps, every)

    # Rectify (the step before the hot path), measured once per run, outside the timed steps: host calibration +
    # H2D of the two original frames + device maps / remap / erode + pyramid
    rectify = None
    if rank == 0:
        from reconstruction_b200 import stage as _stage

        cams = _stage.rig_cameras(2, W, H)
        cal = capi.rectify_calib(cams[0][0], cams[0][1], cams[1][0], cams[1][1], (W, H), w0, L)
        g2 = capi.StereoB200(L, w0, h0, W, H, device=local)
        s2 = torch.cuda.ExternalStream(g2.stream(), device=local)
        ts = []
        for it in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(s2):
                e0.record(s2)
                for j in (0, 1):
                    g2.rectify_view(j, pin_in[0][j], pin_in[0][2 + j], cams[j][0], cal["R_new"][j], cal["P_scaled"][j])
                g2.pair_build()
                e1.record(s2)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        g2.close()
        rectify = {"ms_per_pair": min(ts), "what": "sb200_rectify_view x2 + sb200_pair_build at the top size from pinned original frames "
                   "(H2D 100.7 MB inside), best of 3", "Mpix_per_s": 2 * npx / (min(ts) * 1e-3) / 1e6}

    # the sink's per-pair filter (the step after the hot path) on the points of the last end-to-end pair, measured once
    # per run outside the timed steps: H2D of the f64 points + outlier removal + normals + orientation; D2H of the records excluded
    sink = None
    if rank == 0 and n_pts_e2e > 0:
        from reconstruction_b200 import stage as _stage

        cams = _stage.rig_cameras(2, W, H)
        center = -cams[0][1][:, :3].T @ cams[0][1][:, 3]
        pts = pin_xyz[0][:n_pts_e2e].numpy()
        best = None
        try:
            for it in range(2):
                t0 = time.perf_counter()
                _rec, kept, st_ = capi.sink_filter(pts, 100, 1.0, 2.5, center, device=local)  # CReconstruction.cpp:19 parameters
                wall = time.perf_counter() - t0
                if best is None or st_["device_ms"] < best["device_ms"]:
                    best = {"points_in": int(n_pts_e2e), "points_kept": int(len(kept)), "device_ms": st_["device_ms"], "wall_ms": 1e3 * wall,
                            "Mpts_per_s": n_pts_e2e / (st_["device_ms"] * 1e-3) / 1e6, "widened_queries": st_["widened_queries"],
                            "what": "sb200_sink_filter (meanK 100, stddev x1, normal radius 2.5): float narrowing, two cell sorts, k-nearest mean "
                                    "distances by radix select, host mean/stddev, keep flags + scan, covariance normals; best of 2"}
        except Exception as e:  # noqa: BLE001  (a side measurement must not take the headline down with it)
            best = {"unavailable": str(e)[:200]}
        sink = best

    # native JPEG decode of one frame of the named size (the data format in front of the path): host, one thread
    decode = None
    if rank == 0:
        try:
            import cv2
            import tempfile

            host_bin = os.path.join(ROOT, "reconstruction_b200", "host", "reconstruction")
            ok, jb = cv2.imencode(".jpg", pairs[0].image[0], [cv2.IMWRITE_JPEG_QUALITY, 95])
            if ok and os.path.exists(host_bin):
                with tempfile.NamedTemporaryFile(suffix=".jpg") as tf:
                    tf.write(jb.tobytes())
                    tf.flush()
                    r = subprocess.run([host_bin, "--decode-bench", tf.name, "3"], capture_output=True, text=True, timeout=120)
                    decode = json.loads(r.stdout)
                    decode["what"] = "sbcv::imdecode (baseline JPEG, quality 95, 4:2:0) of one top-size frame, one host thread, best of 3"
        except Exception as e:  # noqa: BLE001
            decode = {"unavailable": str(e)[:200]}

    # totals over ranks
    tot_pts = n_pts * NC
    tot_launch = launches
    if world > 1:
        t = torch.tensor([n_pts * NC, launches], dtype=torch.int64, device=red_dev)
        dist.all_reduce(t)
        tot_pts, tot_launch = int(t[0].item()), int(t[1].item())

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        traffic, traffic_note = None, None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic, traffic_note = tj.get("dram_bytes_per_launch"), tj.get("note")
        achieved = ALGO_BYTES_PER_PX_ITER * top_px / (top_ms * 1e-3) / 1e9 if top_ms > 0 else None
        achieved_all = ALGO_BYTES_PER_PX_ITER * all_px / (all_ms * 1e-3) / 1e9 if all_ms > 0 else None
        sec = ms * 1e-3
        prof_steps = args.steps  # steps the profiled pass ran on context 0
        out = {
            "metric": METRIC, "value": world * NC * npx * args.steps / sec / 1e6, "unit": "Mpix/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc.replace("one adjacent pair per GPU per step", f"{NC} adjacent pair(s) per GPU per step"),
                       "pairs_per_gpu_per_step": NC,
                       "pairs_in_flight": f"{NC} contexts per GPU (one stream and one host thread each): a step matches {NC} camera pairs "
                                          "concurrently, as the C++ CStereoMatching mirror does (SB200_CTX_PER_DEVICE)",
                       "top_size": [W, H], "pyrm_num": L,
                       "l2": f"inputs larger than L2: ~{2.4 * NC:.1f} GB working set per step, two alternating input pairs",
                       "exchange": ("none (1 GPU)" if world == 1 else "off (--no-exchange: A/B run)" if comm is None else
                                    "sb200_exchange_submit / sb200_exchange_drain (C ABI): NCCL all-gather of the counts + one grouped "
                                    "un-padded broadcast per rank, overlapped with the next pair's matching")},
            "exchange": None if xstats is None else {
                "collective_ms_per_step": xstats["collective_ms"] / args.steps, "bytes_received_per_step": xstats["bytes_received"] // args.steps,
                "exchanges_per_step": xstats["exchanges"] / args.steps, "nccl_max_ctas": int(os.environ.get("SB200_NCCL_MAX_CTAS", "16")), "staging_slots": int(os.environ.get("SB200_BENCH_SLOTS", "2")),
                "what": "device time of the collectives on this rank's exchange stream (CUDA events), rank 0; they run beside the matching"},
            "pts_per_s": tot_pts * args.steps / sec, "points_per_pair": n_pts,
            "gpu_launches": tot_launch,
            "e2e": {"value": world * NC * npx * args.steps / (e2e_ms * 1e-3) / 1e6, "unit": "Mpix/s",
                    "h2d_bytes_per_step": 4 * npx * 2 * NC, "d2h_bytes_per_step": int(n_pts_e2e) * 24 * NC,
                    "ms_per_step": e2e_ms / args.steps, "api": "sb200_match_pair_host (pinned host buffers), one call per pair and context"},
            "one_pair_at_a_time": None if single_ms is None else {
                "value": world * npx * args.steps / (single_ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_pair": single_ms / args.steps,
                "gpu_launches": single_launches,
                "what": "the same steps on one context (one pair in flight per GPU), timed the same way just before the headline region; "
                        "roofline, ncc_top_level and stage_ms_per_step are taken from this pass so that another pair's kernels do not "
                        "sit between the timing events"},
            "roofline": {"bound": "hbm", "kernel": "k_refine_fused (DisparityRefine: one launch = several Jacobi sweeps of both matching directions, top pyramid level)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                         "limiter": "instruction issue, not HBM: 61 FP64 instructions per pixel-sweep (fixed by bit-exactness) hold the dispatch "
                                    "port two cycles each, plus ~70 others; DRAM moves 0.41x the algorithmic bytes (DESIGN.md 4, 7)",
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PX_ITER * top_px / top_n if top_n else None,
                         "avg_launch_us": 1e3 * top_ms / top_n if top_n else None, "launches_timed": top_n,
                         "all_levels": {"achieved": achieved_all, "frac": (achieved_all / peak) if achieved_all else None,
                                        "launches_timed": all_n}},
            # the NCC cost-volume stage the north star names (HighLevelInitialMatch, top level, both directions: integer
            # statistics + ranges + dp4a screening + exact FP64 pass): 12 algorithmic B per source-margin pixel
            "ncc_top_level": {"ms_per_step": ncc_ms / prof_steps, "pixels": ncc_px,
                              "achieved_GBps": (12 * ncc_px * prof_steps / (ncc_ms * 1e-3) / 1e9) if ncc_ms > 0 else None,
                              "frac_of_hbm_peak": (12 * ncc_px * prof_steps / (ncc_ms * 1e-3) / 1e9 / peak) if ncc_ms > 0 else None,
                              "exact_fallback_pixels_per_step": int(counters[0]) // max(prof_steps, 1),
                              "kernel": "k_ncc_band (TMA-staged 16 x 128 band, one 2x2 quad per thread, exact integer keys) + k_hole_ranges + "
                                        "list kernels; both directions on two streams (DESIGN.md 3.1)"},
            "rectify": rectify,
            "sink_filter": sink,
            "jpeg_decode": decode,
            "stage_ms_per_step": {k: round(float(stage_ms[i]) / prof_steps, 4) for i, k in enumerate(
                ["pyramid", "FindMargin", "InitialMatch", "Smooth", "Order", "Unique1", "Rematch", "Unique2", "Median", "Refine",
                 "Unique3", "ToCloud", "RefineSweepsOnly"])},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            kind, label = pick_cpu_kind()
            ncpu = os.cpu_count() or 8
            w0c, h0c = (64, 48) if ncpu < 16 else (96, 72)  # ~10-30 s of CPU work
            r = cpu_run(kind, L, w0c, h0c, 1, 0, threads=ncpu, keep_outputs=True)
            out["cpu_baseline"] = {"value": r["mpix_s"], "unit": "Mpix/s", "cores": r["cores"], "kind": label, "sample": r["sample"],
                                   "sec": r["sec_per_step"]}
            # parity: the pair the CPU arm just processed, through the CUDA path, compared bit for bit (the checker's outputs are
            # only read here; nothing measured above depends on them)
            try:
                import numpy as np

                ro = r["outputs"]
                spc = ro["pair"]
                gp = capi.StereoB200(L, w0c, h0c, device=local)
                gp.set_pair(*spc.image, *spc.mask)
                gp.set_calib(spc.Q, spc.R_final, spc.T_final)
                ng = gp.match_pair()
                diff = sum(int((gp.get_disparity(k).view(np.int64) != ro["d"][k].view(np.int64)).sum()) for k in (0, 1))
                gx, _, _ = gp.get_points(ng)
                pdiff = -1 if ng != len(ro["xyz"]) else int((gx.view(np.int64) != ro["xyz"].view(np.int64)).any(axis=1).sum())
                out["parity"] = {"against": label, "config": r["sample"], "px_compared": int(2 * ro["d"][0].size), "px_differing": diff,
                                 "points_compared": int(len(ro["xyz"])), "points_differing": pdiff, "bar": "bit-exact (f64 maps and xyz)",
                                 "larger_cases": "tests/test_gpu_baseline_parity.py: every dump point at 2048x1536 L3, 1536x1152 L5, 1504x1008 L5"}
                gp.close()
            except Exception as e:  # noqa: BLE001
                out["parity"] = {"unavailable": str(e)[:200]}
        emit(out)
    if comm is not None:
        comm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--config", choices=sorted(CONFIGS), default="C")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-exchange", action="store_true", help="N > 1: leave the point all-gather out (A/B measurement of its cost)")
    ap.add_argument("--pairs-in-flight", type=int, default=3, help="camera pairs matched concurrently per GPU (contexts per GPU)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    protect_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
