"""Writes a data set in the reference's on-disk schema (config.yml + calib_camera.yml, CManageData.cpp:26-66;
writer example BatchProcess/main.cpp:47-73) plus the staged/ directory the host mirror's Rectify reads
(reconstruction_b200/host/CStereoMatching.cpp).  Used by the tests and by tools/stage_rig.py; not on the timed path.
"""
from __future__ import annotations

import os

import numpy as np

from . import synth


def _mat(name, a, dt="d"):
    a = np.asarray(a)
    rows, cols = (a.shape[0], 1) if a.ndim == 1 else a.shape
    flat = a.reshape(-1)
    vals = ", ".join(str(int(v)) for v in flat) if dt == "u" else ", ".join(repr(float(v)) for v in flat)
    return f"{name}: !!opencv-matrix\n   rows: {rows}\n   cols: {cols}\n   dt: {dt}\n   data: [ {vals} ]\n"


def write_pnm(path, arr):
    arr = np.ascontiguousarray(arr)
    with open(path, "wb") as f:
        if arr.ndim == 3:  # BGR in memory -> RGB on disk
            f.write(b"P6\n%d %d\n255\n" % (arr.shape[1], arr.shape[0]))
            f.write(arr[:, :, ::-1].tobytes())
        else:
            f.write(b"P5\n%d %d\n255\n" % (arr.shape[1], arr.shape[0]))
            f.write(arr.tobytes())


def rig_cameras(n_cam, width, height):
    """Cameras on a horizontal arc looking at the origin: K and [R|t] per camera."""
    out = []
    for i in range(n_cam):
        ang = np.deg2rad(6.0 * (i - (n_cam - 1) / 2))
        c = np.array([1000.0 * np.sin(ang), 0.0, -1000.0 * np.cos(ang)])
        z = -c / np.linalg.norm(c)
        x = np.cross([0.0, 1.0, 0.0], z)
        x /= np.linalg.norm(x)
        y = np.cross(z, x)
        r = np.stack([x, y, z])
        k = np.array([[0.9 * width, 0, width / 2], [0, 0.9 * width, height / 2], [0, 0, 1.0]])
        out.append((k, np.concatenate([r, (-r @ c)[:, None]], axis=1)))
    return out


def write_dataset(root, pyrm_num, lowest_w, lowest_h, n_pairs=1, isoutput=0, pair_id0=0, origin_scale=1.0, block_style=True, distinct=None):
    """n_pairs adjacent pairs (i, i+1) of an (n_pairs+1)-camera rig; returns (config path, [StagedPair]).
    distinct: synthesise only that many different pairs and repeat them (hard links) - for large throughput data sets."""
    os.makedirs(os.path.join(root, "staged"), exist_ok=True)
    root = os.path.join(os.path.abspath(root), "")
    top_w, top_h = lowest_w << (pyrm_num - 1), lowest_h << (pyrm_num - 1)
    n_syn = n_pairs if not distinct else min(distinct, n_pairs)
    pairs = [synth.make_pair(lowest_w, lowest_h, pyrm_num, pair_id=pair_id0 + p, origin_scale=origin_scale) for p in range(n_syn)]
    pairs = [pairs[p % n_syn] for p in range(n_pairs)]
    ow, oh = pairs[0].origin_size
    n_cam = n_pairs + 1
    cams = rig_cameras(n_cam, ow, oh)
    with open(root + "calib_camera.yml", "w") as f:
        f.write("%YAML:1.0\n---\n")
        for i, (k, rt) in enumerate(cams):
            f.write(_mat(f"intrinsic-{i}", k))
            f.write(_mat(f"extrinsic-{i}", rt))
    names = [f"{1:04d}_Cam{i}.ppm" for i in range(n_cam)]
    with open(root + "config.yml", "w") as f:
        f.write("%YAML:1.0\n---\n")
        f.write(f'filepath: "{root}"\n')
        f.write(f'outfilename: "{root}out.ply"\n')
        f.write(f"isoutput: {isoutput}\n")
        f.write("camera_calib_name: calib_camera.yml\n")
        f.write(f"PyrmNum: {pyrm_num}\nLowestLevelWidth: {lowest_w}\nLowestLevelHeight: {lowest_h}\n")
        f.write(f"OriginWidth: {ow}\nOriginHeight: {oh}\n")
        if block_style:
            f.write("imagelist:\n" + "".join(f'   - "{n}"\n' for n in names))
            f.write("masklist:\n" + "".join(f'   - "mask/{n}"\n' for n in names))
        else:
            f.write("imagelist: [ " + ", ".join(f'"{n}"' for n in names) + " ]\n")
            f.write("masklist: [ " + ", ".join(f'"mask/{n}"' for n in names) + " ]\n")
        f.write(_mat("camID", np.array([[p, p + 1] for p in range(n_pairs)]), dt="u"))
    for p, sp in enumerate(pairs):
        with open(root + f"staged/pair{p}.yml", "w") as f:
            f.write("%YAML:1.0\n---\n")
            f.write(_mat("Q", sp.Q))
            f.write(_mat("R_final", sp.R_final))
            f.write(_mat("T_final", sp.T_final))
            for k in (0, 1):
                f.write(_mat(f"P{k}", np.concatenate([cams[p + k][0] @ cams[p + k][1][:, :3], (cams[p + k][0] @ cams[p + k][1][:, 3])[:, None]], axis=1)))
        for k in (0, 1):
            for what, ext, arr in (("view", "ppm", sp.image[k]), ("mask", "pgm", sp.mask[k])):
                dst = root + f"staged/pair{p}_{what}{k}.{ext}"
                if p >= n_syn:  # a repeated pair: link to the first copy
                    if os.path.exists(dst):
                        os.remove(dst)
                    os.link(root + f"staged/pair{p % n_syn}_{what}{k}.{ext}", dst)
                else:
                    write_pnm(dst, arr)
    return root + "config.yml", pairs


def write_raw_dataset(root, pyrm_num, lowest_w, lowest_h, origin, cams, images, masks, pairs, isoutput=0, fmt="pnm"):
    """A data set of ORIGINAL (unrectified) frames in the reference's schema: config.yml, calib_camera.yml and one image
    + mask per camera.  The host mirror then runs Rectify natively (no staged/ directory).  cams: [(K, Rt)] per camera,
    images / masks: per camera arrays at `origin` size, pairs: [[camA, camB], ...].  fmt: "pnm" (PPM / PGM), or "jpg" / "png"
    written by OpenCV like the reference's own data sets ("%.4d_Cam%d.jpg", BatchProcess/main.cpp:66)."""
    os.makedirs(os.path.join(root, "mask"), exist_ok=True)
    root = os.path.join(os.path.abspath(root), "")
    with open(root + "calib_camera.yml", "w") as f:
        f.write("%YAML:1.0\n---\n")
        for i, (k, rt) in enumerate(cams):
            f.write(_mat(f"intrinsic-{i}", k))
            f.write(_mat(f"extrinsic-{i}", rt))
    ext = {"pnm": ("ppm", "pgm"), "jpg": ("jpg", "jpg"), "png": ("png", "png")}[fmt]
    names = [f"{1:04d}_Cam{i}.{ext[0]}" for i in range(len(cams))]
    mask_names = [f"mask/{1:04d}_Cam{i}.{ext[1]}" for i in range(len(cams))]
    for i, n in enumerate(names):
        img = images[i] if images[i].ndim == 3 else np.repeat(images[i][:, :, None], 3, axis=2)
        if fmt == "pnm":
            write_pnm(root + n, img)
            write_pnm(root + mask_names[i], masks[i])
        else:
            import cv2

            par = [cv2.IMWRITE_JPEG_QUALITY, 95] if fmt == "jpg" else []
            assert cv2.imwrite(root + n, np.ascontiguousarray(img), par) and cv2.imwrite(root + mask_names[i], np.ascontiguousarray(masks[i]), par)
    with open(root + "config.yml", "w") as f:
        f.write("%YAML:1.0\n---\n")
        f.write(f'filepath: "{root}"\n')
        f.write(f'outfilename: "{root}out.ply"\n')
        f.write(f"isoutput: {isoutput}\n")
        f.write("camera_calib_name: calib_camera.yml\n")
        f.write(f"PyrmNum: {pyrm_num}\nLowestLevelWidth: {lowest_w}\nLowestLevelHeight: {lowest_h}\n")
        f.write("imagelist:\n" + "".join(f'   - "{n}"\n' for n in names))
        f.write("masklist:\n" + "".join(f'   - "{n}"\n' for n in mask_names))
        f.write(_mat("camID", np.array(pairs), dt="u"))
    return root + "config.yml"


def read_ply_f32(path):
    """(xyz float32 [n,3], bgr uint8 [n,3]) of a cloud written in the reference's PLY layout."""
    with open(path, "rb") as f:
        n = 0
        while True:
            line = f.readline()
            if line.startswith(b"element vertex"):
                n = int(line.split()[-1])
            if line.strip() == b"end_header":
                break
        rec = np.frombuffer(f.read(15 * n), dtype=np.dtype([("xyz", "<f4", 3), ("bgr", "u1", 3)]))
    return rec["xyz"].copy(), rec["bgr"].copy()
