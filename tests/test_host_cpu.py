"""Host mirror (reconstruction_b200/host) on CPU: it builds, parses the reference's config.yml / calib schema
(CManageData.cpp:26-66) and refuses to run without a GPU (no CPU fallback behind the C ABI)."""
import json
import os
import subprocess

import numpy as np
import pytest

from reconstruction_b200 import capi, stage

HOST = os.path.join(os.path.dirname(capi.HERE), "reconstruction_b200", "host")
BIN = os.path.join(HOST, "reconstruction")


@pytest.fixture(scope="module")
def cli():
    capi.build()
    subprocess.run(["make", "-s", "-C", HOST], check=True)
    return BIN


@pytest.mark.parametrize("block_style", [True, False])
def test_config_parse(cli, tmp_path, block_style):
    cfg, pairs = stage.write_dataset(str(tmp_path), 3, 32, 24, n_pairs=3, origin_scale=1.5, block_style=block_style)
    d = json.loads(subprocess.run([cli, "--dump-config", cfg], check=True, capture_output=True, text=True).stdout)
    assert (d["PyrmNum"], d["LowestLevelWidth"], d["LowestLevelHeight"]) == (3, 32, 24)
    assert (d["OriginWidth"], d["OriginHeight"]) == pairs[0].origin_size
    assert d["CameraNum"] == 4 and len(d["pairs"]) == 3 and d["isoutput"] == 0
    cams = stage.rig_cameras(4, *pairs[0].origin_size)
    for p, pr in enumerate(d["pairs"]):
        for k in (0, 1):
            c = pr[f"cam{k}"]
            assert c["id"] == p + k
            assert c["image"].endswith(f"0001_Cam{p + k}.ppm") and c["mask"].endswith(f"mask/0001_Cam{p + k}.ppm")
            K, Rt = cams[p + k]
            assert np.array_equal(np.array(c["K"]).reshape(3, 3), K)  # repr() round-trips doubles exactly
            assert np.array_equal(np.array(c["Rt"]).reshape(3, 4), Rt)
            assert np.allclose(c["center"], -Rt[:, :3].T @ Rt[:, 3], rtol=0, atol=1e-9)  # CManageData.cpp:61


def test_config_readable_by_opencv(tmp_path):
    """The staged files are the dialect cv::FileStorage reads (what the reference itself would parse)."""
    cv2 = pytest.importorskip("cv2")
    cfg, pairs = stage.write_dataset(str(tmp_path), 2, 32, 24, n_pairs=2)
    fs = cv2.FileStorage(cfg, cv2.FILE_STORAGE_READ)
    assert fs.isOpened()
    assert int(fs.getNode("PyrmNum").real()) == 2
    assert fs.getNode("camID").mat().tolist() == [[0, 1], [1, 2]]
    n = fs.getNode("imagelist")
    assert [n.at(i).string() for i in range(n.size())] == [f"0001_Cam{i}.ppm" for i in range(3)]
    st = cv2.FileStorage(os.path.join(str(tmp_path), "staged", "pair1.yml"), cv2.FILE_STORAGE_READ)
    assert np.array_equal(st.getNode("Q").mat(), pairs[1].Q)


def test_missing_config_and_no_gpu(cli, tmp_path):
    r = subprocess.run([cli, str(tmp_path / "nope.yml")], capture_output=True, text=True)
    assert r.returncode != 0 and "cannot open file" in r.stdout
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    cfg, _ = stage.write_dataset(str(tmp_path), 2, 32, 24)
    r = subprocess.run([cli, cfg], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU path" in r.stdout


def test_config_written_like_the_references_batch_tool(cli, tmp_path):
    """A data-set description produced the way BatchProcess/main.cpp:47-73 produces it — by cv::FileStorage, JPEG frames and JPEG
    masks, camID as an unsigned-char matrix, the reference's pair list {0,1},{2,3},{4,5},{7,6} (:30-36) — parses into the same
    CManageData fields; the original size comes from decoding masklist[0] (CManageData.cpp:68-69) with the native JPEG reader."""
    cv2 = pytest.importorskip("cv2")
    root = str(tmp_path) + os.sep
    os.makedirs(root + "mask")
    n_cam, ow, oh = 8, 72, 56
    rng = np.random.default_rng(3)
    for j in range(n_cam):
        assert cv2.imwrite(root + f"0001_Cam{j}.jpg", rng.integers(0, 256, (oh, ow, 3), dtype=np.uint8))
        assert cv2.imwrite(root + f"mask/0001_Cam{j}.jpg", np.full((oh, ow), 255, np.uint8))
    cams = stage.rig_cameras(n_cam, ow, oh)
    fs = cv2.FileStorage(root + "calib_camera.yml", cv2.FILE_STORAGE_WRITE)
    for j, (K, Rt) in enumerate(cams):
        fs.write(f"intrinsic-{j}", K)
        fs.write(f"extrinsic-{j}", Rt)
    fs.release()
    fs = cv2.FileStorage(root + "config.yml", cv2.FILE_STORAGE_WRITE)
    fs.write("filepath", root)
    fs.write("outfilename", root + "1.ply")
    fs.write("isoutput", 0)
    fs.write("camera_calib_name", "calib_camera.yml")
    fs.write("PyrmNum", 4)
    fs.write("LowestLevelWidth", 160)
    fs.write("LowestLevelHeight", 240)
    fs.write("imagelist", [f"0001_Cam{j}.jpg" for j in range(n_cam)])
    fs.write("masklist", [f"mask\\0001_Cam{j}.jpg" for j in range(n_cam)])  # "mask\\" + name, as BatchProcess/main.cpp:68 writes it
    fs.write("camID", np.array([[0, 1], [2, 3], [4, 5], [7, 6]], np.uint8))
    fs.release()
    d = json.loads(subprocess.run([cli, "--dump-config", root + "config.yml"], check=True, capture_output=True, text=True).stdout)
    assert (d["PyrmNum"], d["LowestLevelWidth"], d["LowestLevelHeight"], d["isoutput"]) == (4, 160, 240, 0)
    assert (d["OriginWidth"], d["OriginHeight"]) == (ow, oh) and d["CameraNum"] == n_cam and d["outfilename"] == root + "1.ply"
    assert [[p["cam0"]["id"], p["cam1"]["id"]] for p in d["pairs"]] == [[0, 1], [2, 3], [4, 5], [7, 6]]
    for p in d["pairs"]:
        for k in ("cam0", "cam1"):
            c = p[k]
            K, Rt = cams[c["id"]]
            assert c["image"] == root + f"0001_Cam{c['id']}.jpg" and c["mask"] == root + f"mask/0001_Cam{c['id']}.jpg"
            assert np.array_equal(np.array(c["K"]).reshape(3, 3), K) and np.array_equal(np.array(c["Rt"]).reshape(3, 4), Rt)


def _ply_body(path):
    raw = open(path, "rb").read()
    return raw[raw.index(b"end_header\n") + 11:]


def test_sink_handover_without_a_gpu(tmp_path):
    """The host mirror's sink around a stubbed sb200_sink_filter (tests/harness/sink_handover.cpp): records filtered ahead by
    worker threads and records filtered inside filter() land in pair order; tmp/cloud_filter.ply - rewritten per pair by a
    background thread that may drop superseded states - holds the LAST pair's records once run() returns; the merged files hold
    every pair."""
    import numpy as np

    exe = str(tmp_path / "sink_handover")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-o", exe, os.path.join(os.path.dirname(os.path.abspath(__file__)), "harness", "sink_handover.cpp"),
                    os.path.join(HOST, "CCloudOptimization.cpp"), os.path.join(HOST, "CManageData.cpp"), os.path.join(HOST, "sbcv.cpp"),
                    os.path.join(HOST, "sbimg.cpp"), "-lz", "-lpthread"], check=True)
    for counts in ([5, 0, 7, 300001, 4], [1000], [3, 3], [40000, 40001, 0]):
        d = tmp_path / ("run" + "_".join(map(str, counts)))
        d.mkdir()
        r = subprocess.run([exe, str(d), str(len(counts))] + [str(c) for c in counts], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        expect = []
        for p, n in enumerate(counts):
            i = np.arange(n)
            i = i[i % 3 != 0]
            dev = 7.0 if p % 2 else 0.0
            rec = np.stack([np.full(len(i), p), i, 0.25 * i + p, np.full(len(i), dev), i, np.full(len(i), n), np.ones(len(i))], 1).astype(np.float32)
            expect.append(rec)
        merged = np.frombuffer(_ply_body(d / "out.ply.normals.ply"), np.float32).reshape(-1, 7)
        assert np.array_equal(merged, np.concatenate(expect)), counts
        last = [e for e, n in zip(expect, counts) if n > 0][-1]  # filter() returns early for an empty pair (nothing is written for it)
        got = np.frombuffer(_ply_body(d / "tmp" / "cloud_filter.ply"), np.float32).reshape(-1, 7)
        assert np.array_equal(got, last), counts
        assert len(_ply_body(d / "out.ply")) == 15 * sum(counts)
        assert f"points {sum(counts)} kept {len(merged)}" in r.stdout, r.stdout
