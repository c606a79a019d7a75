"""The C++ host mirror end to end on the B200: `reconstruction config.yml` (reference CLI, main.cpp:5-24) on a staged
synthetic rig; the clouds it writes must be the oracle's points (f32-rounded as the PLY branch does,
CStereoMatching.cpp:753-756) in the reference order, with the source colours."""
import os
import subprocess

import numpy as np
import pytest

from reconstruction_b200 import capi, stage

pytestmark = pytest.mark.gpu
HOST = os.path.join(os.path.dirname(capi.HERE), "reconstruction_b200", "host")


def test_cli_matches_oracle(oracle, tmp_path):
    capi.build()
    subprocess.run(["make", "-s", "-C", HOST], check=True)
    L, w0, h0, n_pairs = 3, 64, 48, 3
    cfg, pairs = stage.write_dataset(str(tmp_path), L, w0, h0, n_pairs=n_pairs, isoutput=1)
    r = subprocess.run([os.path.join(HOST, "reconstruction"), cfg], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Matching time" in r.stdout and "total time" in r.stdout
    all_xyz = []
    for p, sp in enumerate(pairs):
        o = oracle.CpuStereo("port", L, w0, h0, *sp.origin_size)
        o.set_pair(*sp.image, *sp.mask)
        o.set_calib(sp.Q, sp.R_final, sp.T_final)
        n = o.match_pair()
        oxyz = o.to_cloud()
        obgr, opix = o.get_point_attrs()
        xyz, bgr = stage.read_ply_f32(str(tmp_path / f"cloud{p}.ply"))
        assert len(xyz) == n
        assert np.array_equal(xyz.view(np.int32), oxyz.astype(np.float32).view(np.int32)), f"pair {p}: points differ"
        assert np.array_equal(bgr, obgr)
        all_xyz.append(oxyz.astype(np.float32))
    xyz, _ = stage.read_ply_f32(str(tmp_path / "out.ply"))  # the sink's merged cloud, pair order
    assert np.array_equal(xyz.view(np.int32), np.concatenate(all_xyz).view(np.int32))


def test_cli_native_rectify(tmp_path, golden_dir):
    """`reconstruction config.yml` on ORIGINAL frames: Rectify runs natively (host calibration + device remap/erode); the cloud
    must equal the one the same C ABI calls produce when driven from Python."""
    capi.build()
    subprocess.run(["make", "-s", "-C", HOST], check=True)
    g = np.load(os.path.join(golden_dir, "rectify_cv2.npz"))
    c = "a"
    L, (w0, h0) = int(g[c + "_pyrm_num"]), (int(v) for v in g[c + "_lowest"])
    origin = tuple(int(v) for v in g[c + "_origin"])
    cams = [(g[c + "_K0"], g[c + "_Rt0"]), (g[c + "_K1"], g[c + "_Rt1"])]
    cfg = stage.write_raw_dataset(str(tmp_path), L, w0, h0, origin, cams, [g[c + "_src_image0"], g[c + "_src_image1"]],
                                  [g[c + "_src_mask0"], g[c + "_src_mask1"]], [[0, 1]], isoutput=1)
    r = subprocess.run([os.path.join(HOST, "reconstruction"), cfg], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    xyz, bgr = stage.read_ply_f32(str(tmp_path / "cloud0.ply"))
    st = capi.StereoB200(L, w0, h0, *origin)
    cal = capi.rectify_calib(cams[0][0], cams[0][1], cams[1][0], cams[1][1], origin, w0, L)
    for j in (0, 1):
        st.rectify_view(j, g[c + f"_src_image{j}"], g[c + f"_src_mask{j}"], cams[j][0], cal["R_new"][j], cal["P_scaled"][j])
    st.pair_build()
    st.set_calib(cal["Q"], cal["R_final"], cal["T_final"])
    n = st.match_pair()
    pxyz, pbgr, _ = st.get_points(n)
    assert len(xyz) == n and n > 0
    assert np.array_equal(xyz.view(np.int32), pxyz.astype(np.float32).view(np.int32))
    assert np.array_equal(bgr, pbgr)


@pytest.mark.parametrize("fmt", ["jpg", "png"])
def test_cli_reads_the_references_image_formats(tmp_path, golden_dir, fmt):
    """The reference's data sets are JPEG frames and JPEG masks read with cv::imread (BatchProcess/main.cpp:66,
    CStereoMatching.cpp:147-151).  `reconstruction config.yml` on such a data set decodes them natively (sbimg.cpp); the cloud must
    equal, bit for bit, the one obtained when OpenCV decodes the same files and the same C ABI calls are driven from Python."""
    cv2 = pytest.importorskip("cv2")
    capi.build()
    subprocess.run(["make", "-s", "-C", HOST], check=True)
    g = np.load(os.path.join(golden_dir, "rectify_cv2.npz"))
    c = "a"
    L, (w0, h0) = int(g[c + "_pyrm_num"]), (int(v) for v in g[c + "_lowest"])
    origin = tuple(int(v) for v in g[c + "_origin"])
    cams = [(g[c + "_K0"], g[c + "_Rt0"]), (g[c + "_K1"], g[c + "_Rt1"])]
    cfg = stage.write_raw_dataset(str(tmp_path), L, w0, h0, origin, cams, [g[c + "_src_image0"], g[c + "_src_image1"]],
                                  [g[c + "_src_mask0"], g[c + "_src_mask1"]], [[0, 1]], isoutput=1, fmt=fmt)
    r = subprocess.run([os.path.join(HOST, "reconstruction"), cfg], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    xyz, bgr = stage.read_ply_f32(str(tmp_path / "cloud0.ply"))
    st = capi.StereoB200(L, w0, h0, *origin)
    cal = capi.rectify_calib(cams[0][0], cams[0][1], cams[1][0], cams[1][1], origin, w0, L)
    for j in (0, 1):
        img = cv2.imread(str(tmp_path / f"0001_Cam{j}.{fmt}"), cv2.IMREAD_COLOR)
        msk = cv2.imread(str(tmp_path / "mask" / f"0001_Cam{j}.{fmt}"), cv2.IMREAD_GRAYSCALE)
        st.rectify_view(j, img, msk, cams[j][0], cal["R_new"][j], cal["P_scaled"][j])
    st.pair_build()
    st.set_calib(cal["Q"], cal["R_final"], cal["T_final"])
    n = st.match_pair()
    pxyz, pbgr, _ = st.get_points(n)
    assert len(xyz) == n and n > 0
    assert np.array_equal(xyz.view(np.int32), pxyz.astype(np.float32).view(np.int32))
    assert np.array_equal(bgr, pbgr)


def test_cli_gather_path_on_one_gpu(tmp_path):
    """SB200_ALLGATHER=2 takes the several-device path with a single device (a one-rank communicator, the workers' points kept in
    HBM, the exchange thread, the sink filter running ahead on its own host thread): raw clouds, merged cloud, oriented cloud and
    tmp/cloud_filter.ply must equal the default path's, byte for byte."""
    capi.build()
    subprocess.run(["make", "-s", "-C", HOST], check=True)
    L, w0, h0, n_pairs = 3, 64, 48, 5
    outs = {}
    for tag, env in (("plain", {"SB200_DEVICES": "0", "SB200_ALLGATHER": "0"}), ("gather", {"SB200_DEVICES": "0", "SB200_ALLGATHER": "2", "SB200_CTX_PER_DEVICE": "2"}),
                     ("gather1", {"SB200_DEVICES": "0", "SB200_ALLGATHER": "2", "SB200_CTX_PER_DEVICE": "1"})):
        d = tmp_path / tag
        d.mkdir()
        cfg, _ = stage.write_dataset(str(d), L, w0, h0, n_pairs=n_pairs, isoutput=1)
        r = subprocess.run([os.path.join(HOST, "reconstruction"), cfg], cwd=str(d), capture_output=True, text=True, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stdout + r.stderr
        assert ("point all-gather:" in r.stdout) == (tag != "plain"), r.stdout
        files = [f"cloud{p}.ply" for p in range(n_pairs)] + ["out.ply", "out.ply.normals.ply", os.path.join("tmp", "cloud_filter.ply")]
        outs[tag] = {f: open(d / f, "rb").read() for f in files}
        assert all(len(v) > 200 for v in outs[tag].values())
    def records(raw):  # PointNormal PLY: 7 float32 per vertex after the header
        body = raw[raw.index(b"end_header\n") + 11:]
        return np.frombuffer(body, np.float32).reshape(-1, 7)

    for f in outs["plain"]:
        for other in ("gather", "gather1"):
            if f.startswith("cloud") or f == "out.ply":
                assert outs["plain"][f] == outs[other][f], (f, other)
            else:  # kept points identical; normals are eigenvectors (same code, same inputs - compared numerically all the same)
                a, b = records(outs["plain"][f]), records(outs[other][f])
                assert a.shape == b.shape and len(a) > 0 and np.array_equal(a[:, :3].view(np.int32), b[:, :3].view(np.int32)), (f, other)
                assert np.allclose(a[:, 3:], b[:, 3:], rtol=1e-4, atol=1e-5, equal_nan=True), (f, other)


def test_cli_two_gpus_gathers_the_points_over_nccl(tmp_path):
    """Several devices: the workers keep their points in HBM and the C ABI's NCCL all-gather (sb200_exchange_*) collects them;
    the merged cloud must equal the single-device run's, byte for byte (pair order, CloudOptimization/CCloudOptimization.cpp:123)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    capi.build()
    subprocess.run(["make", "-s", "-C", HOST], check=True)
    L, w0, h0, n_pairs = 3, 64, 48, 5  # an odd number of pairs: the last ticket has a device without a pair
    outs = {}
    for tag, env in (("one", {"SB200_DEVICES": "0"}), ("two", {"SB200_DEVICES": "0,1", "SB200_CTX_PER_DEVICE": "2"})):
        d = tmp_path / tag
        d.mkdir()
        cfg, _ = stage.write_dataset(str(d), L, w0, h0, n_pairs=n_pairs, isoutput=1)
        r = subprocess.run([os.path.join(HOST, "reconstruction"), cfg], cwd=str(d), capture_output=True, text=True, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stdout + r.stderr
        if tag == "two":
            assert "point all-gather:" in r.stdout, r.stdout
        outs[tag] = [open(d / f"cloud{p}.ply", "rb").read() for p in range(n_pairs)] + [open(d / "out.ply", "rb").read()]
    assert outs["one"] == outs["two"]
