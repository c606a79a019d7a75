"""Golden vectors for the native image readers (reconstruction_b200/host/sbimg.cpp): small JPEG / PNG / BMP files written by
OpenCV 4.13 (the build in this image: libjpeg-turbo, libpng) and what cv2.imdecode makes of them, colour and grey.
    python tests/golden/make_decode_golden.py      ->  tests/golden/decode/*.{jpg,png,bmp} + decode_expected.npz
The reference reads its frames with cv::imread (CStereoMatching.cpp:147-151); OpenCV 2.4.5's codecs are not under
/root/reference, so these vectors pin the restatement against the OpenCV that can be run here."""
import os

import cv2
import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "decode")


def texture(h, w, rng):
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([128 + 100 * np.sin(x / 7.0 + c) * np.cos(y / 5.0 - c) + rng.normal(0, 12, (h, w)) for c in range(3)], -1)
    return np.clip(img, 0, 255).astype(np.uint8)


def main():
    os.makedirs(HERE, exist_ok=True)
    rng = np.random.default_rng(20260101)
    files = {}
    S = {"420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, "444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444,
         "440": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440, "411": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411}
    for name, (h, w, q, sf, rst) in {"a_420_q90": (37, 53, 90, "420", 0), "b_422_q50_rst": (48, 40, 50, "422", 2), "c_444_q100": (17, 16, 100, "444", 0),
                                     "d_440_q75": (21, 9, 75, "440", 0), "e_411_q85_rst": (20, 70, 85, "411", 1), "f_420_tiny": (3, 2, 95, "420", 0)}.items():
        ok, b = cv2.imencode(".jpg", texture(h, w, rng), [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, S[sf], cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
        files[name + ".jpg"] = b.tobytes()
    ok, b = cv2.imencode(".jpg", texture(30, 41, rng)[..., 0], [cv2.IMWRITE_JPEG_QUALITY, 80])
    files["g_grey_q80.jpg"] = b.tobytes()
    ok, b = cv2.imencode(".jpg", texture(24, 24, rng), [cv2.IMWRITE_JPEG_OPTIMIZE, 1, cv2.IMWRITE_JPEG_QUALITY, 97])
    files["h_optimised_q97.jpg"] = b.tobytes()
    ok, b = cv2.imencode(".jpg", texture(29, 35, rng), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1, cv2.IMWRITE_JPEG_QUALITY, 88, cv2.IMWRITE_JPEG_RST_INTERVAL, 2])
    files["m_progressive_q88_rst.jpg"] = b.tobytes()
    ok, b = cv2.imencode(".png", texture(19, 23, rng))
    files["i_rgb.png"] = b.tobytes()
    ok, b = cv2.imencode(".png", rng.integers(0, 65536, (11, 13, 3)).astype(np.uint16))
    files["j_rgb16.png"] = b.tobytes()
    ok, b = cv2.imencode(".png", rng.integers(0, 256, (9, 10, 4)).astype(np.uint8))
    files["k_rgba.png"] = b.tobytes()
    ok, b = cv2.imencode(".bmp", texture(13, 7, rng))
    files["l.bmp"] = b.tobytes()
    exp = {}
    for name, data in files.items():
        open(os.path.join(HERE, name), "wb").write(data)
        arr = np.frombuffer(data, np.uint8)
        exp[name + ":color"] = cv2.imdecode(arr, cv2.IMREAD_COLOR)
        exp[name + ":gray"] = cv2.imdecode(arr, cv2.IMREAD_GRAYSCALE)
    np.savez_compressed(os.path.join(HERE, "decode_expected.npz"), **exp)
    print(len(files), "files,", sum(len(v) for v in files.values()), "bytes; OpenCV", cv2.__version__)


if __name__ == "__main__":
    main()
