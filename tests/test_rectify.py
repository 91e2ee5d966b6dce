"""Rectify / resize pre-stage (the reference's "AprilTag Graph" is rectify -> resize -> apriltag).
CPU part: the numpy restatement (oracle/rectify.py) against OpenCV's initUndistortRectifyMap + remap.  GPU part (-m gpu): the fused
kernel against the restatement, bit for bit, and the detections on raw distorted frames against the oracle's on the rectified ones."""
import os

import cv2
import numpy as np
import pytest

from oracle import rectify as R


def camera(w=960, h=720):
    K = np.array([[820.0, 0, w / 2 + 7.5], [0, 815.0, h / 2 - 4.25], [0, 0, 1]])
    D = np.array([-0.28, 0.09, 0.0012, -0.0007, -0.012])
    Rm = cv2.Rodrigues(np.array([0.01, -0.015, 0.004]))[0]
    return K, D, Rm


def distorted_scene(rng, w, h, K, D, Rm, P, out_w, out_h, tags):
    """A rectified scene with tags (ground truth known there), warped into the raw camera image by the inverse mapping."""
    from isaac_ros_apriltag_b200 import synth
    scene, truth = synth.make_frame(rng, out_w, out_h, tags, side_px=(70, 150), max_tilt_deg=25, K=P)
    # raw pixel -> rectified pixel: undistortPoints gives the ideal (rectified) position of every raw pixel
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float32)
    pts = np.stack([xs.ravel(), ys.ravel()], 1).reshape(-1, 1, 2)
    und = cv2.undistortPoints(pts, K, D, R=Rm, P=P).reshape(-1, 2)
    raw = cv2.remap(scene, und[:, 0].reshape(h, w), und[:, 1].reshape(h, w), cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=110)
    return raw, scene, truth


def test_map_and_remap_against_opencv():
    w, h = 960, 720
    K, D, Rm = camera(w, h)
    for out_w, out_h, scale in ((960, 720, 1.0), (640, 480, 2 / 3)):
        P = K.copy()
        P[:2] *= scale
        mx, my = R.rectify_map(K, D, Rm, P, out_w, out_h)
        cx, cy = cv2.initUndistortRectifyMap(K, D, Rm, P, (out_w, out_h), cv2.CV_32FC1)
        assert np.abs(mx - cx).max() < 2e-3 and np.abs(my - cy).max() < 2e-3  # (OpenCV accumulates the ray per column in floats)
        rng = np.random.default_rng(5)
        raw = cv2.GaussianBlur(rng.integers(0, 256, (h, w), dtype=np.uint8), (0, 0), 1.5)
        ours = R.rectify_gray(raw, "mono8", mx, my)
        ref = cv2.remap(raw, cx, cy, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        d = np.abs(ours.astype(int) - ref.astype(int))
        assert d.max() <= 2 and (d > 1).mean() < 1e-3, (d.max(), (d > 1).mean())
        bgr = np.repeat(raw[:, :, None], 3, axis=2)
        assert np.array_equal(R.rectify_gray(bgr, "bgr8", mx, my), ours)


def test_oracle_detects_on_rectified_frames():
    """End to end on the CPU side: tags drawn in the rectified scene, seen through the distorting camera, are found at their
    rectified positions after the pre-stage."""
    from oracle import oracle as O
    w, h = 960, 720
    K, D, Rm = camera(w, h)
    P = K.copy()
    rng = np.random.default_rng(2)
    raw, scene, truth = distorted_scene(rng, w, h, K, D, Rm, P, w, h, [("tag36h11", 3), ("tag36h11", 44)])
    mx, my = R.rectify_map(K, D, Rm, P, w, h)
    rect = R.rectify_gray(raw, "mono8", mx, my)
    dets = O.Oracle(("tag36h11",)).detect(rect)
    assert sorted(d["id"] for d in dets if d["hamming"] == 0) == [3, 44]
    for d in dets:
        tr = [t for t in truth if t["id"] == d["id"]][0]
        assert np.abs(d["p"] - tr["p"]).max() < 1.5  # (two interpolations between the drawn scene and the detector)
    # without the pre-stage the corners are off by the distortion
    draw = O.Oracle(("tag36h11",)).detect(raw)
    assert all(np.abs(d["p"] - [t for t in truth if t["id"] == d["id"]][0]["p"]).max() > 2 for d in draw if d["hamming"] == 0)


@pytest.mark.gpu
@pytest.mark.parametrize("enc,out", [("mono8", (960, 720)), ("bgr8", (960, 720)), ("rgba8", (640, 480))])
def test_fused_prestage_on_the_gpu(enc, out):
    import parity_util as pu
    if os.environ.get("B200AT_TEST_EMU") == "1":
        saved = pu.use_emulator()
    else:
        import torch
        if not torch.cuda.is_available():
            pytest.skip("no CUDA device")
        saved = None
    try:
        from isaac_ros_apriltag_b200 import capi
        from oracle import oracle as O
        w, h = 960, 720
        out_w, out_h = out
        K, D, Rm = camera(w, h)
        P = K.copy()
        P[:2] *= out_w / w
        rng = np.random.default_rng(7)
        raws, truths = [], []
        for _ in range(2):
            raw, scene, truth = distorted_scene(rng, w, h, K, D, Rm, P, out_w, out_h, [("tag36h11", 3), ("tag36h11", 44)])
            raws.append(raw)
            truths.append(truth)
        raws = np.stack(raws)
        ch = {"mono8": 1, "bgr8": 3, "rgba8": 4}[enc]
        frames = raws if ch == 1 else np.ascontiguousarray(np.repeat(raws[:, :, :, None], ch, axis=3))
        if ch > 1:
            frames[..., 0] = np.clip(frames[..., 0].astype(int) + 9, 0, 255)  # channels differ: the luma weights matter
        det = capi.Detector(out_w, out_h, intrinsics=(P[0, 0], P[1, 1], P[0, 2], P[1, 2]), tag_size=0.22, encoding=enc, max_batch=2, max_tags=16)
        det.set_rectification(w, h, K, D, Rm, P)
        t, ptrs, pitch = pu.upload(frames)
        gd = det.detect_device(ptrs, pitch, pu.current_stream())
        mx, my = R.rectify_map(K, D, Rm, P, out_w, out_h)
        orc = O.Oracle(("tag36h11",))
        for i in range(2):
            want = R.rectify_gray(frames[i], enc, mx, my)
            got = det.read_buffer(capi.BUF_RECTIFIED, i)
            assert np.array_equal(got, want), (i, int((got != want).sum()), int(np.abs(got.astype(int) - want.astype(int)).max()))
            od = orc.detect(want)
            assert [(d["id"], d["hamming"]) for d in od] == [(int(a), int(b)) for a, b in zip(gd[i]["id"], gd[i]["hamming"])]
            assert sorted(int(a) for a in gd[i]["id"]) == [3, 44]
            for a, b in zip(gd[i], od):
                assert np.abs(a["p"] - b["p"]).max() <= 1e-3
                tr = [x for x in truths[i] if x["id"] == b["id"]][0]
                assert np.abs(a["p"] - tr["p"]).max() < 1.5
        # the host-buffer entry point refuses while the pre-stage is on; switching it off restores the plain path
        with pytest.raises(capi.B200ATError):
            det.detect_host(frames)
        det.set_rectification(0, 0, None, None, None, None)
        if (out_w, out_h) == (w, h):
            plain = det.detect_device(ptrs, pitch, pu.current_stream())
            od = orc.detect(frames[0] if ch == 1 else O.to_gray(frames[0], enc))
            assert [d["id"] for d in od] == [int(a) for a in plain[0]["id"]]
        det.close()
    finally:
        if saved is not None:
            pu.restore(saved)
