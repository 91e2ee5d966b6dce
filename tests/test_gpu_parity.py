"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same seeded
frames.  Bar: bit-exact for every integer/byte/index stage (decimated image, tile min/max, threshold, labels, component
sizes, cluster keys and point sets, sort order), tag IDs and Hamming distance exact; float corners of the quad fit
bit-exact in the exact mode (B200AT_TUNE=qf_exact=1) and within TOL_QUAD_PX in the default windowed mode; refined corners /
homography / decision margin / pose within the tolerances stated below (refine_edges' atan2f/cosf/sinf differ in the last
bit between glibc and CUDA)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_REFINED_PX = 5e-2   # refined corners of ALL quads incl. undecodable junk quads whose edge fits are ill-conditioned (atan2f of ~0/~0: a last-bit difference of the angle moves such a corner by up to ~0.02 px)
TOL_CORNER_PX = 1e-3    # detection corners / centre
TOL_MARGIN = 5e-3       # decision margin (gray levels, typically 40-100): a 1e-4 px corner difference moves a bilinear sample on a tag edge by ~0.02
TOL_POSE_T = 1e-5       # metres (relative to tag distance ~1-5 m)
TOL_POSE_R = 1e-5


@pytest.fixture(scope="module")
def pu():
    """B200AT_TEST_EMU=1 (developer switch, no GPU needed) runs this same suite against the kernels compiled for the CPU
    SIMT emulator of tools/emu instead of the GPU: a logic check before spending GPU time, never a substitute for -m gpu."""
    import parity_util
    if os.environ.get("B200AT_TEST_EMU") == "1":
        saved = parity_util.use_emulator()
        yield parity_util
        parity_util.restore(saved)
        return
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    yield parity_util


def needs_real_gpu(pu):
    if pu.EMU:
        pytest.skip("needs a real GPU (torch device tensors / CUDA graphs)")


TOL_QUAD_PX = 1e-3      # windowed quad fit: float quad corners vs the oracle's (bit-identical in the exact mode)
MAX_QUAD_SET_DIFF = 0.005  # ... and the fraction of candidate quads that may be accepted on one side only (ties within rounding)


def exact_mode():
    return "qf_exact=1" in os.environ.get("B200AT_TUNE", "")


def assert_exact(res, rep):
    """Bit-exact for every integer / index stage and for tag id, Hamming distance, family.  Quad corners: bit-identical to the
    oracle in the exact quad-fit mode (B200AT_TUNE=qf_exact=1); in the default windowed mode (prefix sums associated differently,
    see k_quad2.cu) within TOL_QUAD_PX, the set of accepted candidate quads equal up to MAX_QUAD_SET_DIFF."""
    assert res["status"] == 0, res
    for k in ("dec", "tile", "thr", "labels", "sizes", "clusters", "points", "det_n", "det_id"):
        assert res[k] == 0, (k, res, rep)
    assert res.get("order", 0) == 0, res
    if exact_mode():
        assert res["quads_n"] == 0 and res["quads_bits"] == 0 and res["refined_n"] == 0, res
    else:
        allowed = int(MAX_QUAD_SET_DIFF * res["n_quads"])
        assert res["quads_n"] <= allowed and res["refined_n"] <= allowed and res["quads_max"] <= TOL_QUAD_PX, res
    assert res["refined_max"] <= TOL_REFINED_PX and res["det_corner_max"] <= TOL_CORNER_PX and res["det_margin_max"] <= TOL_MARGIN, res


def _as_encoding(frames, enc):
    if enc == "mono8":
        return frames
    ch = 3 if enc in ("rgb8", "bgr8") else 4
    col = np.repeat(frames[:, :, :, None], ch, axis=3)
    if ch == 4:
        col[..., 3] = 255
    return np.ascontiguousarray(col)


@pytest.mark.parametrize("config,n,enc", [("C1", 2, "mono8"), ("C2", 2, "mono8"), ("C2", 2, "bgr8"), ("C3", 8, "mono8"), ("C4", 1, "mono8"),
                                           ("C5", 2, "mono8"), ("C5", 2, "bgr8")])
def test_stage_parity_configs(pu, config, n, enc):
    """Every BASELINE.json config, stage by stage against the oracle (C3 = 4K with eight frames; C2 / C5 also in the node's bgr8 wire
    format)."""
    from isaac_ros_apriltag_b200 import synth
    if pu.EMU and config == "C3":
        n = 1  # (the emulator runs a 4K frame in ~20 s)
    frames, truths, K, ts, fams = synth.make_config_frames(config, n)
    rep = []
    res, gdets = pu.compare_stages(_as_encoding(frames, enc), enc, fams, report=rep)
    assert_exact(res, rep)
    assert res["n_det"] >= n  # the frames do contain detectable tags
    # and the detections are the ground truth tags
    for g, tr in zip(gdets, truths):
        want = sorted((t["family"], t["id"]) for t in tr)
        got = sorted((["tag36h11", "tag25h9", "tag16h5", "tag36h10"][d["family"]], int(d["id"])) for d in g if d["hamming"] == 0)
        if config != "C4":
            assert got == want
        else:
            assert len(set(got) & set(want)) >= 0.9 * len(want)


@pytest.mark.parametrize("tune", ["qf_exact=1", "ccl_tma=0", "ccl_tma=0,qf_exact=1", "qf_bucket_limit=3"])
def test_every_tune_variant_on_the_gpu(pu, tune, monkeypatch):
    """The kernel variants behind B200AT_TUNE (csrc/detector.h, struct Tune: the bit-exact quad fit, the CCL sweep without TMA staging,
    the quad-fit sort's fallback for degenerate outlines -- qf_bucket_limit=3 sends nearly every cluster through it) run on the GPU
    against the oracle like the defaults do -- a variant that only ever ran under the emulator is inventory."""
    from isaac_ros_apriltag_b200 import synth
    monkeypatch.setenv("B200AT_TUNE", tune)
    frames, truths, K, ts, fams = synth.make_config_frames("C2", 2)
    big = synth.make_frame(np.random.default_rng(11), 1600, 1200, [("tag36h11", 42)], side_px=(680, 720), max_tilt_deg=10.0, noise_sigma=0.0)[0]
    for fr, fam in ((_as_encoding(frames, "bgr8"), fams), (big[None], ("tag36h11",))):
        rep = []
        res, gd = pu.compare_stages(fr, "bgr8" if fr.ndim == 4 else "mono8", fam, report=rep)
        assert_exact(res, rep)
        assert res["n_det"] >= 1
    if "qf_exact=1" in tune:
        assert res["quads_bits"] == 0


@pytest.mark.parametrize("family,ids", [("tag16h5", [0, 7, 29]), ("tag36h10", [0, 1000, 2319]), ("tag25h9", [0, 34])])
def test_other_families_decode_on_the_gpu(pu, family, ids):
    """The families beyond tag36h11 that have a code table (apriltag_node.cpp:47-58): rendered, detected on the GPU, compared with the
    oracle stage by stage; id / Hamming distance / family exact.  (tag16h5 is known for false positives on texture: both sides must
    report the SAME ones.)"""
    from isaac_ros_apriltag_b200 import synth
    rng = np.random.default_rng(len(family))
    g, truth = synth.make_frame(rng, 960, 720, [(family, i) for i in ids], side_px=(90, 170), max_tilt_deg=30.0)
    rep = []
    res, gd = pu.compare_stages(np.stack([g, g[:, ::-1].copy()]), "mono8", (family,), report=rep)
    assert_exact(res, rep)
    got = sorted(int(d["id"]) for d in gd[0] if d["hamming"] == 0)
    assert set(ids) <= set(got), (ids, got)
    # and with every family registered at once (multi-family decode path, like config C5)
    rep = []
    res, gd = pu.compare_stages(g[None], "mono8", ("tag36h11", "tag25h9", "tag16h5", "tag36h10"), report=rep)
    assert_exact(res, rep)
    fi = ["tag36h11", "tag25h9", "tag16h5", "tag36h10"].index(family)
    assert set(ids) <= set(int(d["id"]) for d in gd[0] if d["family"] == fi and d["hamming"] == 0)


@pytest.mark.parametrize("encoding", ["bgr8", "rgb8", "rgba8", "bgra8"])
def test_colour_encodings(pu, encoding):
    from isaac_ros_apriltag_b200 import synth
    rng = np.random.default_rng(21)
    frames = []
    for _ in range(2):
        g, _ = synth.make_frame(rng, 640, 480, [("tag36h11", 5), ("tag36h11", 9)], side_px=(60, 140))
        ch = 3 if encoding in ("rgb8", "bgr8") else 4
        col = rng.integers(0, 40, g.shape + (ch,), dtype=np.int16) + g[:, :, None].astype(np.int16) - 20
        frames.append(np.clip(col, 0, 255).astype(np.uint8))
    rep = []
    res, _ = pu.compare_stages(np.stack(frames), encoding, ("tag36h11",), report=rep)
    assert_exact(res, rep)


@pytest.mark.parametrize("w,h", [(1226, 370), (751, 481), (322, 242)])
def test_odd_sizes_partial_tiles(pu, w, h):
    """Widths/heights that are not multiples of the tile size or of 16: partial-tile threshold path, padded row pitch,
    unaligned (generic) load path."""
    from isaac_ros_apriltag_b200 import synth
    rng = np.random.default_rng(w)
    g, _ = synth.make_frame(rng, w, h, [("tag36h11", 17)], side_px=(60, 120))
    rep = []
    res, _ = pu.compare_stages(g[None], "mono8", ("tag36h11",), report=rep)
    assert_exact(res, rep)


@pytest.mark.parametrize("opts", [dict(quad_decimate=1.0), dict(quad_decimate=1.5), dict(quad_decimate=3.0), dict(quad_sigma=0.8), dict(quad_sigma=-0.8),
                                  dict(tile_size=8), dict(refine_edges=0), dict(decode_sharpening=0.0), dict(max_hamming=1),
                                  dict(min_white_black_diff=20)])
def test_detector_knobs(pu, opts):
    from isaac_ros_apriltag_b200 import synth
    rng = np.random.default_rng(5)
    g, _ = synth.make_frame(rng, 800, 600, [("tag36h11", 1), ("tag36h11", 2), ("tag36h11", 3)], side_px=(70, 160))
    rep = []
    res, _ = pu.compare_stages(np.stack([g, g[::-1].copy()]), "mono8", ("tag36h11",), report=rep, **opts)
    assert_exact(res, rep)


def test_empty_blank_and_noise_frames(pu):
    rng = np.random.default_rng(1)
    frames = np.stack([np.full((480, 640), 128, np.uint8), np.zeros((480, 640), np.uint8),
                       rng.integers(0, 256, (480, 640), dtype=np.uint8), np.full((480, 640), 255, np.uint8)])
    rep = []
    res, gd = pu.compare_stages(frames, "mono8", ("tag36h11",), report=rep)
    assert_exact(res, rep)
    assert len(gd[0]) == 0 and len(gd[1]) == 0 and len(gd[3]) == 0


def test_pose_against_oracle_and_truth(pu):
    import torch
    from isaac_ros_apriltag_b200 import capi, synth
    from oracle import oracle as O
    frames, truths, K, ts, fams = synth.make_config_frames("C2", 2)
    H, W = frames.shape[1:]
    det = capi.Detector(W, H, intrinsics=(K[0, 0], K[1, 1], K[0, 2], K[1, 2]), tag_size=ts, families=fams, encoding="mono8",
                        max_batch=2, max_tags=64)
    t, ptrs, pitch = pu.upload(frames)
    gd = det.detect_device(ptrs, pitch, pu.current_stream())
    orc = O.Oracle(fams)
    fx, fy, cx, cy = np.float32(K[0, 0]), np.float32(K[1, 1]), np.float32(K[0, 2]), np.float32(K[1, 2])
    for i in range(2):
        od = orc.detect(frames[i])
        assert [d["id"] for d in od] == list(gd[i]["id"])
        for k, d in enumerate(gd[i]):
            best, p1, p2 = orc.estimate_pose(k, float(fx), float(fy), float(cx), float(cy), float(np.float32(ts)))
            assert np.abs(d["t"] - best["t"]).max() <= TOL_POSE_T * max(1.0, best["t"][2])
            assert np.abs(d["R"].reshape(3, 3) - best["R"]).max() <= TOL_POSE_R
            tr = [x for x in truths[i] if x["id"] == d["id"]][0]
            assert np.abs(d["t"] - tr["t"]).max() < 0.05 * tr["t"][2]
    det.close()


def test_pol_golden_through_cuapriltags_abi(pu):
    """The reference's POL golden test (isaac_ros_apriltag_pol_test.py:117-175) through the drop-in entry points
    nvCreateAprilTagsDetector / cuAprilTagsDetect and through the node core, with the reference's own tolerances."""
    needs_real_gpu(pu)
    import ctypes as C
    import torch
    from isaac_ros_apriltag_b200 import capi, node, synth
    exp = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "apriltag0_expected.json")))
    bgr, K = synth.make_apriltag0()
    t = torch.from_numpy(bgr).cuda()
    n = node.AprilTagNode(tag_family="tag36h11", backends="CUDA", size=0.22, max_tags=64, tile_size=4)
    for _ in range(2):  # second frame exercises the already-initialised path
        dets = n.on_frame("bgr8", 1920, 1080, 1920 * 3, t.data_ptr(), K)
    assert len(dets) >= 1
    for d in dets:
        assert d["id"] == 0 and d["family"] == "tag36h11" and d["child_frame_id"] == "tag36h11:0"
        assert np.abs(d["center"] - np.array(exp["center"])).max() <= 2
        assert np.abs(d["corners"] - np.array(exp["corners"])).max() <= 2
        assert np.abs(d["position"] - np.array(exp["translation"])).max() <= 0.01
        x, y, z, w = d["orientation_xyzw"]
        q = np.array([w, x, y, z])
        ew = np.array(exp["quaternion_wxyz"])
        assert min(np.abs(q - ew).max(), np.abs(q + ew).max()) <= 0.01
    # unsupported encoding on the cuAprilTags strategy throws like apriltag_node.cpp:469-476
    with pytest.raises(RuntimeError, match="only supports 'rgb8' or 'bgr8'"):
        n.on_frame("mono8", 1920, 1080, 1920, t.data_ptr(), K)
    n.close()
    # mono8 through the other strategy (isaac_ros_apriltag_mono8_test.py:105-141: >= 1 detection)
    gray = torch.from_numpy(np.ascontiguousarray(bgr[:, :, 0])).cuda()
    n2 = node.AprilTagNode(tag_family="tag36h11", backends="CPU")
    d2 = n2.on_frame("mono8", 1920, 1080, 1920, gray.data_ptr(), K)
    assert len(d2) >= 1 and d2[0]["id"] == 0
    # backends-compare test (isaac_ros_apriltag_backends_compare_test.py:154-249): index-wise agreement of the two strategies
    assert len(d2) == len(dets)
    for a, b in zip(dets, d2):
        assert a["id"] == b["id"] and np.abs(a["center"] - b["center"]).max() <= 2 and np.abs(a["corners"] - b["corners"]).max() <= 2
        assert np.abs(a["position"] - b["position"]).max() <= 0.01
        assert min(np.abs(a["orientation_xyzw"] - b["orientation_xyzw"]).max(), np.abs(a["orientation_xyzw"] + b["orientation_xyzw"]).max()) <= 0.01
    n2.close()


def test_registered_reversed_border_family(pu):
    """b200AprilTagsRegisterFamily with a synthetic REVERSED-BORDER family (tag36h11's codes and bit layout, white border ring on a
    black surround -- the polarity of upstream's tagStandard* / tagCircle* / tagCustom* families, whose tables are not built in):
    exercises fit_quad's polarity gate, refine_edges' flipped normals, quad_decode's polarity check and, with a normal family
    registered beside it, the per-family polarity filter (apriltag.c: `if (family->reversed_border != quad->reversed_border)`)."""
    from isaac_ros_apriltag_b200 import capi, families as F, synth
    from oracle import oracle as O
    fam = dict(F.families()["tag36h11"])
    fam["reversed_border"] = True
    F.add_family("custom0", fam)
    capi.register_family(4, fam)
    O.register_family(4, fam)
    rng = np.random.default_rng(9)
    g, truth = synth.make_frame(rng, 960, 720, [("custom0", 5), ("custom0", 77), ("tag36h11", 5)], side_px=(100, 170), max_tilt_deg=25)
    for fams, want in ((("custom0",), [(4, 5), (4, 77)]), (("tag36h11", "custom0"), [(0, 5), (4, 5), (4, 77)]), (("tag36h11",), [(0, 5)])):
        rep = []
        res, gd = pu.compare_stages(np.stack([g, g[::-1].copy()]), "mono8", fams, report=rep)
        assert_exact(res, rep)
        assert sorted((int(d["family"]), int(d["id"])) for d in gd[0] if d["hamming"] == 0) == want, fams
    # argument validation of the registration entry point
    with pytest.raises(capi.B200ATError):
        capi.register_family(2, fam)          # built-in slots cannot be overwritten
    bad = dict(fam)
    bad["bit_x"] = [9] * fam["nbits"]          # outside the tag
    with pytest.raises(capi.B200ATError):
        capi.register_family(5, bad)
    with pytest.raises(capi.B200ATError):
        capi.Detector(640, 480, families=("custom1",))  # an unregistered slot is UNSUPPORTED at create


def test_node_loads_a_family_table_file(pu, tmp_path, monkeypatch):
    """The node's non-CUDA strategy accepts the nine family names of apriltag_node.cpp:47-58; for the five without a built-in table it
    reads $B200AT_FAMILY_PATH/<family>.txt (the fields of upstream's tag<Family>.c).  Mechanism test with a synthetic reversed-border
    table under the name custom48h12: detections are published with that family string and tf child frame."""
    needs_real_gpu(pu)
    import torch
    from isaac_ros_apriltag_b200 import families as F, node, synth
    fam = dict(F.families()["tag36h11"])
    fam["reversed_border"] = True
    F.add_family("custom0", fam)
    with open(tmp_path / "custom48h12.txt", "w") as f:
        f.write(f"{fam['nbits']} {len(fam['codes'])} {fam['width_at_border']} {fam['total_width']} 1\n")
        f.write(" ".join(str(v) for v in fam["bit_x"]) + "\n" + " ".join(str(v) for v in fam["bit_y"]) + "\n")
        f.write("\n".join(hex(c) for c in fam["codes"]) + "\n")
    g, truth = synth.make_frame(np.random.default_rng(9), 960, 720, [("custom0", 5), ("custom0", 77)], side_px=(100, 170), max_tilt_deg=25)
    t = torch.from_numpy(g).cuda()
    K = synth.default_K(960, 720)
    n = node.AprilTagNode(tag_family="custom48h12", backends="CPU")
    with pytest.raises(RuntimeError, match="no code table for family"):
        n.on_frame("mono8", 960, 720, 960, t.data_ptr(), K)   # no B200AT_FAMILY_PATH: the create error of the first frame
    n.close()
    monkeypatch.setenv("B200AT_FAMILY_PATH", str(tmp_path))
    n = node.AprilTagNode(tag_family="custom48h12", backends="CPU")
    dets = n.on_frame("mono8", 960, 720, 960, t.data_ptr(), K)
    assert sorted(d["id"] for d in dets) == [5, 77] and all(d["family"] == "custom48h12" for d in dets)
    assert sorted(d["child_frame_id"] for d in dets) == ["custom48h12:5", "custom48h12:77"]
    n.close()


def test_multi_tag_frame_through_cuapriltags_abi_and_node(pu):
    """Ten tags in one frame through cuAprilTagsDetect (array of cuAprilTagsID_t: 88-byte stride, CUDA's 8-byte aligned float2) and
    through both node strategies: every entry, not only the first, carries the id / corners / pose of the batch entry point."""
    needs_real_gpu(pu)
    import ctypes as C
    import torch
    from isaac_ros_apriltag_b200 import capi, node, synth
    frames, truths, K, ts, fams = synth.make_config_frames("C2", 1)
    bgr = _as_encoding(frames, "bgr8")
    H, W = frames.shape[1:]
    t = torch.from_numpy(bgr).cuda()
    det = capi.Detector(W, H, intrinsics=(K[0, 0], K[1, 1], K[0, 2], K[1, 2]), tag_size=ts, families=fams, encoding="bgr8", max_batch=1, max_tags=64)
    want = det.detect_device([t.data_ptr()], W * 3, 0)[0]
    assert len(want) == 10
    L = capi.lib()
    hdl = C.c_void_p()
    cam = capi.Intrinsics(float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]))
    assert L.nvCreateAprilTagsDetector(C.byref(hdl), W, H, 4, 0, C.byref(cam), C.c_float(ts)) == 0
    img = capi.ImageInput(t.data_ptr(), W * 3, W, H)
    tags = np.zeros(64, capi.ID_DTYPE)
    ntags = C.c_uint32()
    for stream in (0, torch.cuda.Stream().cuda_stream):  # legacy default stream and a real one (graph replay from the third call on)
        for _ in range(3):
            tags[:] = 0
            rc = L.cuAprilTagsDetect(hdl, C.byref(img), tags.ctypes.data_as(C.POINTER(capi.TagID)), C.byref(ntags), 64, C.c_void_p(stream))
            assert rc == 0 and ntags.value == 10
            for k in range(10):
                assert int(tags[k]["id"]) == int(want[k]["id"]) and int(tags[k]["hamming_error"]) == int(want[k]["hamming"])
                assert np.abs(tags[k]["corners"] - want[k]["p"][::-1].astype(np.float32)).max() < 1e-3
                assert np.abs(tags[k]["translation"] - want[k]["t"]).max() < 1e-5
                assert np.abs(tags[k]["orientation"].reshape(3, 3).T - want[k]["R"].reshape(3, 3)).max() < 1e-5
    # a caller's max_tags smaller than the scene: a truncated list and success, as a full cuAprilTags list would be
    rc = L.cuAprilTagsDetect(hdl, C.byref(img), tags.ctypes.data_as(C.POINTER(capi.TagID)), C.byref(ntags), 4, None)
    assert rc == 0 and ntags.value == 4
    L.cuAprilTagsDestroy(hdl)
    for backends in ("CUDA", "CPU"):
        n = node.AprilTagNode(tag_family="tag36h11", backends=backends, size=ts, max_tags=64, tile_size=4)
        for _ in range(3):
            dets = n.on_frame("bgr8", W, H, W * 3, t.data_ptr(), K)
        assert [d["id"] for d in dets] == [int(i) for i in want["id"]]
        for d, w_ in zip(dets, want):
            assert np.abs(d["corners"] - w_["p"][::-1]).max() < 1e-2 and np.abs(d["center"] - w_["c"]).max() < 0.6
            assert np.abs(d["position"] - w_["t"]).max() < 1e-4
        n.close()
    det.close()


def test_full_batch_properties(pu):
    """BASELINE size (batch 256 @1080p would take the oracle minutes): size-independent properties instead.
    Every replica of the same frame inside one batch gives the byte-identical result (idempotence / no cross-frame
    leakage), the host-buffer entry point agrees with the device-pointer one, and the first distinct frames match the oracle."""
    import torch
    from isaac_ros_apriltag_b200 import capi, synth
    from oracle import oracle as O
    distinct, B = 4, 64
    frames, truths, K, ts, fams = synth.make_config_frames("C2", distinct)
    batch = np.ascontiguousarray(frames[np.arange(B) % distinct])
    H, W = frames.shape[1:]
    det = capi.Detector(W, H, intrinsics=(K[0, 0], K[1, 1], K[0, 2], K[1, 2]), tag_size=ts, families=fams, encoding="mono8",
                        max_batch=B, max_tags=64)
    t, ptrs, pitch = pu.upload(batch)
    gd = det.detect_device(ptrs, pitch, pu.current_stream())
    assert det.status() == 0
    for i in range(B):
        assert gd[i].tobytes() == gd[i % distinct].tobytes(), i
    hd = det.detect_host(batch)
    for i in range(B):
        assert hd[i].tobytes() == gd[i].tobytes()
    orc = O.Oracle(fams)
    for i in range(distinct):
        od = orc.detect(frames[i])
        assert [(d["id"], d["hamming"]) for d in od] == [(int(a), int(b)) for a, b in zip(gd[i]["id"], gd[i]["hamming"])]
    # smaller batch through a larger workspace and chunking through a smaller one
    det2 = capi.Detector(W, H, families=fams, encoding="mono8", max_batch=3, max_tags=64)
    hd2 = det2.detect_host(batch[:7])
    for i in range(7):
        assert list(hd2[i]["id"]) == list(gd[i]["id"])
    det.close()
    det2.close()


def test_overflow_is_reported_not_ub(pu):
    import torch
    from isaac_ros_apriltag_b200 import capi, synth
    frames, *_ = synth.make_config_frames("C1", 1)
    H, W = frames.shape[1:]
    det = capi.Detector(W, H, encoding="mono8", max_batch=1, hash_slots_per_frame=1024, points_per_frame=4096, clusters_per_frame=64)
    t, ptrs, pitch = pu.upload(frames)
    with pytest.raises(capi.B200ATError) as e:
        det.detect_device(ptrs, pitch, pu.current_stream())
    assert e.value.code == 5 and det.status() != 0
    det.close()


def test_cuda_graph_replay_matches_plain_launches(pu):
    """EnqueueBatch caches a CUDA graph of the whole batch when the caller passes a real stream: call 1 runs plain,
    call 2 captures + launches, later calls replay.  Results must be identical, and a replay must pick up NEW frame
    pointers (the frame table is a pinned buffer read at execution time)."""
    needs_real_gpu(pu)
    import torch
    from isaac_ros_apriltag_b200 import capi, synth
    frames, truths, K, ts, fams = synth.make_config_frames("C1", 4)
    H, W = frames.shape[1:]
    det = capi.Detector(W, H, families=fams, encoding="mono8", max_batch=2, max_tags=64)
    t, ptrs, pitch = pu.upload(frames)
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    ref = capi.Detector(W, H, families=fams, encoding="mono8", max_batch=2, max_tags=64)
    want01 = ref.detect_device(ptrs[0:2], pitch, 0)
    want23 = ref.detect_device(ptrs[2:4], pitch, 0)
    got = []
    for it in range(5):
        sel = ptrs[0:2] if it % 2 == 0 else ptrs[2:4]
        got.append(det.detect_device(sel, pitch, st.cuda_stream))
    for it in range(5):
        want = want01 if it % 2 == 0 else want23
        for a, b in zip(got[it], want):
            assert a.tobytes() == b.tobytes(), it
    assert sum(len(x) for x in want01) >= 2
    det.close()
    ref.close()


def test_refine_kernels_agree(pu):
    """Batches of at most four frames refine with k_refine_cta (one CTA per quad, one warp per edge), larger ones with the
    warp-per-quad kernel: the same frames must give byte-identical detections either way."""
    from isaac_ros_apriltag_b200 import capi, synth
    frames, truths, K, ts, fams = synth.make_config_frames("C1", 6)
    H, W = frames.shape[1:]
    det = capi.Detector(W, H, families=fams, encoding="mono8", max_batch=6, max_tags=64)
    t, ptrs, pitch = pu.upload(frames)
    big = det.detect_device(ptrs, pitch, pu.current_stream())
    small = det.detect_device(ptrs[:3], pitch, pu.current_stream()) + det.detect_device(ptrs[3:], pitch, pu.current_stream())
    for a, b in zip(big, small):
        assert a.tobytes() == b.tobytes()
    assert sum(len(x) for x in big) >= 6
    det.close()


def test_two_device_batches_in_flight(pu):
    """b200AprilTagsEnqueueBatch twice before the first CollectBatch (same stream): the second batch queues behind the first on
    the one workspace, each has its own pinned result buffers and cached graph; results equal the synchronous calls', in order.
    A third batch, or another stream, is refused while two / one are in flight."""
    needs_real_gpu(pu)
    import torch
    from isaac_ros_apriltag_b200 import capi, synth
    frames, truths, K, ts, fams = synth.make_config_frames("C1", 6)
    H, W = frames.shape[1:]
    det = capi.Detector(W, H, families=fams, encoding="mono8", max_batch=2, max_tags=64)
    t, ptrs, pitch = pu.upload(frames)
    st, st2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    want = [det.detect_device(ptrs[2 * k:2 * k + 2], pitch, st.cuda_stream) for k in range(3)]
    for rounds in range(3):  # plain launches, capture, replay -- per slot
        got = []
        det.enqueue(ptrs[0:2], pitch, st.cuda_stream)
        det.enqueue(ptrs[2:4], pitch, st.cuda_stream)
        with pytest.raises(capi.B200ATError):
            det.enqueue(ptrs[4:6], pitch, st.cuda_stream)   # two already in flight
        got.append(det.collect())
        with pytest.raises(capi.B200ATError):
            det.enqueue(ptrs[4:6], pitch, st2.cuda_stream)  # another stream while one is in flight
        det.enqueue(ptrs[4:6], pitch, st.cuda_stream)
        got.append(det.collect())
        got.append(det.collect())
        for k in range(3):
            for a, b in zip(got[k], want[k]):
                assert a.tobytes() == b.tobytes(), (rounds, k)
    assert sum(len(x) for w in want for x in w) >= 6
    det.close()


def test_bench_workload_frames_match_oracle(pu):
    """The 32 distinct frames bench.py tiles into its 256-frame batch (C2, bgr8): tag IDs and Hamming distance exact,
    corners / centre within TOL_CORNER_PX, pose within tolerance, for every frame, through the batch entry point."""
    import torch
    from isaac_ros_apriltag_b200 import capi, synth
    from oracle import oracle as O
    frames, truths, K, ts, fams = synth.make_config_frames("C2", 32)
    bgr = np.ascontiguousarray(np.repeat(frames[:, :, :, None], 3, axis=3))
    H, W = frames.shape[1:]
    det = capi.Detector(W, H, intrinsics=(K[0, 0], K[1, 1], K[0, 2], K[1, 2]), tag_size=ts, families=fams, encoding="bgr8",
                        max_batch=32, max_tags=64)
    t, ptrs, pitch = pu.upload(bgr)
    gd = det.detect_device(ptrs, pitch, pu.current_stream())
    assert det.status() == 0
    od, _ = O.detect_batch(bgr, fams, nthreads=min(16, os.cpu_count() or 1), encoding="bgr8")
    ndet = 0
    for i in range(32):
        assert [(d["id"], d["hamming"]) for d in od[i]] == [(int(a), int(b)) for a, b in zip(gd[i]["id"], gd[i]["hamming"])], i
        found = set(d["id"] for d in od[i] if d["hamming"] == 0)
        truth_ids = set(tr["id"] for tr in truths[i])
        assert found <= truth_ids and len(found) >= len(truth_ids) - 1, i  # detector recall on the synthetic truth (both arms alike)
        for a, b in zip(gd[i], od[i]):
            assert np.abs(a["p"] - b["p"]).max() <= TOL_CORNER_PX and np.abs(a["c"] - b["c"]).max() <= TOL_CORNER_PX
            assert abs(float(a["decision_margin"]) - b["decision_margin"]) <= TOL_MARGIN
            ndet += 1
    assert ndet >= 0.97 * 320
    det.close()


def test_async_host_calls_two_in_flight(pu, monkeypatch):
    """b200AprilTagsEnqueueBatchHost / CollectBatchHost: two batches in flight, the second one's DMA and quad detection overlapping
    the first one's decode / pose / D2H.  Batches of different content and length, several rounds, both staging modes: every batch
    comes back byte-identical to the device-pointer entry point's result for ITS frames, in order."""
    from isaac_ros_apriltag_b200 import capi, synth
    monkeypatch.setenv("B200AT_SPARSE_DEBUG", "1")
    frames, truths, K, ts, fams = synth.make_config_frames("C1", 10)
    frames = _as_encoding(frames, "bgr8")
    H, W = frames.shape[1:3]
    det = capi.Detector(W, H, families=fams, encoding="bgr8", max_batch=4, max_tags=16)
    t, ptrs, pitch = pu.upload(frames)
    want = [det.detect_device([ptrs[i]], pitch, pu.current_stream())[0] for i in range(10)]
    assert all(len(w) == 1 for w in want)
    if pu.EMU:
        host = frames
    else:
        import torch
        host = torch.from_numpy(frames).pin_memory().numpy()
    batches = [list(range(0, 4)), list(range(4, 7)), list(range(7, 10)), [9, 3, 5, 1], [2], list(range(0, 8))]
    for mode in ("1", "0"):
        monkeypatch.setenv("B200AT_SPARSE_H2D", mode)
        for sub in (None, "1"):
            if sub:
                monkeypatch.setenv("B200AT_HOST_SUB", sub)
            keep = [np.ascontiguousarray(host[b]) if pu.EMU else __import__("torch").from_numpy(np.ascontiguousarray(host[b])).pin_memory().numpy() for b in batches]
            got = []
            det.enqueue_host(keep[0])
            for i in range(1, len(batches)):
                det.enqueue_host(keep[i])          # two in flight
                with pytest.raises(capi.B200ATError):
                    det.enqueue_host(keep[i])      # a third is refused (and leaves the two untouched)
                got.append(det.collect_host())
                assert det.counters()["sparse_h2d"] == int(mode)
            got.append(det.collect_host())
            for b, g in zip(batches, got):
                assert len(g) == len(b)
                for i, a in zip(b, g):
                    assert a.tobytes() == want[i].tobytes(), (mode, sub, b, i)
            if sub:
                monkeypatch.delenv("B200AT_HOST_SUB")
    # the synchronous entry point refuses to run while asynchronous batches are in flight, and works again afterwards
    det.enqueue_host(keep[0])
    with pytest.raises(capi.B200ATError):
        det.detect_host(keep[1])
    det.collect_host()
    assert det.detect_host(keep[1])[0].tobytes() == want[4].tobytes()
    det.close()


def test_full_size_batch_every_position_against_oracle(pu):
    """BASELINE config C2 at its full batch (256 x 1080p bgr8, the 32 distinct seeded frames bench.py tiles): EVERY one of the 256
    positions of the batch is compared with the oracle's result for the frame it holds -- tag ids and Hamming distance exact,
    corners / centre / decision margin within tolerance."""
    if pu.EMU:
        pytest.skip("batch 256 x 1080p is a GPU-size test")
    import torch
    from isaac_ros_apriltag_b200 import capi, synth
    from oracle import oracle as O
    distinct, B = 32, 256
    frames, truths, K, ts, fams = synth.make_config_frames("C2", distinct)
    bgr = _as_encoding(frames, "bgr8")
    H, W = frames.shape[1:]
    det = capi.Detector(W, H, intrinsics=(K[0, 0], K[1, 1], K[0, 2], K[1, 2]), tag_size=ts, families=fams, encoding="bgr8", max_batch=B, max_tags=64)
    dev = torch.from_numpy(bgr).cuda()
    fb = dev[0].numel()
    # position i holds distinct frame (5 i + i // 32) % 32: neighbouring positions differ, every frame appears at 8 positions
    order = [(5 * i + i // 32) % distinct for i in range(B)]
    ptrs = [dev.data_ptr() + order[i] * fb for i in range(B)]
    gd = det.detect_device(ptrs, W * 3, pu.current_stream())
    assert det.status() == 0
    od, _ = O.detect_batch(bgr, fams, nthreads=min(16, os.cpu_count() or 1), encoding="bgr8")
    ndet = 0
    for i in range(B):
        o = od[order[i]]
        assert [(d["id"], d["hamming"]) for d in o] == [(int(a), int(b)) for a, b in zip(gd[i]["id"], gd[i]["hamming"])], i
        for a, b in zip(gd[i], o):
            assert np.abs(a["p"] - b["p"]).max() <= TOL_CORNER_PX and np.abs(a["c"] - b["c"]).max() <= TOL_CORNER_PX, i
            assert abs(float(a["decision_margin"]) - b["decision_margin"]) <= TOL_MARGIN
            ndet += 1
    assert ndet >= 0.97 * 10 * B
    # the host entry point on the same batch (pinned frames, sparse staging, four sub-batches, two workspace views)
    host = torch.from_numpy(bgr[order]).pin_memory().numpy()
    hd = det.detect_host(host)
    for i in range(B):
        assert hd[i].tobytes() == gd[i].tobytes(), i
    det.close()


def test_two_devices_one_process(pu):
    """One process, one handle per GPU (b200AprilTagsOptions_t::device): the per-device kernel attributes, workspaces and
    streams are independent, the caller's current device is left untouched, results are identical on both GPUs."""
    needs_real_gpu(pu)
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from isaac_ros_apriltag_b200 import capi, synth
    frames, truths, K, ts, fams = synth.make_config_frames("C2", 2)
    H, W = frames.shape[1:]
    torch.cuda.set_device(0)
    dets, bufs = [], []
    for dev in (0, 1):
        d = capi.Detector(W, H, families=fams, encoding="mono8", max_batch=2, max_tags=64, device=dev)
        assert torch.cuda.current_device() == 0
        t = torch.from_numpy(frames).to(f"cuda:{dev}")
        bufs.append(t)
        dets.append(d)
    res = []
    for dev, (d, t) in enumerate(zip(dets, bufs)):
        fb = t[0].numel()
        res.append(d.detect_device([t.data_ptr(), t.data_ptr() + fb], W, 0))
        assert torch.cuda.current_device() == 0
    for a, b in zip(res[0], res[1]):
        assert a.tobytes() == b.tobytes() and len(a) == 10
    for d in dets:
        d.close()


@pytest.mark.parametrize("config,enc", [("C2", "bgr8"), ("C4", "mono8")])
def test_sparse_host_path_matches_full_copy(pu, config, enc, monkeypatch):
    """b200AprilTagsDetectBatchHost with pinned frames: staging only every f-th row by DMA and fetching the rows around the quads
    on demand (B200AT_SPARSE_H2D=1) gives byte-identical results to the full copy and to the device-pointer entry point, and
    moves fewer bytes over PCIe.  B200AT_SPARSE_DEBUG poisons the staging slots."""
    from isaac_ros_apriltag_b200 import capi, synth
    monkeypatch.setenv("B200AT_SPARSE_DEBUG", "1")
    n = 6
    frames, truths, K, ts, fams = synth.make_config_frames(config, n)
    if enc == "bgr8":
        frames = np.ascontiguousarray(np.repeat(frames[:, :, :, None], 3, axis=3))
    H, W = frames.shape[1:3]
    det = capi.Detector(W, H, families=fams, encoding=enc, max_batch=4, max_tags=128)
    t, ptrs, pitch = pu.upload(frames)
    want = det.detect_device(ptrs[:4], pitch, pu.current_stream()) + det.detect_device(ptrs[4:], pitch, pu.current_stream())
    if pu.EMU:
        host = frames
    else:
        import torch
        host = torch.from_numpy(frames).pin_memory().numpy()
    monkeypatch.setenv("B200AT_SPARSE_H2D", "0")
    full = det.detect_host(host)
    cf = det.counters()
    monkeypatch.setenv("B200AT_SPARSE_H2D", "1")
    got = det.detect_host(host)
    cs = det.counters()
    assert cf["sparse_h2d"] == 0 and cf["h2d_bytes"] == frames.nbytes
    # C2 (10 tags / frame) needs a fraction of the odd rows; C4 (a 13 x 7 grid of tags) needs nearly all of them
    assert cs["sparse_h2d"] == 1 and cs["h2d_bytes"] < (0.85 if config == "C2" else 1.05) * frames.nbytes
    for i in range(n):
        assert full[i].tobytes() == want[i].tobytes(), i
        assert got[i].tobytes() == want[i].tobytes(), i
    assert sum(len(x) for x in got) >= n
    # sub-batches of 1 and 2 frames: the 6 frames then cycle through the three staging slots, the two workspace views and the four
    # counter blocks of the pipelined path (FRONT(k+1) on the compute stream while the fetch stream serves sub-batch k)
    for sub in ("1", "2"):
        monkeypatch.setenv("B200AT_HOST_SUB", sub)
        for mode in ("0", "1"):
            monkeypatch.setenv("B200AT_SPARSE_H2D", mode)
            for _ in range(2):
                got4 = det.detect_host(host)
                c4 = det.counters()
                assert c4["sparse_h2d"] == int(mode) and c4["detections"] == sum(len(x) for x in want)
                for i in range(n):
                    assert got4[i].tobytes() == want[i].tobytes(), (sub, mode, i)
    monkeypatch.delenv("B200AT_HOST_SUB")
    # separately allocated (pinned) frames: not contiguous in the caller's memory, so every frame gets its own DMA instead of one
    # 2-D copy per run of frames
    if not pu.EMU:
        import torch
        keep = [torch.from_numpy(frames[i].copy()).pin_memory() for i in range(n)]
        for mode in ("0", "1"):
            monkeypatch.setenv("B200AT_SPARSE_H2D", mode)
            got5 = det.detect_host([k.numpy() for k in keep])
            assert det.counters()["sparse_h2d"] == int(mode)
            for i in range(n):
                assert got5[i].tobytes() == want[i].tobytes(), (mode, i)
    monkeypatch.setenv("B200AT_SPARSE_H2D", "1")
    # frames in pageable memory: the call falls back to the full copy by itself
    if not pu.EMU:
        got2 = det.detect_host(frames.copy())
        assert det.counters()["sparse_h2d"] == 0
        for i in range(n):
            assert got2[i].tobytes() == want[i].tobytes(), i
    det.close()
