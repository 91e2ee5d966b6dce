"""Kernel LOGIC check without a GPU: the detector's own .cu sources, compiled for the CPU SIMT emulator of tools/emu (one
fiber per CUDA thread; barriers, warp collectives, atomics and shared-memory aliasing behave as on the device), run through
the same C ABI and compared stage by stage with the oracle.

This is test infrastructure, like the oracle: the emulated library is built under tools/emu/_build, is loaded only here
(and by `B200AT_TEST_EMU=1 pytest -m gpu`, the developer switch), is never timed and never shipped -- the package itself
only ever loads libb200apriltags.so and fails without a GPU.  The parity tests proper remain tests/test_gpu_parity.py
(-m gpu, on the B200); what this file adds is that a kernel change which breaks a stage is caught in the CPU suite already.
What the emulator cannot show: memory-model races between CTAs, TMA staging (the plain-load staging path is used), timing."""
import numpy as np
import pytest

from test_gpu_parity import assert_exact


@pytest.fixture(scope="module")
def pu():
    import parity_util
    saved = parity_util.use_emulator()
    yield parity_util
    parity_util.restore(saved)


def small_frame(seed, w, h, tags, side=(40, 90)):
    from isaac_ros_apriltag_b200 import synth
    rng = np.random.default_rng(seed)
    g, _ = synth.make_frame(rng, w, h, tags, side_px=side)
    return g


def test_emulated_chain_matches_oracle_mono8(pu):
    frames = np.stack([small_frame(7, 322, 242, [("tag36h11", 3), ("tag36h11", 17)], side=(50, 90)),
                       small_frame(8, 322, 242, [("tag36h11", 100)], side=(60, 120))])
    rep = []
    res, gd = pu.compare_stages(frames, "mono8", ("tag36h11",), report=rep)
    assert_exact(res, rep)
    assert sorted(int(i) for i in gd[0]["id"]) == [3, 17] and list(gd[1]["id"]) == [100]
    assert res["n_clusters"] > 20 and res["n_quads"] >= 3


@pytest.mark.parametrize("encoding", ["bgr8", "rgba8"])
def test_emulated_chain_colour(pu, encoding):
    g = small_frame(21, 400, 300, [("tag36h11", 5), ("tag25h9", 9)], side=(60, 110))
    rng = np.random.default_rng(3)
    ch = 3 if encoding == "bgr8" else 4
    col = rng.integers(0, 40, g.shape + (ch,), dtype=np.int16) + g[:, :, None].astype(np.int16) - 20
    frames = np.clip(col, 0, 255).astype(np.uint8)[None]
    rep = []
    res, gd = pu.compare_stages(frames, encoding, ("tag36h11", "tag25h9"), report=rep)
    assert_exact(res, rep)
    assert len(gd[0]) == 2


@pytest.mark.parametrize("opts", [dict(quad_decimate=1.0), dict(quad_decimate=1.5), dict(quad_decimate=3.0), dict(quad_sigma=0.8),
                                  dict(tile_size=8), dict(refine_edges=0)])
def test_emulated_knobs(pu, opts):
    g = small_frame(5, 333, 251, [("tag36h11", 1), ("tag36h11", 2)], side=(60, 100))  # odd size: partial tiles, padded pitch
    rep = []
    res, _ = pu.compare_stages(np.stack([g, g[::-1].copy()]), "mono8", ("tag36h11",), report=rep, **opts)
    assert_exact(res, rep)


def test_emulated_large_cluster_bins(pu):
    """One big tag: its outer border cluster has thousands of points, which exercises the multi-warp quad-fit bins (keys in
    shared memory, pipelined prefix scan) that small frames never reach."""
    from isaac_ros_apriltag_b200 import synth
    rng = np.random.default_rng(11)
    g, _ = synth.make_frame(rng, 1600, 1200, [("tag36h11", 42)], side_px=(680, 720), max_tilt_deg=10.0, noise_sigma=0.0)
    rep = []
    res, gd = pu.compare_stages(g[None], "mono8", ("tag36h11",), report=rep)
    assert_exact(res, rep)
    assert list(gd[0]["id"]) == [42]
    assert res["n_points"] > 4096


def test_emulated_host_entry_point_matches_device_entry_point(pu):
    from isaac_ros_apriltag_b200 import capi
    frames = np.stack([small_frame(30 + i, 322, 242, [("tag36h11", 10 + i)], side=(60, 110)) for i in range(5)])
    bgr = np.ascontiguousarray(np.repeat(frames[:, :, :, None], 3, axis=3))
    det = capi.Detector(322, 242, encoding="bgr8", max_batch=2, max_tags=16)
    t, ptrs, pitch = pu.upload(bgr)
    want = [det.detect_device(ptrs[i:i + 1], pitch, 0)[0] for i in range(5)]
    got = det.detect_host(bgr)   # 5 frames through a 2-frame workspace: sub-batches + staging slots
    for a, b in zip(got, want):
        assert a.tobytes() == b.tobytes() and len(a) == 1
    # separately allocated frames (not contiguous in the caller's memory: one DMA per frame instead of one per run of frames)
    pad = [np.empty((242 * 322 * 3 + 64 * (i + 1),), np.uint8) for i in range(5)]
    sep = []
    for i in range(5):
        v = pad[i][64 * (i + 1):].reshape(242, 322, 3)
        v[...] = bgr[i]
        sep.append(v)
    got = det.detect_host(sep)
    for a, b in zip(got, want):
        assert a.tobytes() == b.tobytes() and len(a) == 1
    det.close()


def test_emulated_refine_kernels_agree(pu):
    """Batches of at most four frames refine with k_refine_cta (one CTA per quad, one warp per edge), larger ones with the
    warp-per-quad kernel: the same frames must give byte-identical detections either way."""
    from isaac_ros_apriltag_b200 import capi
    frames = np.stack([small_frame(70 + i, 322, 242, [("tag36h11", 40 + i)], side=(60, 110)) for i in range(6)])
    det = capi.Detector(322, 242, encoding="mono8", max_batch=6, max_tags=16)
    t, ptrs, pitch = pu.upload(frames)
    big = det.detect_device(ptrs, pitch, 0)                                               # six frames: warp per quad
    small = det.detect_device(ptrs[:3], pitch, 0) + det.detect_device(ptrs[3:], pitch, 0)  # three frames: CTA per quad
    for a, b in zip(big, small):
        assert a.tobytes() == b.tobytes() and len(a) == 1
    det.close()


def test_emulated_two_device_batches_in_flight(pu):
    """Two b200AprilTagsEnqueueBatch calls before the first CollectBatch: per-slot result buffers, oldest batch collected first."""
    from isaac_ros_apriltag_b200 import capi
    frames = np.stack([small_frame(40 + i, 322, 242, [("tag36h11", 20 + i)], side=(60, 110)) for i in range(6)])
    det = capi.Detector(322, 242, encoding="mono8", max_batch=2, max_tags=16)
    t, ptrs, pitch = pu.upload(frames)
    want = [det.detect_device(ptrs[2 * k:2 * k + 2], pitch, 0) for k in range(3)]
    det.enqueue(ptrs[0:2], pitch, 0)
    det.enqueue(ptrs[2:4], pitch, 0)
    with pytest.raises(capi.B200ATError):
        det.enqueue(ptrs[4:6], pitch, 0)
    got = [det.collect()]
    det.enqueue(ptrs[4:6], pitch, 0)
    got += [det.collect(), det.collect()]
    for k in range(3):
        for a, b in zip(got[k], want[k]):
            assert a.tobytes() == b.tobytes() and len(a) == 1
    det.close()


@pytest.mark.parametrize("enc,dec", [("bgr8", 2.0), ("mono8", 2.0), ("rgba8", 3.0)])
def test_emulated_sparse_host_path(pu, enc, dec, monkeypatch):
    """Sparse staging of the host entry point (only every f-th row by DMA, the rows around the quads fetched on demand): same
    bytes out as the device-pointer path.  B200AT_SPARSE_DEBUG poisons the staging slot, and the emulator poisons fresh
    device memory, so a full-resolution pixel that is read without having been fetched changes the result."""
    from isaac_ros_apriltag_b200 import capi
    monkeypatch.setenv("B200AT_SPARSE_DEBUG", "1")
    frames = np.stack([small_frame(40 + i, 416, 320, [("tag36h11", 20 + i), ("tag36h11", 50 + i)], side=(50, 120)) for i in range(3)])
    ch = {"bgr8": 3, "mono8": 1, "rgba8": 4}[enc]
    if ch > 1:
        frames = np.ascontiguousarray(np.repeat(frames[:, :, :, None], ch, axis=3))
    det = capi.Detector(416, 320, encoding=enc, max_batch=2, max_tags=16, quad_decimate=dec)
    t, ptrs, pitch = pu.upload(frames)
    want = [det.detect_device(ptrs[i:i + 1], pitch, 0)[0] for i in range(3)]
    monkeypatch.setenv("B200AT_SPARSE_H2D", "1")
    got = det.detect_host(frames)
    c = det.counters()
    assert c["sparse_h2d"] == 1 and c["h2d_bytes"] < 0.8 * frames.nbytes
    for a, b in zip(got, want):
        assert a.tobytes() == b.tobytes() and len(a) >= 1
    # sub-batches of one frame: FETCH(k) on its own stream between FRONT(k) and BACK(k), three staging slots, two workspace views
    monkeypatch.setenv("B200AT_HOST_SUB", "1")
    for _ in range(2):
        got4 = det.detect_host(frames)
        c4 = det.counters()
        assert c4["sparse_h2d"] == 1 and c4["detections"] == sum(len(x) for x in want)
        for a, b in zip(got4, want):
            assert a.tobytes() == b.tobytes()
    monkeypatch.delenv("B200AT_HOST_SUB")
    # pageable (not device-mapped) frames fall back to the full copy
    monkeypatch.setenv("B200AT_EMU_HOSTMEM", "pageable")
    got2 = det.detect_host(frames)
    c2 = det.counters()
    assert c2["sparse_h2d"] == 0 and c2["h2d_bytes"] == frames.nbytes
    for a, b in zip(got2, want):
        assert a.tobytes() == b.tobytes()
    det.close()


@pytest.mark.parametrize("tune", ["qf_exact=1", "ccl_tma=0", "ccl_tma=0,qf_exact=1", "qf_bucket_limit=3"])
def test_emulated_kernel_variants(pu, tune, monkeypatch):
    """The kernel variants behind B200AT_TUNE (csrc/detector.h, struct Tune) against the oracle (the GPU suite runs the same list:
    tests/test_gpu_parity.py::test_every_tune_variant_on_the_gpu)."""
    monkeypatch.setenv("B200AT_TUNE", tune)
    frames = np.stack([small_frame(50, 400, 300, [("tag36h11", 7), ("tag36h11", 8)], side=(60, 120)),
                       small_frame(51, 400, 300, [("tag36h11", 9)], side=(150, 200))])
    rep = []
    res, gd = pu.compare_stages(frames, "mono8", ("tag36h11",), report=rep)
    assert_exact(res, rep)
    assert sorted(int(i) for i in gd[0]["id"]) == [7, 8] and list(gd[1]["id"]) == [9]


def test_emulated_host_schedule_under_random_stream_interleavings(pu, monkeypatch):
    """The emulator's asynchronous mode queues every stream operation and runs the queues, at the synchronisation points, in a
    RANDOM interleaving that respects stream order and event dependencies and nothing else: a missing cudaStreamWaitEvent between
    the copy / compute / fetch / tail streams of the pipelined host path turns into wrong results for some seeds (checked by
    fault injection: dropping the wait of BACK on FETCH, or of BACK on the tail of k-2, fails several of the seeds).  Synchronous
    calls and two asynchronous calls in flight (the second call's FRONT is queued before the first call's last BACK)."""
    import ctypes
    from isaac_ros_apriltag_b200 import capi
    monkeypatch.setenv("B200AT_SPARSE_DEBUG", "1")
    frames = np.stack([small_frame(60 + i, 320, 240, [("tag36h11", 30 + i)], side=(60, 110)) for i in range(6)])
    frames = np.ascontiguousarray(np.repeat(frames[:, :, :, None], 3, axis=3))  # (pitch 960 B: a multiple of 16, as sparse staging needs)
    det = capi.Detector(320, 240, encoding="bgr8", max_batch=4, max_tags=16)
    t, ptrs, pitch = pu.upload(frames)
    want = det.detect_device(ptrs[:4], pitch, 0) + det.detect_device(ptrs[4:], pitch, 0)
    L = capi.lib()
    L.b200at_emu_async.argtypes = [ctypes.c_int, ctypes.c_uint]
    monkeypatch.setenv("B200AT_SPARSE_H2D", "1")
    monkeypatch.setenv("B200AT_HOST_SUB", "1")
    a, b = np.ascontiguousarray(frames[:3]), np.ascontiguousarray(frames[3:])
    try:
        for seed in range(6):
            L.b200at_emu_async(1, seed)
            # (the device-pointer path too: the quad-fit bins fork onto seven side streams and join again)
            dev = det.detect_device(ptrs[:4], pitch, 0)
            for i in range(4):
                assert dev[i].tobytes() == want[i].tobytes(), ("device path", seed, i)
            got = det.detect_host(frames)
            c = det.counters()
            assert c["sparse_h2d"] == 1 and c["detections"] == sum(len(x) for x in want), seed
            for i, (x, y) in enumerate(zip(got, want)):
                assert x.tobytes() == y.tobytes(), (seed, i)
            det.enqueue_host(a)
            det.enqueue_host(b)
            ga = det.collect_host()
            det.enqueue_host(a)
            gb = det.collect_host()
            ga2 = det.collect_host()
            for i, (x, y) in enumerate(zip(ga + gb, want)):
                assert x.tobytes() == y.tobytes(), ("async", seed, i)
            for i, (x, y) in enumerate(zip(ga2, want[:3])):
                assert x.tobytes() == y.tobytes(), ("async, third call", seed, i)
    finally:
        L.b200at_emu_async(0, 0)
    det.close()


def test_emulated_host_path_after_encoding_change(pu):
    """b200AprilTagsSetInputEncoding between host calls: the staging slots are re-sized for the larger rows (mono8 -> bgr8)."""
    from isaac_ros_apriltag_b200 import capi
    gray = np.stack([small_frame(70 + i, 320, 240, [("tag36h11", 40 + i)], side=(60, 110)) for i in range(3)])
    bgr = np.ascontiguousarray(np.repeat(gray[:, :, :, None], 3, axis=3))
    det = capi.Detector(320, 240, encoding="mono8", max_batch=2, max_tags=16)
    got_m = det.detect_host(gray)
    assert capi.lib().b200AprilTagsSetInputEncoding(det.h, capi.ENCODINGS["bgr8"]) == 0
    got_c = det.detect_host(bgr)
    t, ptrs, pitch = pu.upload(bgr)
    want = [det.detect_device(ptrs[i:i + 1], pitch, 0)[0] for i in range(3)]
    for a, b, c in zip(got_m, got_c, want):
        assert len(c) == 1 and b.tobytes() == c.tobytes()
        assert list(a["id"]) == list(c["id"])  # same image content through the mono8 path
    det.close()


def test_emulated_sparse_staging_backs_off_on_tag_covered_frames(pu, monkeypatch):
    """A frame covered with tags needs most full-resolution rows: after a sparse call that moved more than the back-off fraction of
    the input (0.85 by default, B200AT_SPARSE_BACKOFF; 0.5 here so that a small frame triggers it) the next calls take the full
    copy by themselves (B200AT_SPARSE_H2D unset), with identical results."""
    from isaac_ros_apriltag_b200 import capi, synth
    monkeypatch.setenv("B200AT_SPARSE_BACKOFF", "0.5")
    rng = np.random.default_rng(4)
    tags = [("tag36h11", i) for i in range(20)]
    g, _ = synth.make_frame(rng, 400, 320, tags, grid=(5, 4, 80, 66))
    frames = np.ascontiguousarray(np.repeat(g[None, :, :, None], 3, axis=3))
    det = capi.Detector(400, 320, encoding="bgr8", max_batch=2, max_tags=32)
    modes, res = [], []
    for _ in range(3):
        res.append(det.detect_host(frames)[0])
        c = det.counters()
        modes.append((c["sparse_h2d"], c["h2d_bytes"] / frames.nbytes))
    assert modes[0][0] == 1 and modes[0][1] > 0.5, modes
    assert modes[1][0] == 0 and modes[2][0] == 0 and modes[1][1] == 1.0, modes
    assert len(res[0]) >= 4 and all(r.tobytes() == res[0].tobytes() for r in res)
    det.close()
