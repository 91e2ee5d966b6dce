"""Stage-by-stage comparison of the CUDA path (through the C ABI) with the CPU oracle on the same frames.

The oracle is the checker here, never the thing shipped (see oracle/apriltag_oracle.h)."""
import numpy as np
from scipy import ndimage

from isaac_ros_apriltag_b200 import capi
from oracle import oracle as O


# Set by tests/test_emu_parity.py (use_emulator()): the kernels then run under the CPU SIMT emulator of tools/emu (test
# infrastructure; "device" memory is host memory there), otherwise on the GPU through libb200apriltags.so.
EMU = False


def use_emulator():
    """Point the ctypes binding at the emulated build of the SAME kernel sources (tools/emu/build_emu.py)."""
    global EMU
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "emu", "build_emu.py")
    spec = importlib.util.spec_from_file_location("build_emu", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    saved = (capi.LIB_PATH, capi._lib, EMU)
    capi.LIB_PATH = mod.build()
    capi._lib = None
    EMU = True
    return saved


def restore(saved):
    global EMU
    capi.LIB_PATH, capi._lib, EMU = saved


def current_stream():
    if EMU:
        return 0
    import torch
    return torch.cuda.current_stream().cuda_stream


class _HostFrames:
    """numpy stand-in for the torch tensor that keeps the frames alive (emulator runs only)"""

    def __init__(self, a):
        self.a = a


def upload(frames):
    """host (n,H,W[,C]) uint8 -> torch cuda tensor; returns (tensor, ptrs, pitch)"""
    if EMU:
        a = np.ascontiguousarray(frames)
        fb = a[0].size
        pitch = a.shape[2] * (a.shape[3] if a.ndim == 4 else 1)
        return _HostFrames(a), [a.ctypes.data + i * fb for i in range(a.shape[0])], pitch
    import torch
    t = torch.from_numpy(np.ascontiguousarray(frames)).cuda()
    n = t.shape[0]
    frame_bytes = t[0].numel()
    pitch = t.shape[2] * (t.shape[3] if t.dim() == 4 else 1)
    ptrs = [t.data_ptr() + i * frame_bytes for i in range(n)]
    return t, ptrs, pitch


def to_gray_batch(frames, encoding):
    if encoding == "mono8":
        return frames
    return np.stack([O.to_gray(f, encoding) for f in frames])


def compare_stages(frames, encoding="mono8", families=("tag36h11",), report=None, **opts):
    """Runs the batch through the GPU detector and every frame through the oracle; returns a dict of per-stage
    mismatch counts (all zeros = bit-exact) plus float deviations for the tolerance-level stages."""
    n, H, W = frames.shape[:3]
    det = capi.Detector(W, H, families=families, encoding=encoding, max_batch=n, max_tags=256, **opts)
    t, ptrs, pitch = upload(frames)
    gdets = det.detect_device(ptrs, pitch, current_stream(), strict=False)
    res = {"status": det.status(), "dec": 0, "tile": 0, "thr": 0, "labels": 0, "sizes": 0, "clusters": 0, "points": 0,
           "quads_n": 0, "quads_bits": 0, "quads_max": 0.0, "refined_n": 0, "refined_bits": 0, "refined_max": 0.0,
           "det_n": 0, "det_id": 0, "det_margin_max": 0.0, "det_corner_max": 0.0, "det_H_max": 0.0, "n_det": 0, "n_quads": 0,
           "n_clusters": 0, "n_points": 0}
    gray = to_gray_batch(frames, encoding)
    okw = {}
    for k in ("quad_decimate", "quad_sigma", "refine_edges", "decode_sharpening", "min_white_black_diff", "max_nmaxima",
              "critical_rad", "max_line_fit_mse", "max_hamming", "tile_size"):
        if k in opts:
            okw[k] = opts[k]
    orc = O.Oracle(families, **okw)
    gclu = det.read_buffer(capi.BUF_CLUSTERS)
    gkeys = det.read_buffer(capi.BUF_POINTS)
    graw = det.read_buffer(capi.BUF_POINTS_RAW)
    gquads = det.read_buffer(capi.BUF_QUADS)
    gref = det.read_buffer(capi.BUF_QUADS_REFINED)
    ts = opts.get("tile_size", 4)
    for i in range(n):
        odets = orc.detect(gray[i])
        # dense stages
        res["dec"] += int((det.read_buffer(capi.BUF_DECIMATED, i) != orc.quad_image()).sum())
        omn, omx = orc.tile_minmax()
        gmn = ndimage.minimum_filter(det.read_buffer(capi.BUF_TILE_MIN, i), size=3, mode="nearest")
        gmx = ndimage.maximum_filter(det.read_buffer(capi.BUF_TILE_MAX, i), size=3, mode="nearest")
        res["tile"] += int((gmn != omn).sum() + (gmx != omx).sum())
        res["thr"] += int((det.read_buffer(capi.BUF_THRESHOLD, i) != orc.threshold()).sum())
        olab, osz = orc.labels()
        glab = det.read_buffer(capi.BUF_LABELS, i)
        gsz = det.read_buffer(capi.BUF_SIZES, i)
        res["labels"] += int((glab != olab).sum())
        # GPU stores the size at the representative only (and never counts 127 pixels' singletons elsewhere)
        gsz_full = gsz.reshape(-1)[glab.reshape(-1)].reshape(glab.shape)
        res["sizes"] += int((gsz_full != osz).sum())
        # clusters: same key set, same sorted point sequence
        ocl = orc.clusters()
        sel = gclu[gclu["frame"] == i]
        gmap = {int(r["key"]): r for r in sel}
        res["n_clusters"] += len(ocl)
        if set(gmap) != set(k for k, _ in ocl):
            res["clusters"] += len(set(gmap) ^ set(k for k, _ in ocl))
        for key, pts in ocl:
            r = gmap.get(key)
            if r is None:
                continue
            res["n_points"] += len(pts)
            lo, hi = int(r["offset"]), int(r["offset"]) + int(r["count"])
            # (1) emitted point multiset == oracle's (emission order inside a cluster is free)
            rw = graw[lo:hi]
            rp = ((rw & np.uint32(0x3fff)) | (((rw >> np.uint32(14)) & np.uint32(0x3fff)) << np.uint32(16))).astype(np.uint32)
            if len(rp) != len(pts) or not np.array_equal(np.sort(rp), np.sort(pts)):
                res["points"] += 1
                continue
            # (2) clusters that reach the sort (bbox / polarity gates passed on both sides) have the same sequence
            gk = gkeys[lo:hi]
            gp = ((gk & np.uint64(0xffff)) | (((gk >> np.uint64(16)) & np.uint64(0xffff)) << np.uint64(16))).astype(np.uint32)
            if np.array_equal(np.sort(gp), np.sort(pts)):
                res["n_sorted"] = res.get("n_sorted", 0) + 1
                if not np.array_equal(gp, pts):
                    res["order"] = res.get("order", 0) + 1
        # quads
        for which, garr, tagn, tagb, tagm in ((False, gquads, "quads_n", "quads_bits", "quads_max"),
                                              (True, gref, "refined_n", "refined_bits", "refined_max")):
            oq = orc.quads(refined=which)
            if which:  # the oracle applies refine_edges inside the decode loop; recompute refined corners here
                pass
            gq = garr[garr["frame"] == i]
            gq = gq[np.argsort(gq["key"], kind="stable")]
            if not which:
                res["n_quads"] += len(oq)
            # quads are matched by their cluster key (a cluster yields at most one quad): set difference counted, corners of
            # the common ones compared
            omap = {int(q["key"]): q["p"] for q in oq}
            gmap2 = {int(k): p for k, p in zip(gq["key"], gq["p"])}
            res[tagn] += len(set(omap) ^ set(gmap2)) + (0 if len(gmap2) == len(gq) else 1)
            common = sorted(set(omap) & set(gmap2))
            if common:
                op = np.stack([omap[k] for k in common])
                gp2 = np.stack([gmap2[k] for k in common])
                res[tagb] += int((gp2.view(np.uint32) != op.view(np.uint32)).sum())
                res[tagm] = max(res[tagm], float(np.abs(gp2 - op).max()))
        # detections
        g = gdets[i]
        res["n_det"] += len(odets)
        if len(g) != len(odets):
            res["det_n"] += 1
            if report is not None:
                report.append(f"frame {i}: gpu {len(g)} dets vs oracle {len(odets)}: "
                              f"gpu ids {list(g['id'])} oracle ids {[d['id'] for d in odets]}")
            continue
        for a, b in zip(g, odets):
            if a["id"] != b["id"] or a["hamming"] != b["hamming"] or capi.FAMILY_NAMES[a["family"]] != b["family"]:
                res["det_id"] += 1
                continue
            res["det_margin_max"] = max(res["det_margin_max"], abs(float(a["decision_margin"]) - b["decision_margin"]))
            res["det_corner_max"] = max(res["det_corner_max"], float(np.abs(a["p"] - b["p"]).max()),
                                        float(np.abs(a["c"] - b["c"]).max()))
            res["det_H_max"] = max(res["det_H_max"], float(np.abs(a["H"].reshape(3, 3) - b["H"]).max()))
    det.close()
    del t
    return res, gdets
