"""ctypes binding of the CPU ORACLE (oracle/libapriltag_oracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module; nothing under isaac_ros_apriltag_b200/ does.  See oracle/apriltag_oracle.h for provenance.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libapriltag_oracle.so")

FAMILY_NAMES = ["tag36h11", "tag25h9", "tag16h5", "tag36h10", "custom0", "custom1"]
ENCODINGS = {"mono8": 0, "rgb8": 1, "bgr8": 2, "rgba8": 3, "bgra8": 4}


class Params(C.Structure):
    _fields_ = [("quad_decimate", C.c_float), ("quad_sigma", C.c_float), ("refine_edges", C.c_int),
                ("decode_sharpening", C.c_double), ("min_cluster_pixels", C.c_int), ("max_nmaxima", C.c_int),
                ("critical_rad", C.c_float), ("max_line_fit_mse", C.c_float), ("min_white_black_diff", C.c_int),
                ("tile_size", C.c_int), ("max_hamming", C.c_int), ("family_mask", C.c_uint32)]


class Detection(C.Structure):
    _fields_ = [("family", C.c_int), ("id", C.c_int), ("hamming", C.c_int), ("decision_margin", C.c_float),
                ("H", C.c_double * 9), ("c", C.c_double * 2), ("p", (C.c_double * 2) * 4)]


class Pose(C.Structure):
    _fields_ = [("R", C.c_double * 9), ("t", C.c_double * 3), ("err", C.c_double)]


class Quad(C.Structure):
    _fields_ = [("p", (C.c_float * 2) * 4), ("reversed_border", C.c_int), ("key", C.c_uint64)]


class Times(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("decimate", "blur", "threshold", "unionfind", "clusters", "fit_quads", "decode", "reconcile", "total")]


def build(force=False):
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "apriltag_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.ato_create.restype = C.c_void_p
        L.ato_create.argtypes = [C.POINTER(Params)]
        L.ato_destroy.argtypes = [C.c_void_p]
        L.ato_detect.restype = C.c_int
        L.ato_detect.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Detection), C.c_int]
        L.ato_get_times.argtypes = [C.c_void_p, C.POINTER(Times)]
        L.ato_get_quad_dims.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        for fn in ("ato_get_quad_image", "ato_get_threshold"):
            getattr(L, fn).argtypes = [C.c_void_p, C.c_void_p]
        L.ato_get_tile_minmax.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ato_get_labels.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ato_num_clusters.argtypes = [C.c_void_p]
        L.ato_num_clusters.restype = C.c_int
        L.ato_get_cluster.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.c_void_p, C.c_int]
        L.ato_get_cluster.restype = C.c_int
        L.ato_num_points_total.argtypes = [C.c_void_p]
        L.ato_num_points_total.restype = C.c_int
        L.ato_get_quads.argtypes = [C.c_void_p, C.POINTER(Quad), C.c_int, C.c_int]
        L.ato_get_quads.restype = C.c_int
        L.ato_estimate_pose.argtypes = [C.POINTER(Detection)] + [C.c_double] * 5 + [C.POINTER(Pose)] * 3
        L.ato_to_gray.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ato_detect_batch.argtypes = [C.POINTER(Params), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(Detection), C.POINTER(C.c_int), C.c_int, C.POINTER(Times)]
        L.ato_detect_batch.restype = C.c_int
        L.ato_detect_batch_enc.argtypes = [C.POINTER(Params), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.POINTER(Detection), C.POINTER(C.c_int), C.c_int, C.POINTER(Times)]
        L.ato_detect_batch_enc.restype = C.c_int
        L.ato_rotate90.argtypes = [C.c_uint64, C.c_int]
        L.ato_rotate90.restype = C.c_uint64
        L.ato_family_info.argtypes = [C.c_int] + [C.POINTER(C.c_int)] * 4
        L.ato_family_code.argtypes = [C.c_int, C.c_int]
        L.ato_family_code.restype = C.c_uint64
        L.ato_register_family.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def register_family(slot, fam):
    """fam: dict with nbits, width_at_border, total_width, reversed_border, bit_x, bit_y, codes (isaac_ros_apriltag_b200.families layout);
    slot 4 or 5 ("custom0" / "custom1")."""
    bx = np.asarray(fam["bit_x"], np.int8)
    by = np.asarray(fam["bit_y"], np.int8)
    codes = np.asarray(fam["codes"], np.uint64)
    rc = lib().ato_register_family(slot, FAMILY_NAMES[slot].encode(), int(fam["nbits"]), len(codes), int(fam["width_at_border"]),
                                   int(fam["total_width"]), int(bool(fam.get("reversed_border", False))), bx.ctypes.data, by.ctypes.data,
                                   codes.ctypes.data)
    assert rc == 0, rc


def default_params(families=("tag36h11",), **kw):
    p = Params()
    lib().ato_default_params(C.byref(p))
    mask = 0
    for f in families:
        mask |= 1 << FAMILY_NAMES.index(f)
    p.family_mask = mask
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def det_to_dict(d):
    return {"family": FAMILY_NAMES[d.family], "id": d.id, "hamming": d.hamming, "decision_margin": float(d.decision_margin),
            "H": np.array(d.H[:]).reshape(3, 3), "c": np.array(d.c[:]), "p": np.array([list(r) for r in d.p])}


def to_gray(img, encoding):
    """colour -> gray with the boundary's fixed-point formula (same as cv2.cvtColor for 8-bit)."""
    img = np.ascontiguousarray(img)
    h, w = img.shape[:2]
    out = np.empty((h, w), np.uint8)
    lib().ato_to_gray(img.ctypes.data, ENCODINGS[encoding], w, h, img.strides[0], out.ctypes.data)
    return out


class Oracle:
    """One detector instance; keeps the intermediates of the last detect() for stage parity tests."""

    def __init__(self, families=("tag36h11",), **kw):
        self.params = default_params(families, **kw)
        self.h = lib().ato_create(C.byref(self.params))
        self.max_out = 4096

    def __del__(self):
        try:
            if self.h:
                lib().ato_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def detect(self, gray):
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        assert gray.ndim == 2
        out = (Detection * self.max_out)()
        n = lib().ato_detect(self.h, gray.ctypes.data, gray.shape[1], gray.shape[0], gray.strides[0], out, self.max_out)
        assert n >= 0
        self._raw = out
        return [det_to_dict(out[i]) for i in range(n)]

    def raw_detections(self):
        return self._raw

    def times(self):
        t = Times()
        lib().ato_get_times(self.h, C.byref(t))
        return {n: getattr(t, n) for n, _ in Times._fields_}

    def quad_dims(self):
        w, h = C.c_int(), C.c_int()
        lib().ato_get_quad_dims(self.h, C.byref(w), C.byref(h))
        return w.value, h.value

    def quad_image(self):
        w, h = self.quad_dims()
        a = np.empty((h, w), np.uint8)
        lib().ato_get_quad_image(self.h, a.ctypes.data)
        return a

    def threshold(self):
        w, h = self.quad_dims()
        a = np.empty((h, w), np.uint8)
        lib().ato_get_threshold(self.h, a.ctypes.data)
        return a

    def tile_minmax(self):
        w, h = self.quad_dims()
        ts = self.params.tile_size
        mn = np.empty((h // ts, w // ts), np.uint8)
        mx = np.empty_like(mn)
        lib().ato_get_tile_minmax(self.h, mn.ctypes.data, mx.ctypes.data)
        return mn, mx

    def labels(self):
        w, h = self.quad_dims()
        lab = np.empty((h, w), np.uint32)
        sz = np.empty((h, w), np.uint32)
        lib().ato_get_labels(self.h, lab.ctypes.data, sz.ctypes.data)
        return lab, sz

    def clusters(self):
        """list of (key, packed points uint32 x|y<<16 in line-fit order) for clusters with 24 <= n <= bound."""
        n = lib().ato_num_clusters(self.h)
        res = []
        buf = np.empty(1 << 16, np.uint32)
        for i in range(n):
            key = C.c_uint64()
            cnt = lib().ato_get_cluster(self.h, i, C.byref(key), buf.ctypes.data, buf.size)
            if cnt > buf.size:
                buf = np.empty(cnt, np.uint32)
                cnt = lib().ato_get_cluster(self.h, i, C.byref(key), buf.ctypes.data, buf.size)
            res.append((key.value, buf[:cnt].copy()))
        return res

    def num_points_total(self):
        return lib().ato_num_points_total(self.h)

    def quads(self, refined=False):
        out = (Quad * 8192)()
        n = lib().ato_get_quads(self.h, out, 8192, 1 if refined else 0)
        n = min(n, 8192)
        return [{"p": np.array([list(r) for r in out[i].p], np.float32), "reversed_border": out[i].reversed_border,
                 "key": out[i].key} for i in range(n)]

    def estimate_pose(self, i, fx, fy, cx, cy, tagsize):
        best, p1, p2 = Pose(), Pose(), Pose()
        lib().ato_estimate_pose(C.byref(self._raw[i]), fx, fy, cx, cy, tagsize, C.byref(best), C.byref(p1), C.byref(p2))
        f = lambda p: {"R": np.array(p.R[:]).reshape(3, 3), "t": np.array(p.t[:]), "err": p.err}
        return f(best), f(p1), f(p2)


def detect_batch(frames, families=("tag36h11",), nthreads=1, max_out=256, encoding="mono8", **kw):
    """frames: (n,h,w[,c]) uint8.  Returns (list of list of det dicts, summed stage times dict)."""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    n, h, w = frames.shape[:3]
    p = default_params(families, **kw)
    out = (Detection * (n * max_out))()
    counts = (C.c_int * n)()
    t = Times()
    lib().ato_detect_batch_enc(C.byref(p), frames.ctypes.data, ENCODINGS[encoding], n, w, h, nthreads, out, counts, max_out,
                               C.byref(t))
    res = [[det_to_dict(out[i * max_out + k]) for k in range(counts[i])] for i in range(n)]
    return res, {nm: getattr(t, nm) for nm, _ in Times._fields_}
