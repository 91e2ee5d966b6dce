"""ctypes binding of the C ABI in include/b200_apriltags.h (libb200apriltags.so).

This is the only way Python reaches the detector: there is NO CPU fallback.  If the shared library is
missing or no CUDA device is present the calls raise (B200ATError / OSError); nothing under
isaac_ros_apriltag_b200/ ever imports oracle/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200apriltags.so")

FAMILY_NAMES = ["tag36h11", "tag25h9", "tag16h5", "tag36h10", "custom0", "custom1"]  # custom*: b200AprilTagsRegisterFamily slots
ENCODINGS = {"mono8": 0, "rgb8": 1, "bgr8": 2, "rgba8": 3, "bgra8": 4}
BPP = {"mono8": 1, "rgb8": 3, "bgr8": 3, "rgba8": 4, "bgra8": 4}
ERRORS = {1: "INVALID_ARG", 2: "UNSUPPORTED", 3: "CUDA", 4: "NOMEM", 5: "OVERFLOW", 6: "NO_DEVICE"}
STAGES = ["preprocess", "threshold", "ccl", "cluster", "quadfit", "decode", "finalize", "d2h"]
(BUF_DECIMATED, BUF_TILE_MIN, BUF_TILE_MAX, BUF_THRESHOLD, BUF_LABELS, BUF_SIZES, BUF_CLUSTERS, BUF_POINTS, BUF_QUADS,
 BUF_QUADS_REFINED, BUF_POINTS_RAW, BUF_RECTIFIED) = range(12)

# every symbol include/b200_apriltags.h declares (checked by tests/test_abi.py without a GPU)
EXPORTED_SYMBOLS = [
    "nvCreateAprilTagsDetector", "cuAprilTagsDetect", "cuAprilTagsDestroy",
    "b200AprilTagsDefaultOptions", "b200AprilTagsRegisterFamily", "b200AprilTagsCreate", "b200AprilTagsSetInputEncoding", "b200AprilTagsSetRectification", "b200AprilTagsDetectBatch",
    "b200AprilTagsDetectBatchHost", "b200AprilTagsEnqueueBatchHost", "b200AprilTagsCollectBatchHost", "b200AprilTagsEnqueueBatch", "b200AprilTagsCollectBatch", "b200AprilTagsLastStatus",
    "b200AprilTagsEnableStageTiming", "b200AprilTagsGetStageTimes", "b200AprilTagsGetCounters", "b200AprilTagsGetDims",
    "b200AprilTagsReadBuffer", "b200AprilTagsVersion",
]


class B200ATError(RuntimeError):
    def __init__(self, code, what):
        super().__init__(f"{what} failed: error {code} ({ERRORS.get(code, '?')})")
        self.code = code


class Float2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


class TagID(C.Structure):  # cuAprilTagsID_t: 88 bytes (CUDA's float2 is 8-byte aligned, so the struct has 4 bytes of tail padding)
    _fields_ = [("corners", Float2 * 4), ("id", C.c_uint16), ("hamming_error", C.c_uint8),
                ("orientation", C.c_float * 9), ("translation", C.c_float * 3), ("_tail_pad", C.c_uint32)]


class ImageInput(C.Structure):  # cuAprilTagsImageInput_t
    _fields_ = [("dev_ptr", C.c_void_p), ("pitch", C.c_size_t), ("width", C.c_uint16), ("height", C.c_uint16)]


class Intrinsics(C.Structure):  # cuAprilTagsCameraIntrinsics_t
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float)]


class Options(C.Structure):  # b200AprilTagsOptions_t
    _fields_ = [("struct_size", C.c_uint32), ("family_mask", C.c_uint32), ("max_batch", C.c_uint32), ("max_tags", C.c_uint32),
                ("tile_size", C.c_uint32), ("quad_decimate", C.c_float), ("quad_sigma", C.c_float), ("refine_edges", C.c_int32),
                ("decode_sharpening", C.c_double), ("min_white_black_diff", C.c_int32), ("max_nmaxima", C.c_int32),
                ("critical_rad", C.c_float), ("max_line_fit_mse", C.c_float), ("max_hamming", C.c_int32),
                ("input_encoding", C.c_int32), ("device", C.c_int32), ("hash_slots_per_frame", C.c_uint32),
                ("points_per_frame", C.c_uint32), ("clusters_per_frame", C.c_uint32), ("quads_per_frame", C.c_uint32)]


class Detection(C.Structure):  # b200AprilTagsDetection_t
    _fields_ = [("family", C.c_int32), ("id", C.c_int32), ("hamming", C.c_int32), ("decision_margin", C.c_float),
                ("H", C.c_double * 9), ("c", C.c_double * 2), ("p", (C.c_double * 2) * 4), ("R", C.c_double * 9),
                ("t", C.c_double * 3), ("pose_err", C.c_double)]


class FamilyDesc(C.Structure):  # b200AprilTagsFamilyDesc_t
    _fields_ = [("struct_size", C.c_uint32), ("nbits", C.c_uint32), ("ncodes", C.c_uint32), ("width_at_border", C.c_uint32),
                ("total_width", C.c_uint32), ("reversed_border", C.c_uint32), ("bit_x", C.c_void_p), ("bit_y", C.c_void_p),
                ("codes", C.c_void_p)]


class Rectify(C.Structure):  # b200AprilTagsRectify_t
    _fields_ = [("struct_size", C.c_uint32), ("src_width", C.c_uint32), ("src_height", C.c_uint32), ("K", C.c_double * 9),
                ("D", C.c_double * 8), ("R", C.c_double * 9), ("P", C.c_double * 9)]


class Frame(C.Structure):  # b200AprilTagsFrame_t
    _fields_ = [("ptr", C.c_void_p), ("pitch", C.c_size_t)]


DET_DTYPE = np.dtype([("family", "<i4"), ("id", "<i4"), ("hamming", "<i4"), ("decision_margin", "<f4"), ("H", "<f8", (9,)),
                      ("c", "<f8", (2,)), ("p", "<f8", (4, 2)), ("R", "<f8", (9,)), ("t", "<f8", (3,)), ("pose_err", "<f8")])
ID_DTYPE = np.dtype([("corners", "<f4", (4, 2)), ("id", "<u2"), ("hamming_error", "u1"), ("_pad", "u1"),
                     ("orientation", "<f4", (9,)), ("translation", "<f4", (3,)), ("_tail_pad", "<u4")])
QUAD_DTYPE = np.dtype([("key", "<u8"), ("p", "<f4", (4, 2)), ("frame", "<u4"), ("reversed_border", "<u4")])
CLUSTER_DTYPE = np.dtype([("key", "<u8"), ("offset", "<u4"), ("count", "<u4"), ("frame", "<u4"), ("pad", "<u4")])
assert DET_DTYPE.itemsize == C.sizeof(Detection) and ID_DTYPE.itemsize == C.sizeof(TagID)


def build(force=False):
    """Compile libb200apriltags.so for sm_100a with the committed Makefile (nvcc cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    if force:
        subprocess.check_call(["make", "-C", csrc, "-s", "clean"])
    subprocess.check_call(["make", "-C", csrc, "-s", "-j8"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError(f"{LIB_PATH} is missing: run isaac_ros_apriltag_b200.capi.build() "
                          "(there is no CPU fallback for the detector)")
        L = C.CDLL(LIB_PATH)
        vp, u32 = C.c_void_p, C.c_uint32
        L.nvCreateAprilTagsDetector.argtypes = [C.POINTER(vp), u32, u32, u32, C.c_int, C.POINTER(Intrinsics), C.c_float]
        L.cuAprilTagsDetect.argtypes = [vp, C.POINTER(ImageInput), C.POINTER(TagID), C.POINTER(u32), u32, vp]
        L.cuAprilTagsDetect.restype = u32
        L.cuAprilTagsDestroy.argtypes = [vp]
        L.b200AprilTagsDefaultOptions.argtypes = [C.POINTER(Options)]
        L.b200AprilTagsDefaultOptions.restype = None
        L.b200AprilTagsRegisterFamily.argtypes = [C.c_int32, C.POINTER(FamilyDesc)]
        L.b200AprilTagsCreate.argtypes = [C.POINTER(vp), u32, u32, C.POINTER(Intrinsics), C.c_float, C.POINTER(Options)]
        L.b200AprilTagsSetInputEncoding.argtypes = [vp, C.c_int32]
        L.b200AprilTagsSetRectification.argtypes = [vp, C.POINTER(Rectify)]
        L.b200AprilTagsDetectBatch.argtypes = [vp, C.POINTER(Frame), u32, vp, vp, vp, vp]
        L.b200AprilTagsDetectBatchHost.argtypes = [vp, C.POINTER(Frame), u32, vp, vp, vp]
        L.b200AprilTagsEnqueueBatchHost.argtypes = [vp, C.POINTER(Frame), u32]
        L.b200AprilTagsCollectBatchHost.argtypes = [vp, vp, vp, vp]
        L.b200AprilTagsEnqueueBatch.argtypes = [vp, C.POINTER(Frame), u32, vp]
        L.b200AprilTagsCollectBatch.argtypes = [vp, vp, vp, vp]
        L.b200AprilTagsLastStatus.argtypes = [vp, C.POINTER(u32)]
        L.b200AprilTagsEnableStageTiming.argtypes = [vp, C.c_int]
        L.b200AprilTagsGetStageTimes.argtypes = [vp, C.POINTER(C.c_float)]
        L.b200AprilTagsGetCounters.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.b200AprilTagsGetDims.argtypes = [vp] + [C.POINTER(u32)] * 4
        L.b200AprilTagsReadBuffer.argtypes = [vp, C.c_int, u32, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.b200AprilTagsVersion.restype = C.c_char_p
        _lib = L
    return _lib


def register_family(slot, fam):
    """b200AprilTagsRegisterFamily: fam = dict with nbits, width_at_border, total_width, reversed_border, bit_x, bit_y, codes;
    slot = 4 ("custom0") or 5 ("custom1")."""
    bx = np.asarray(fam["bit_x"], np.int8)
    by = np.asarray(fam["bit_y"], np.int8)
    codes = np.asarray(fam["codes"], np.uint64)
    d = FamilyDesc(C.sizeof(FamilyDesc), int(fam["nbits"]), len(codes), int(fam["width_at_border"]), int(fam["total_width"]),
                   int(bool(fam.get("reversed_border", False))), bx.ctypes.data, by.ctypes.data, codes.ctypes.data)
    rc = lib().b200AprilTagsRegisterFamily(slot, C.byref(d))
    if rc != 0:
        raise B200ATError(rc, "b200AprilTagsRegisterFamily")


def default_options(**kw):
    o = Options()
    lib().b200AprilTagsDefaultOptions(C.byref(o))
    fams = kw.pop("families", None)
    if fams is not None:
        o.family_mask = 0
        for f in fams:
            o.family_mask |= 1 << FAMILY_NAMES.index(f)
    enc = kw.pop("encoding", None)
    if enc is not None:
        o.input_encoding = ENCODINGS[enc]
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


class Detector:
    """Thin RAII wrapper over the handle.  Device pointers are plain integers (e.g. torch.Tensor.data_ptr())."""

    def __init__(self, width, height, intrinsics=(1.0, 1.0, 0.0, 0.0), tag_size=1.0, **options):
        self.width, self.height = int(width), int(height)
        self.opt = default_options(**options)
        self.encoding = {v: k for k, v in ENCODINGS.items()}[self.opt.input_encoding]
        self.max_batch, self.max_tags = self.opt.max_batch, self.opt.max_tags
        cam = Intrinsics(*[float(v) for v in intrinsics])
        self.h = C.c_void_p()
        rc = lib().b200AprilTagsCreate(C.byref(self.h), self.width, self.height, C.byref(cam), float(tag_size),
                                       C.byref(self.opt))
        if rc != 0:
            self.h = None
            raise B200ATError(rc, "b200AprilTagsCreate")
        self._dets = np.zeros((self.max_batch, self.max_tags), DET_DTYPE)
        self._counts = np.zeros(self.max_batch, np.uint32)

    def close(self):
        if getattr(self, "h", None):
            lib().cuAprilTagsDestroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _frames(self, ptrs, pitch):
        arr = (Frame * len(ptrs))()
        for i, p in enumerate(ptrs):
            arr[i].ptr = int(p)
            arr[i].pitch = int(pitch[i] if hasattr(pitch, "__len__") else pitch)
        return arr

    def frame_table(self, ptrs, pitch):
        """The ctypes frame array of a batch (device addresses + pitch); enqueue() accepts it in place of the address list, so a
        caller that processes the same buffers again and again builds it once."""
        return self._frames(ptrs, pitch)

    def enqueue(self, ptrs, pitch=0, stream=0):
        """Queues one batch on `stream` (b200AprilTagsEnqueueBatch; up to two may be in flight, same stream).  `ptrs`: device
        addresses of the frames' first pixels, or a frame_table()."""
        fr = ptrs if isinstance(ptrs, C.Array) else self._frames(ptrs, pitch)
        rc = lib().b200AprilTagsEnqueueBatch(self.h, fr, len(fr), C.c_void_p(int(stream)))
        if rc != 0:
            raise B200ATError(rc, "b200AprilTagsEnqueueBatch")
        self._dev_q = getattr(self, "_dev_q", []) + [(len(fr), fr)]

    def collect(self, strict=True, copy=True):
        """Waits for the oldest batch queued with enqueue() and returns its detections (list of DET_DTYPE arrays; copy=False:
        views into a buffer that the next collect() overwrites)."""
        n, _keep = self._dev_q.pop(0)
        rc = lib().b200AprilTagsCollectBatch(self.h, self._dets.ctypes.data, None, self._counts.ctypes.data)
        if rc != 0 and (strict or rc != 5):
            raise B200ATError(rc, "b200AprilTagsCollectBatch")
        if copy:
            return [self._dets[i, :self._counts[i]].copy() for i in range(n)]
        return [self._dets[i, :self._counts[i]] for i in range(n)]

    def detect_device(self, ptrs, pitch, stream=0, strict=True):
        """ptrs: device addresses of the frames' first pixels.  Returns a list (per frame) of DET_DTYPE arrays."""
        self.enqueue(ptrs, pitch, stream)
        return self.collect(strict)

    def detect_host(self, frames, strict=True):
        """frames: (n, H, W[, C]) uint8 numpy array in host memory (H2D copies happen inside the call)."""
        if isinstance(frames, (list, tuple)):  # separately allocated frames (each C-contiguous), e.g. one buffer per camera
            frames = [np.ascontiguousarray(f, np.uint8) for f in frames]
            n = len(frames)
            fr = self._frames([f.ctypes.data for f in frames], [f.strides[0] for f in frames])
        else:
            frames = np.ascontiguousarray(frames, np.uint8)
            n = frames.shape[0]
            pitch = frames.strides[1]
            fr = self._frames([frames[i].ctypes.data for i in range(n)], pitch)
        dets = np.zeros((n, self.max_tags), DET_DTYPE)
        counts = np.zeros(n, np.uint32)
        rc = lib().b200AprilTagsDetectBatchHost(self.h, fr, n, dets.ctypes.data, None, counts.ctypes.data)
        if rc != 0 and (strict or rc != 5):
            raise B200ATError(rc, "b200AprilTagsDetectBatchHost")
        return [dets[i, :counts[i]].copy() for i in range(n)]

    def enqueue_host(self, frames):
        """Asynchronous half of detect_host: queues one batch (up to two may be in flight); `frames` must stay alive and
        unmodified until the matching collect_host()."""
        assert frames.dtype == np.uint8 and frames.flags["C_CONTIGUOUS"]
        n = frames.shape[0]
        fr = self._frames([frames[i].ctypes.data for i in range(n)], frames.strides[1])
        rc = lib().b200AprilTagsEnqueueBatchHost(self.h, fr, n)
        if rc != 0:
            raise B200ATError(rc, "b200AprilTagsEnqueueBatchHost")
        self._host_q = getattr(self, "_host_q", []) + [(n, frames)]

    def collect_host(self, strict=True):
        """Waits for the oldest batch queued with enqueue_host and returns its detections (list of DET_DTYPE arrays)."""
        n, _keep = self._host_q.pop(0)
        dets = np.zeros((n, self.max_tags), DET_DTYPE)
        counts = np.zeros(n, np.uint32)
        rc = lib().b200AprilTagsCollectBatchHost(self.h, dets.ctypes.data, None, counts.ctypes.data)
        if rc != 0 and (strict or rc != 5):
            raise B200ATError(rc, "b200AprilTagsCollectBatchHost")
        return [dets[i, :counts[i]].copy() for i in range(n)]

    def set_rectification(self, src_width, src_height, K, D, R, P):
        """Fused rectify / resize pre-stage for the device-pointer entry points (include/b200_apriltags.h); None for K disables."""
        if K is None:
            rc = lib().b200AprilTagsSetRectification(self.h, None)
        else:
            r = Rectify()
            r.struct_size = C.sizeof(Rectify)
            r.src_width, r.src_height = int(src_width), int(src_height)
            r.K[:] = [float(v) for v in np.asarray(K).reshape(-1)]
            d = list(np.asarray(D, float).reshape(-1)) + [0.0] * 8
            r.D[:] = d[:8]
            r.R[:] = [float(v) for v in np.asarray(R).reshape(-1)]
            r.P[:] = [float(v) for v in np.asarray(P).reshape(-1)]
            rc = lib().b200AprilTagsSetRectification(self.h, C.byref(r))
        if rc != 0:
            raise B200ATError(rc, "b200AprilTagsSetRectification")

    def status(self):
        s = C.c_uint32()
        lib().b200AprilTagsLastStatus(self.h, C.byref(s))
        return s.value

    def enable_timing(self, on=True):
        lib().b200AprilTagsEnableStageTiming(self.h, 1 if on else 0)

    def stage_times(self):
        ms = (C.c_float * len(STAGES))()
        lib().b200AprilTagsGetStageTimes(self.h, ms)
        return dict(zip(STAGES, [float(v) for v in ms]))

    def counters(self):
        c = (C.c_uint64 * 8)()
        lib().b200AprilTagsGetCounters(self.h, c)
        return {"launches": c[0], "points": c[1], "clusters": c[2], "quads": c[3], "detections": c[5], "h2d_bytes": c[6],
                "sparse_h2d": c[7]}

    def dims(self):
        v = [C.c_uint32() for _ in range(4)]
        lib().b200AprilTagsGetDims(self.h, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)  # Wd, Hd, tw, th

    def read_buffer(self, which, frame=0):
        wd, hd, tw, th = self.dims()
        n = C.c_size_t()
        lib().b200AprilTagsReadBuffer(self.h, which, frame, None, 0, C.byref(n))
        if which == BUF_RECTIFIED:
            out = np.empty((self.height, self.width), np.uint8)
        elif which in (BUF_DECIMATED, BUF_THRESHOLD):
            out = np.empty((hd, wd), np.uint8)
        elif which in (BUF_TILE_MIN, BUF_TILE_MAX):
            out = np.empty((th, tw), np.uint8)
        elif which in (BUF_LABELS, BUF_SIZES):
            out = np.empty((hd, wd), np.uint32)
        elif which == BUF_CLUSTERS:
            out = np.empty(n.value, CLUSTER_DTYPE)
        elif which == BUF_POINTS:
            out = np.empty(n.value, np.uint64)
        elif which == BUF_POINTS_RAW:
            out = np.empty(n.value, np.uint32)
        else:
            out = np.empty(n.value, QUAD_DTYPE)
        if out.nbytes:
            rc = lib().b200AprilTagsReadBuffer(self.h, which, frame, out.ctypes.data, out.nbytes, C.byref(n))
            if rc != 0:
                raise B200ATError(rc, "b200AprilTagsReadBuffer")
        return out
