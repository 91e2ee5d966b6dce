"""Python view of the node core (libb200apriltag_node.so): the AprilTagNode plugin surface without ROS.

Mirrors nvidia::isaac_ros::apriltag::AprilTagNode (/root/reference/isaac_ros_apriltag/src/apriltag_node.cpp:562-623):
same parameter names/defaults, same constructor-time validation errors, same per-frame behaviour."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200apriltag_node.so")


class DetectionMsg(C.Structure):
    _fields_ = [("family", C.c_char * 32), ("id", C.c_int32), ("center", C.c_double * 2), ("corners", (C.c_double * 2) * 4),
                ("position", C.c_double * 3), ("orientation_xyzw", C.c_double * 4), ("child_frame_id", C.c_char * 48)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
        L = C.CDLL(LIB_PATH)
        L.b200NodeCreate.argtypes = [C.POINTER(C.c_void_p), C.c_char_p, C.c_char_p, C.c_double, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
        L.b200NodeDestroy.argtypes = [C.c_void_p]
        L.b200NodeUsingCuAprilTagImpl.argtypes = [C.c_void_p]
        L.b200NodeOnFrame.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_double),
                                      C.c_uint32, C.c_uint32, C.c_char_p, C.POINTER(DetectionMsg), C.c_int, C.POINTER(C.c_int),
                                      C.c_char_p, C.c_size_t]
        L.b200NodeRotationToQuaternion.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.b200NodeParseBackends.argtypes = [C.c_char_p]
        L.b200NodeParseBackends.restype = C.c_uint32
        _lib = L
    return _lib


class AprilTagNode:
    """AprilTagNode(tag_family='tag36h11', backends='CUDA', size=0.22, max_tags=64, tile_size=4).
    Raises RuntimeError from the constructor exactly where the reference throws std::runtime_error."""

    def __init__(self, tag_family="tag36h11", backends="CUDA", size=0.22, max_tags=64, tile_size=4):
        self.h = C.c_void_p()
        err = C.create_string_buffer(1024)
        rc = lib().b200NodeCreate(C.byref(self.h), tag_family.encode(), backends.encode(), float(size), int(max_tags), int(tile_size),
                                  err, len(err))
        if rc != 0:
            self.h = None
            raise RuntimeError(err.value.decode())
        self.max_tags = max_tags

    def using_cuapriltag_impl(self):
        return bool(lib().b200NodeUsingCuAprilTagImpl(self.h))

    def on_frame(self, encoding, width, height, step, dev_ptr, K, ci_width=None, ci_height=None, frame_id="camera"):
        out = (DetectionMsg * self.max_tags)()
        n = C.c_int()
        err = C.create_string_buffer(1024)
        K9 = (C.c_double * 9)(*[float(v) for v in np.asarray(K).reshape(-1)])
        rc = lib().b200NodeOnFrame(self.h, encoding.encode(), width, height, step, C.c_void_p(int(dev_ptr)), K9,
                                   ci_width or width, ci_height or height, frame_id.encode(), out, self.max_tags, C.byref(n), err, len(err))
        if rc != 0:
            raise RuntimeError(err.value.decode())
        res = []
        for i in range(n.value):
            d = out[i]
            res.append({"family": d.family.decode(), "id": d.id, "center": np.array(d.center[:]),
                        "corners": np.array([list(r) for r in d.corners]), "position": np.array(d.position[:]),
                        "orientation_xyzw": np.array(d.orientation_xyzw[:]), "child_frame_id": d.child_frame_id.decode()})
        return res

    def close(self):
        if getattr(self, "h", None):
            lib().b200NodeDestroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
