"""The C-ABI library loads without a GPU and exports every symbol include/b200_apriltags.h declares."""
import ctypes as C
import os
import re

import pytest

from isaac_ros_apriltag_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "b200_apriltags.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b((?:nv|cu|b200)[A-Za-z0-9_]+)\s*\(", src)
    return sorted(set(n for n in names if not n.endswith("_t")))


def test_exports_match_header():
    L = capi.lib()
    fns = header_functions()
    assert {"nvCreateAprilTagsDetector", "cuAprilTagsDetect", "cuAprilTagsDestroy"} <= set(fns)
    for name in fns:
        assert hasattr(L, name), name
    assert set(capi.EXPORTED_SYMBOLS) == set(fns)
    assert b"sm_100a" in L.b200AprilTagsVersion()


def test_struct_layouts():
    assert C.sizeof(capi.TagID) == capi.ID_DTYPE.itemsize == 88 and capi.TagID.id.offset == 32 and capi.TagID.orientation.offset == 36
    assert C.sizeof(capi.ImageInput) == 24 and C.sizeof(capi.Intrinsics) == 16
    assert C.sizeof(capi.Detection) == capi.DET_DTYPE.itemsize == 272
    o = capi.default_options()
    assert o.struct_size == C.sizeof(capi.Options)
    assert (o.max_tags, o.tile_size, o.quad_decimate, o.refine_edges, o.max_hamming) == (64, 4, 2.0, 1, 2)
    assert o.input_encoding == capi.ENCODINGS["bgr8"] and o.family_mask == 1


def test_argument_validation_and_no_cpu_fallback():
    L = capi.lib()
    h = C.c_void_p()
    o = capi.default_options()
    cam = capi.Intrinsics(1, 1, 0, 0)
    assert L.b200AprilTagsCreate(C.byref(h), 0, 480, C.byref(cam), 1.0, C.byref(o)) == 1  # INVALID_ARG
    o.quad_decimate = 2.5
    assert L.b200AprilTagsCreate(C.byref(h), 640, 480, C.byref(cam), 1.0, C.byref(o)) == 2  # UNSUPPORTED (only integers and 1.5)
    o = capi.default_options()
    o.family_mask = 1 << 7
    assert L.b200AprilTagsCreate(C.byref(h), 640, 480, C.byref(cam), 1.0, C.byref(o)) == 2
    assert L.nvCreateAprilTagsDetector(C.byref(h), 640, 480, 4, 5, C.byref(cam), 1.0) == 2  # only NVAT_TAG36H11
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        with pytest.raises(capi.B200ATError) as e:
            capi.Detector(640, 480)
        assert e.value.code == 6  # NO_DEVICE: the product path fails loudly, it never falls back to a CPU detector
