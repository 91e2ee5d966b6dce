#!/usr/bin/env python3
"""Single-frame 720p latency through the drop-in entry point (cuAprilTagsDetect on a real stream: CUDA-graph replay), the way
bench.py's latency_720p measures it.  Plain run: prints median / p90.  Under `ncu --metrics gpu__time_duration.sum` it yields the
launch list of one frame (tools/launch_list.py)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from isaac_ros_apriltag_b200 import capi, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
f720, _, K7, ts7, _ = synth.make_config_frames("C1", 1)
bgr = np.ascontiguousarray(np.repeat(f720[0][:, :, None], 3, axis=2))
t720 = torch.from_numpy(bgr).cuda()
L = capi.lib()
hdl = C.c_void_p()
cam = capi.Intrinsics(float(K7[0, 0]), float(K7[1, 1]), float(K7[0, 2]), float(K7[1, 2]))
assert L.nvCreateAprilTagsDetector(C.byref(hdl), 1280, 720, 4, 0, C.byref(cam), C.c_float(ts7)) == 0
img = capi.ImageInput(t720.data_ptr(), 1280 * 3, 1280, 720)
tags = (capi.TagID * 64)()
ntags = C.c_uint32()
st = torch.cuda.Stream()
lat = []
for i in range(n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rc = L.cuAprilTagsDetect(hdl, C.byref(img), tags, C.byref(ntags), 64, C.c_void_p(st.cuda_stream))
    lat.append((time.perf_counter() - t0) * 1e3)
L.cuAprilTagsDestroy(hdl)
k = min(10, n // 2)
print(f"latency_720p median {np.median(lat[k:]):.4f} ms p90 {np.percentile(lat[k:], 90):.4f} ms tags {ntags.value} rc {rc}")
