#!/usr/bin/env python3
"""H2D rate of the host path's DMA pattern: every quad_decimate-th row of pinned 1080p bgr8 frames (cudaMemcpy2DAsync with a
doubled source pitch) against one contiguous copy of whole frames -- the ceiling the sparse staging can reach on this box."""
import time
import numpy as np
import torch
from cuda import cudart

B, H, W, C = 256, 1080, 1920, 3
host = torch.empty((B, H, W, C), dtype=torch.uint8).pin_memory()
dev = torch.empty((B, H, W, C), dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()
row = W * C


def timed(fn, nbytes, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


def full():
    cudart.cudaMemcpyAsync(dev.data_ptr(), host.data_ptr(), B * H * row, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, st.cuda_stream)
    st.synchronize()


def rows2(per_call_frames):
    def f():
        for i in range(0, B, per_call_frames):
            n = min(per_call_frames, B - i)
            if per_call_frames == 1:
                cudart.cudaMemcpy2DAsync(dev.data_ptr() + i * H * row, row, host.data_ptr() + i * H * row, 2 * row, row, H // 2,
                                         cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, st.cuda_stream)
            else:  # several frames as one 2-D copy: the frames are contiguous, so rows 0, 2, 4 ... continue across frames (H is even)
                cudart.cudaMemcpy2DAsync(dev.data_ptr() + i * H * row, row, host.data_ptr() + i * H * row, 2 * row, row, n * H // 2,
                                         cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, st.cuda_stream)
        st.synchronize()
    return f


print("contiguous whole frames: %.1f GB/s" % timed(full, B * H * row))
print("every 2nd row, one 2-D copy per frame: %.1f GB/s" % timed(rows2(1), B * (H // 2) * row))
print("every 2nd row, one 2-D copy per 64 frames: %.1f GB/s" % timed(rows2(64), B * (H // 2) * row))
