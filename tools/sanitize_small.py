"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck / synccheck): every default kernel of the device path in both
quad-fit modes and both CCL staging modes, an odd-sized mono8 frame with blur (generic load path, partial tiles), a large cluster
(multi-warp sort bins, multi-chunk windows), the sparse host path and two asynchronous host calls in flight."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from isaac_ros_apriltag_b200 import capi, synth

rng = np.random.default_rng(3)
frames = np.stack([synth.make_frame(rng, 640, 480, [("tag36h11", 5), ("tag36h11", 9)], side_px=(60, 140))[0] for _ in range(4)])
bgr = np.ascontiguousarray(np.repeat(frames[:, :, :, None], 3, axis=3))
t = torch.from_numpy(bgr).cuda()
fb = t[0].numel()
for tune in ("", "qf_exact=1", "ccl_tma=0", "qf_bucket_limit=3"):
    if tune:
        os.environ["B200AT_TUNE"] = tune
    else:
        os.environ.pop("B200AT_TUNE", None)
    det = capi.Detector(640, 480, encoding="bgr8", max_batch=4, max_tags=64)
    r = det.detect_device([t.data_ptr() + i * fb for i in range(4)], 640 * 3, 0)
    print(f"bgr8 [{tune or 'default'}] ids", [list(x["id"]) for x in r], "status", det.status())
    if not tune:
        host = torch.from_numpy(bgr).pin_memory().numpy()
        for mode in ("1", "0"):
            os.environ["B200AT_SPARSE_H2D"] = mode
            os.environ["B200AT_HOST_SUB"] = "1"
            h = det.detect_host(host)
            det.enqueue_host(host)
            det.enqueue_host(host)
            a, b = det.collect_host(), det.collect_host()
            print(f"host sparse={mode} ids", [list(x["id"]) for x in h], "async ok", all(x.tobytes() == y.tobytes() for x, y in zip(a + b, h + h)),
                  "sparse", det.counters()["sparse_h2d"])
        os.environ.pop("B200AT_SPARSE_H2D")
        os.environ.pop("B200AT_HOST_SUB")
    det.close()
os.environ.pop("B200AT_TUNE", None)
# six frames (more than kRefineCtaFrames: the warp-per-quad refine kernel), two device batches in flight on one stream
frames6 = np.stack([synth.make_frame(rng, 640, 480, [("tag36h11", 20 + i)], side_px=(60, 140))[0] for i in range(6)])
t6 = torch.from_numpy(np.ascontiguousarray(np.repeat(frames6[:, :, :, None], 3, axis=3))).cuda()
fb6 = t6[0].numel()
det = capi.Detector(640, 480, encoding="bgr8", max_batch=6, max_tags=64)
st = torch.cuda.Stream()
p6 = [t6.data_ptr() + i * fb6 for i in range(6)]
sync = det.detect_device(p6, 640 * 3, st.cuda_stream)
for _ in range(3):  # plain, capture, replay on both slots
    det.enqueue(p6, 640 * 3, st.cuda_stream)
    det.enqueue(p6[:3], 640 * 3, st.cuda_stream)
    a, b = det.collect(), det.collect()
print("six frames ids", [list(x["id"]) for x in sync], "two in flight ok",
      all(x.tobytes() == y.tobytes() for x, y in zip(a, sync)) and all(x.tobytes() == y.tobytes() for x, y in zip(b, sync[:3])), "status", det.status())
det.close()
g, _ = synth.make_frame(rng, 751, 481, [("tag36h11", 17)], side_px=(60, 120))
det = capi.Detector(751, 481, encoding="mono8", max_batch=1, max_tags=64, quad_sigma=0.8)
t2 = torch.from_numpy(g).cuda()
r = det.detect_device([t2.data_ptr()], 751, 0)
print("mono8 odd ids", [list(x["id"]) for x in r], "status", det.status())
det.close()
big, _ = synth.make_frame(np.random.default_rng(11), 1600, 1200, [("tag36h11", 42)], side_px=(680, 720), max_tilt_deg=10.0, noise_sigma=0.0)
det = capi.Detector(1600, 1200, encoding="mono8", max_batch=1, max_tags=16)
t3 = torch.from_numpy(big).cuda()
r = det.detect_device([t3.data_ptr()], 1600, 0)
print("large cluster ids", [list(x["id"]) for x in r], "points", det.counters()["points"], "status", det.status())
det.close()
