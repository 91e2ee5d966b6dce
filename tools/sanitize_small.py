"""tiny end-to-end run for compute-sanitizer (memcheck / racecheck): 2 frames 640x480 bgr8 + 1 frame odd size mono8"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from isaac_ros_apriltag_b200 import capi, synth
rng = np.random.default_rng(3)
frames = np.stack([synth.make_frame(rng, 640, 480, [("tag36h11", 5), ("tag36h11", 9)], side_px=(60, 140))[0] for _ in range(2)])
bgr = np.ascontiguousarray(np.repeat(frames[:, :, :, None], 3, axis=3))
det = capi.Detector(640, 480, encoding="bgr8", max_batch=2, max_tags=64)
t = torch.from_numpy(bgr).cuda()
fb = t[0].numel()
r = det.detect_device([t.data_ptr(), t.data_ptr() + fb], 640 * 3, 0)
print("bgr8 ids", [list(x["id"]) for x in r], "status", det.status())
det.close()
g, _ = synth.make_frame(rng, 751, 481, [("tag36h11", 17)], side_px=(60, 120))
det = capi.Detector(751, 481, encoding="mono8", max_batch=1, max_tags=64, quad_sigma=0.8)
t2 = torch.from_numpy(g).cuda()
r = det.detect_device([t2.data_ptr()], 751, 0)
print("mono8 odd ids", [list(x["id"]) for x in r], "status", det.status())
det.close()
