import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from isaac_ros_apriltag_b200 import capi, synth
from oracle import oracle as O
import parity_util as pu
frames, truths, K, ts, fams = synth.make_config_frames("C2", 2)
n, H, W = frames.shape[:3]
orc = O.Oracle(fams)
oq_all = []
for i in range(n):
    orc.detect(frames[i]); oq_all.append({int(q["key"]): q["p"].copy() for q in orc.quads(refined=False)})
for rep in range(3):
    det = capi.Detector(W, H, families=fams, encoding="mono8", max_batch=n, max_tags=256)
    t, ptrs, pitch = pu.upload(frames)
    for call in range(3):
        det.detect_device(ptrs, pitch, pu.current_stream(), strict=False)
        gclu = det.read_buffer(capi.BUF_CLUSTERS); gq = det.read_buffer(capi.BUF_QUADS)
        size = {(int(r["frame"]), int(r["key"])): int(r["count"]) for r in gclu}
        bad = []
        for i in range(n):
            g = {int(k): p for k, p in zip(gq[gq["frame"] == i]["key"], gq[gq["frame"] == i]["p"])}
            o = oq_all[i]
            for k in set(g) ^ set(o):
                bad.append((i, "gpu-only" if k in g else "oracle-only", size.get((i, k))))
            for k in set(g) & set(o):
                d = float(np.abs(g[k] - o[k]).max())
                if d > 1e-3: bad.append((i, "corner %.3g" % d, size.get((i, k))))
        print("rep", rep, "call", call, "quads", len(gq), "bad", len(bad), sorted(bad, key=lambda b: b[2] or 0)[:40], flush=True)
    det.close()
