#!/usr/bin/env python3
"""Diagnostic: run stage-by-stage GPU-vs-oracle parity on seeded synthetic frames and print a summary."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

import numpy as np

from isaac_ros_apriltag_b200 import synth
import parity_util as pu


def main():
    cases = sys.argv[1:] or ["C1:2", "C2:2"]
    for c in cases:
        name, n = c.split(":")
        frames, truths, K, ts, fams = synth.make_config_frames(name, int(n))
        rep = []
        t = time.time()
        res, _ = pu.compare_stages(frames, "mono8", fams, report=rep)
        print(name, json.dumps(res), "%.1fs" % (time.time() - t), flush=True)
        for r in rep:
            print("   ", r)


if __name__ == "__main__":
    main()
