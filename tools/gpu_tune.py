#!/usr/bin/env python3
"""One-process sweep of the per-handle performance knobs (B200AT_TUNE, csrc/detector.h struct Tune) and of the host-path
staging modes on the bench workload (C2: 256 x 1080p bgr8).  For every configuration: results compared byte for byte with the
default configuration's (the knobs must never change results), whole-batch time from CUDA events, per-stage times.
Prints one JSON line per configuration; run on the GPU box:  python tools/gpu_tune.py > gpurun_out/tune.jsonl"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from isaac_ros_apriltag_b200 import capi, synth  # noqa: E402

ALL_OFF = "thr_early=0,ccl_sweep=0,cluster_eager=0,decode_split=0,qf_mc=0,qf_keys23=0"   # the round-1 kernels
ALL_ON = ""                                                                               # library defaults
# what the next GPU session should measure first (all bit-exact under the emulator, none measured yet except decode_pair)
DEVICE_CONFIGS = ["qf_exact=1", "", "ccl_tma=0", "qf_bucket_limit=3", ""]
if os.environ.get("B200AT_TUNE_CONFIGS"):
    DEVICE_CONFIGS = os.environ["B200AT_TUNE_CONFIGS"].split(";")
# host entry point: (knobs, sparse staging (-1 = library default), sub-batch (0 = default), streams, pipelined fetch level, ramp, copy streams)
HOST_CONFIGS = [("", -1, 0, 1, -1, -1, 1), ("", 1, 64, 1, 2, 1, 1), ("", 1, 48, 1, 2, 1, 1), ("", 1, 32, 1, 2, 1, 1)]


def emit(**kw):
    print(json.dumps(kw), flush=True)


def same(a, b):
    """ids / Hamming / family / order identical on every frame; returns (ok, bytes_identical, max corner difference in px)"""
    if len(a) != len(b):
        return False, False, None
    ident, mx = True, 0.0
    for x, y in zip(a, b):
        if len(x) != len(y) or not (np.array_equal(x["id"], y["id"]) and np.array_equal(x["hamming"], y["hamming"]) and np.array_equal(x["family"], y["family"])):
            return False, False, None
        ident = ident and x.tobytes() == y.tobytes()
        if len(x):
            mx = max(mx, float(np.abs(x["p"] - y["p"]).max()))
    return True, ident, mx


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--distinct", type=int, default=32)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--host-only", action="store_true")
    ap.add_argument("--device-only", action="store_true")
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    t_start = time.time()
    frames, truths, K, tagsize, fams = synth.make_config_frames("C2", args.distinct)
    frames = np.ascontiguousarray(np.repeat(frames[:, :, :, None], 3, axis=3))
    H, W = frames.shape[1:3]
    B = args.batch
    reps = (B + args.distinct - 1) // args.distinct
    dev = torch.from_numpy(frames).cuda().repeat((reps, 1, 1, 1))[:B].contiguous()
    fb = dev[0].numel()
    ptrs = [dev.data_ptr() + i * fb for i in range(B)]
    pitch = W * 3
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream
    emit(event="setup", seconds=time.time() - t_start, gpu=torch.cuda.get_device_name(0))

    def make(tune):
        if tune:
            os.environ["B200AT_TUNE"] = tune
        else:
            os.environ.pop("B200AT_TUNE", None)
        return capi.Detector(W, H, intrinsics=(K[0, 0], K[1, 1], K[0, 2], K[1, 2]), tag_size=tagsize, families=fams, encoding="bgr8",
                             max_batch=B, max_tags=64)

    base = None
    configs = DEVICE_CONFIGS[:4] + [ALL_ON] if args.quick else DEVICE_CONFIGS
    if args.host_only:
        det = make("")
        base = det.detect_device(ptrs, pitch, sh)
        det.close()
        configs = []
    for tune in configs:
        try:
            det = make(tune)
            for _ in range(2):
                dets = det.detect_device(ptrs, pitch, sh)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            ev0.record(stream)
            for _ in range(args.steps):
                det.detect_device(ptrs, pitch, sh)
            ev1.record(stream)
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / args.steps
            det.enable_timing(True)
            acc = {}
            for _ in range(2):
                det.detect_device(ptrs, pitch, sh)
                for k, v in det.stage_times().items():
                    acc[k] = acc.get(k, 0.0) + v / 2
            det.enable_timing(False)
            if base is None:
                base = dets
            ok, ident, mx = same(dets, base)
            emit(event="device", tune=tune or "default", ms_per_step=ms, fps=B / ms * 1e3, parity=ok, identical=ident, max_corner_diff=mx, status=det.status(),
                 detections=int(sum(len(d) for d in dets)), stages_ms={k: round(v, 4) for k, v in acc.items()})
            det.close()
        except Exception as e:  # keep sweeping: one bad configuration must not cost the others
            emit(event="device", tune=tune or "default", error=repr(e))
    # ---- device path in sub-batch sized calls (what the host path's sub-batching costs by itself) ----
    for sub in ():
        try:
            os.environ.pop("B200AT_TUNE", None)
            det = capi.Detector(W, H, intrinsics=(K[0, 0], K[1, 1], K[0, 2], K[1, 2]), tag_size=tagsize, families=fams, encoding="bgr8",
                                max_batch=sub, max_tags=64)
            for i0 in range(0, B, sub):
                det.detect_device(ptrs[i0:i0 + sub], pitch, sh)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            ev0.record(stream)
            for _ in range(2):
                for i0 in range(0, B, sub):
                    det.detect_device(ptrs[i0:i0 + sub], pitch, sh)
            ev1.record(stream)
            torch.cuda.synchronize()
            emit(event="device_subbatched", sub=sub, ms_per_256=ev0.elapsed_time(ev1) / 2)
            det.close()
        except Exception as e:
            emit(event="device_subbatched", sub=sub, error=repr(e))
    # ---- host path: staging mode x sub-batch size x knobs ----
    if args.device_only:
        emit(event="done", seconds=time.time() - t_start)
        return
    host = torch.from_numpy(frames).pin_memory().repeat((reps, 1, 1, 1))[:B].contiguous().pin_memory().numpy()
    for tune, sparse_on, sub_i, streams_i, pipe_i, ramp_i, ncopy_i in HOST_CONFIGS:
        mode, sub, streams = str(sparse_on), str(sub_i), str(streams_i)
        try:
            for k in ("B200AT_SPARSE_H2D", "B200AT_HOST_SUB", "B200AT_HOST_STREAMS", "B200AT_HOST_PIPE", "B200AT_HOST_RAMP"):
                os.environ.pop(k, None)
            if sparse_on >= 0:
                os.environ["B200AT_SPARSE_H2D"] = mode
            if sub_i > 0:
                os.environ["B200AT_HOST_SUB"] = sub
            os.environ["B200AT_HOST_STREAMS"] = streams
            os.environ["B200AT_HOST_COPY_STREAMS"] = str(ncopy_i)
            if pipe_i >= 0:
                os.environ["B200AT_HOST_PIPE"] = str(pipe_i)
            if ramp_i >= 0:
                os.environ["B200AT_HOST_RAMP"] = str(ramp_i)
            det = make(tune)
            det.detect_host(host)
            t0 = time.perf_counter()
            for _ in range(4):
                r = det.detect_host(host)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 4
            c = det.counters()
            emit(event="host", tag=args.tag, tune=tune or "default", sparse=int(c["sparse_h2d"]), host_sub=int(sub), streams=int(streams), pipe=pipe_i, ramp=ramp_i, copy_streams=ncopy_i,
                 ms_per_step=dt * 1e3, fps=B / dt, h2d_bytes=int(c["h2d_bytes"]), input_bytes=int(host.nbytes), parity=same(r, base)[0], status=det.status())
            det.close()
        except Exception as e:
            emit(event="host", tune=tune or "default", sparse=int(mode), host_sub=int(sub), streams=int(streams), pipe=pipe_i, error=repr(e))
    os.environ.pop("B200AT_HOST_PIPE", None)
    os.environ.pop("B200AT_HOST_RAMP", None)
    os.environ.pop("B200AT_HOST_COPY_STREAMS", None)
    os.environ.pop("B200AT_SPARSE_H2D", None)
    os.environ.pop("B200AT_HOST_SUB", None)
    os.environ.pop("B200AT_HOST_STREAMS", None)
    emit(event="done", seconds=time.time() - t_start)


if __name__ == "__main__":
    main()
