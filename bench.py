#!/usr/bin/env python3
"""bench.py -- headline benchmark: 1080p tag36h11 frames/s on B200 (BASELINE.json metric), contract per the
build instructions.

One "step" = one pass of the full detection hot path (gray+decimate, threshold, union-find, gradient
clusters, quad fit, refine+decode, reconcile, pose) over one batch of synthetic frames.
  value    : whole-job frames/s with the batch already resident in HBM (CUDA events on the launch stream).
  e2e      : same metric through the reference-facing C ABI with HOST (pinned) buffers
             (b200AprilTagsDetectBatchHost): H2D of every frame and D2H of the results inside the timed region.
  roofline : threshold kernel (the metric's named kernel): algorithmic bytes 2*Pd per frame / live event time.
  cpu_baseline : the CPU oracle (port of AprilRobotics apriltag) timed on this box's host cores, rank 0, N=1.

`--impl reference` times the CPU oracle port on the same workload (the reference's closed cuAprilTags / VPI
libraries and the AprilRobotics sources are absent from this image).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames_per_sec_1080p_tag36h11"
UNIT = "frames/s"


NCU_TRAFFIC_FILE = "r04_ncu_full_dense_batch256.csv"  # ncu --set full of the same command and batch (tools/gpu_final.sh)


def load_ncu_traffic(kernel="k_threshold4"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture
    (profiles/, same command and batch as this bench); None when the file is missing."""
    import csv
    path = os.path.join(ROOT, "profiles", NCU_TRAFFIC_FILE)
    try:
        rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("#"))]
        hdr, units = rows[0], rows[1]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            if kernel in r[0]:
                tot = 0.0
                for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    i = hdr.index(name)
                    tot += float(r[i]) * scale[units[i]]
                return tot
    except Exception:
        return None
    return None


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa(gpu_index):
    """Pin this rank's threads to the CPUs next to its GPU's PCIe root BEFORE the pinned staging memory is allocated (first touch
    then places it on that NUMA node): with 8 ranks on one box the host->device path is what bounds `e2e` (SCALE_r01: 0.59 at N=8
    with every rank on the same CPUs).  Returns what was done, for the JSON line."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
        bdf = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(gpu_index)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bdf.startswith("00000000:"):
            bdf = bdf[4:]
        base = f"/sys/bus/pci/devices/{bdf}"
        node = open(base + "/numa_node").read().strip()
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        local = sorted(cpus & set(allowed))
        if local and len(local) < len(allowed):
            os.sched_setaffinity(0, local)
            return {"numa_node": node, "cpus": f"{local[0]}-{local[-1]} ({len(local)})", "bound": True}
        return {"numa_node": node, "cpus": f"{allowed[0]}-{allowed[-1]} ({len(allowed)})", "bound": False,
                "why": "the GPU's local CPUs are all the CPUs this process may use"}
    except Exception as e:  # no sysfs / nvidia-smi: run unbound
        return {"bound": False, "why": repr(e)[:80]}


def make_workload(config, distinct, encoding):
    from isaac_ros_apriltag_b200 import synth
    frames, truths, K, tagsize, fams = synth.make_config_frames(config, distinct)
    if encoding != "mono8":
        ch = 3 if encoding in ("rgb8", "bgr8") else 4
        col = np.repeat(frames[:, :, :, None], ch, axis=3)
        if ch == 4:
            col[..., 3] = 255
        frames = np.ascontiguousarray(col)
    return frames, truths, K, tagsize, fams


def run_reference(args):
    """CPU arm: the oracle port with all host threads, on the same workload, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    frames, _, _, _, fams = make_workload(args.config, args.distinct, args.encoding)
    n_step = args.batch  # the same step as the GPU arm: one batch of the workload (256 frames of C2 take the 16-core port ~1 s)
    idx = np.arange(n_step) % frames.shape[0]
    sample = np.ascontiguousarray(frames[idx])
    for _ in range(args.warmup):
        O.detect_batch(sample[:max(cores, 1)], fams, nthreads=cores, encoding=args.encoding)
    t0 = time.perf_counter()
    ndet = 0
    for _ in range(args.steps):
        res, _times = O.detect_batch(sample, fams, nthreads=cores, encoding=args.encoding)
        ndet += sum(len(r) for r in res)
    dt = time.perf_counter() - t0
    fps = n_step * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(args), "l2": "n/a (CPU arm)", "frames_per_step": n_step,
                       "detections_per_step": ndet // max(args.steps, 1)},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n_step} frames/step x {args.steps} steps, frame-parallel, one detector per thread"},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def workload_name(args):
    desc = {"C2": "1920x1080 tag36h11, 10 tags/frame", "C1": "1280x720, 1 tag36h11", "C3": "3840x2160 tag36h11, 4 tags/frame",
            "C4": "1280x720 dense grid 91 tag36h11", "C5": "1920x1080 6 tag36h11 + 4 tag25h9"}[args.config]
    return f"{args.config}: {desc}, batch={args.batch}/GPU, {args.encoding}, {args.distinct} distinct seeded frames tiled"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--distinct", type=int, default=32)
    ap.add_argument("--encoding", default="bgr8")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from isaac_ros_apriltag_b200 import capi, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the detector has no CPU fallback")
    torch.cuda.set_device(local_rank)
    affinity = bind_to_gpu_numa(local_rank)
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    frames, truths, K, tagsize, fams = make_workload(args.config, args.distinct, args.encoding)
    H, W = frames.shape[1:3]
    B = args.batch
    # device-resident batch: `distinct` seeded frames tiled to the batch size (inputs >> L2: no flush needed)
    dev_distinct = torch.from_numpy(frames).cuda()
    reps = (B + args.distinct - 1) // args.distinct
    dev_batch = dev_distinct.repeat((reps,) + (1,) * (dev_distinct.dim() - 1))[:B].contiguous()
    frame_bytes = dev_batch[0].numel()
    pitch = frame_bytes // H
    ptrs = [dev_batch.data_ptr() + i * frame_bytes for i in range(B)]
    max_tags = 128 if args.config == "C4" else 64  # the dense-grid config has 91 tags per frame
    det = capi.Detector(W, H, intrinsics=(K[0, 0], K[1, 1], K[0, 2], K[1, 2]), tag_size=tagsize, families=fams,
                        encoding=args.encoding, max_batch=B, max_tags=max_tags, device=local_rank)
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream

    # ---- device-resident throughput ----
    for _ in range(args.warmup):
        dets = det.detect_device(ptrs, pitch, sh)
    n_det = sum(len(d) for d in dets)
    status = det.status()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    # two batches in flight (b200AprilTagsEnqueueBatch / CollectBatch): the host side of step k + 1 -- frame table, graph launch,
    # unpacking the results of step k -- overlaps the kernels of step k; every step's results are copied out and unpacked
    ftab = det.frame_table(ptrs, pitch)
    ev0.record(stream)
    launches = 0
    det.enqueue(ftab, stream=sh)
    for _ in range(args.steps - 1):
        det.enqueue(ftab, stream=sh)
        det.collect(copy=False)
        launches += det.counters()["launches"]
    ev1.record(stream)
    det.collect(copy=False)
    launches += det.counters()["launches"]
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    # the sharding module's rule (tests/test_sharding_gloo.py runs the same two functions at world size 2 on gloo): the slowest
    # rank's device time defines the step, `value` = frames all ranks processed / that time
    ms_local = ev0.elapsed_time(ev1)
    ms_total = sharding.max_over_ranks(ms_local, "cuda")
    ms_per_step = ms_total / args.steps
    value = sharding.whole_job_throughput(B * args.steps, ms_local / 1e3, "cuda")
    counters = det.counters()
    # per-stage device times (CUDA events between the stages): measured on extra, UNPIPELINED steps outside the timed
    # region -- with stage timing on, the library runs the batch as one chunk so the stages do not overlap
    det.enable_timing(True)
    stage_acc = {}
    n_stage_steps = 3
    for _ in range(n_stage_steps):
        det.detect_device(ptrs, pitch, sh)
        for k, v in det.stage_times().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    det.enable_timing(False)
    torch.cuda.synchronize()

    # ---- end to end through the C ABI with host buffers ----
    # The host entry point stages the pinned frames in sub-batches (copy of k+1 overlaps the kernels of k).  Two staging
    # modes exist: "full_copy" moves every byte of every frame; "sparse" DMAs only the rows quad detection reads (every
    # quad_decimate-th) and fetches the full-resolution rows around the fitted quads on demand from the pinned frames.
    # `e2e` is the library default (B200AT_SPARSE_H2D unset); both modes are reported, with the bytes they actually moved.
    e2e = None
    e2e_modes = {}
    if not args.no_e2e:
        host = torch.from_numpy(frames).pin_memory()
        host_batch = host.repeat((reps,) + (1,) * (host.dim() - 1))[:B].contiguous().pin_memory()
        hb = host_batch.numpy()
        e2e_steps = max(2, args.steps)
        d2h = B * max_tags * capi.DET_DTYPE.itemsize + B * 4 + 4 * 4 * 32 * 4  # detections, counts, four sub-batches x four counter blocks

        def measure_e2e(mode, pipelined):
            """Every step copies every frame of the batch host->device and every result device->host inside the timed region.
            pipelined: b200AprilTagsEnqueueBatchHost / CollectBatchHost with two batches in flight (step k+1's DMA and quad
            detection overlap step k's decode / pose / D2H); otherwise one synchronous b200AprilTagsDetectBatchHost per step."""
            if mode is None:
                os.environ.pop("B200AT_SPARSE_H2D", None)
            else:
                os.environ["B200AT_SPARSE_H2D"] = "1" if mode == "sparse" else "0"
            for _ in range(max(1, args.warmup - 1)):
                det.detect_host(hb)
            barrier()
            t0 = time.perf_counter()
            if pipelined:
                det.enqueue_host(hb)
                for _ in range(e2e_steps - 1):
                    det.enqueue_host(hb)
                    r = det.collect_host()
                r = det.collect_host()
            else:
                for _ in range(e2e_steps):
                    r = det.detect_host(hb)
            torch.cuda.synchronize()
            dt = sharding.max_over_ranks(time.perf_counter() - t0, "cuda")
            c = det.counters()
            os.environ.pop("B200AT_SPARSE_H2D", None)
            return {"value": world * B * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(c["h2d_bytes"]),
                    "input_bytes_per_step": int(B * frame_bytes), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "staging": "sparse" if c["sparse_h2d"] else "full_copy", "detections_per_step": int(sum(len(x) for x in r)),
                    "calls_in_flight": 2 if pipelined else 1,
                    "h2d_gbs": world * c["h2d_bytes"] * e2e_steps / dt / 1e9,
                    "timer": "host wall clock from the first enqueue to the last collect" if pipelined else "host wall clock around the synchronous C-ABI calls"}

        def h2d_ceiling():
            """What the box gives: plain cudaMemcpyAsync of the pinned batch, all ranks at the same time (aggregate GB/s)."""
            dst = torch.empty_like(dev_batch)
            cs = torch.cuda.Stream()
            with torch.cuda.stream(cs):
                dst.copy_(host_batch, non_blocking=True)
            cs.synchronize()
            barrier()
            t0 = time.perf_counter()
            with torch.cuda.stream(cs):
                for _ in range(3):
                    dst.copy_(host_batch, non_blocking=True)
            cs.synchronize()
            dt = sharding.max_over_ranks(time.perf_counter() - t0, "cuda")
            del dst
            return world * 3 * host_batch.numel() / dt / 1e9

        e2e = measure_e2e(None, True)
        e2e["h2d_ceiling_gbs"] = h2d_ceiling()
        e2e["h2d_ceiling_note"] = "plain pinned cudaMemcpyAsync of the same batch, all ranks concurrently (aggregate)"
        e2e["affinity"] = affinity
        e2e_modes[e2e["staging"] + "_pipelined"] = e2e
        e2e_modes[e2e["staging"] + "_synchronous"] = measure_e2e(None, False)
        other = "full_copy" if e2e["staging"] == "sparse" else "sparse"
        alt = measure_e2e(other, True)
        if alt["staging"] == other:
            e2e_modes[other + "_pipelined"] = alt

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the threshold kernel (metric's named kernel) + per-stage table ----
    peak, peak_kind = load_peaks()
    wd, hd, tw, th = det.dims()
    Pd = wd * hd
    bpp = capi.BPP[args.encoding]
    stage_ms = {k: v / n_stage_steps for k, v in stage_acc.items()}
    alg = {"preprocess": (bpp + 1) * Pd, "threshold": 2 * Pd, "ccl": 5 * Pd, "cluster": 2 * 5 * Pd}
    stage_gbs = {k: (alg[k] * B / (stage_ms[k] / 1e3) / 1e9) if stage_ms.get(k, 0) > 0 else None for k in alg}
    thr_gbs = stage_gbs["threshold"]
    roofline = {"kernel": "k_threshold4", "bound": "hbm", "achieved": thr_gbs, "peak": peak, "unit": "GB/s",
                "frac": (thr_gbs / peak) if thr_gbs else None,
                "traffic": load_ncu_traffic() if (args.config == "C2" and B == 256 and args.encoding == "bgr8") else None,
                "traffic_source": "profiles/" + NCU_TRAFFIC_FILE + " (ncu --set full of this command, bytes per launch; null if not captured)",
                "peak_source": peak_kind,
                "algorithmic_bytes_per_launch": 2 * Pd * B, "launch_ms": stage_ms.get("threshold")}
    dominant = max(stage_ms, key=lambda k: stage_ms[k]) if stage_ms else None
    # SURVEY 8d end-to-end figure: (input bytes + 14 Pd) per frame at the measured frame rate against the HBM peak -- what the whole
    # pipeline achieves of a dense, staged, HBM-bound pipeline; and the same for the slowest stage (quad fit: 4 B packed point in,
    # 8 B sorted point out and in again = 20 B per boundary point kept; the other stages as in stages_gbs)
    in_bytes = frame_bytes
    per_frame = in_bytes + 14 * Pd
    roofline_pipeline = {"bound": "hbm", "bytes_per_frame": int(per_frame), "achieved": per_frame * value / world / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": per_frame * value / world / 1e9 / peak,
                         "note": "whole pipeline: (input bytes + 14 Pd) x frames/s per GPU (SURVEY 8d); the irregular stages are latency / issue bound, see DESIGN.md"}
    alg_dom = dict(alg)
    alg_dom["quadfit"] = 20 * counters["points"] / max(B, 1)
    roofline_dominant = None
    if dominant in alg_dom and stage_ms.get(dominant, 0) > 0:
        gbs = alg_dom[dominant] * B / (stage_ms[dominant] / 1e3) / 1e9
        roofline_dominant = {"stage": dominant, "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                             "algorithmic_bytes_per_frame": int(alg_dom[dominant]), "ms_per_step": stage_ms[dominant],
                             "share_of_step": stage_ms[dominant] / max(sum(v for k, v in stage_ms.items() if k != "d2h"), 1e-9)}

    # ---- single-frame latency through the drop-in entry point (how the reference's README numbers were taken:
    # 720p, one frame in flight; /root/reference/README.md:69 quotes 2.0-2.9 ms for cuAprilTags on other hardware) ----
    latency = None
    if args.gpus == 1 and not args.no_e2e:
        import ctypes as C
        from isaac_ros_apriltag_b200 import synth
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
        lat = []
        lat_stream = torch.cuda.Stream()  # a real (non-legacy) stream, like the reference node's stream_ (apriltag_node.cpp:460):
        lsh = lat_stream.cuda_stream      # lets the library replay its cached CUDA graph of the whole frame
        for i in range(60):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rc = L.cuAprilTagsDetect(hdl, C.byref(img), tags, C.byref(ntags), 64, C.c_void_p(lsh))
            lat.append((time.perf_counter() - t0) * 1e3)
        L.cuAprilTagsDestroy(hdl)
        latency = {"workload": "1280x720 bgr8, 1 tag36h11, one frame in flight, cuAprilTagsDetect (synchronous)",
                   "median_ms": float(np.median(lat[10:])), "p90_ms": float(np.percentile(lat[10:], 90)), "tags": int(ntags.value), "rc": int(rc)}

    cpu_baseline = None
    if not args.no_cpu_baseline and args.gpus == 1:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        n_s = min(max(cores, 16), 256)
        idx = np.arange(n_s) % frames.shape[0]
        sample = np.ascontiguousarray(frames[idx])
        O.detect_batch(sample[:cores], fams, nthreads=cores, encoding=args.encoding)
        t0 = time.perf_counter()
        reps_cpu = 0
        while True:
            res, times = O.detect_batch(sample, fams, nthreads=cores, encoding=args.encoding)
            reps_cpu += 1
            if time.perf_counter() - t0 > 8.0 or reps_cpu >= 20:
                break
        dtc = time.perf_counter() - t0
        t1 = time.perf_counter()
        res1, times1 = O.detect_batch(sample[:4], fams, nthreads=1, encoding=args.encoding)
        dt1 = time.perf_counter() - t1
        cpu_baseline = {"value": n_s * reps_cpu / dtc, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{n_s} frames x {reps_cpu} passes, frame-parallel (one single-thread detector per core)",
                        "single_thread_fps": 4 / dt1,
                        "stage_share_single_thread": {k: round(v / max(times1["total"], 1e-9), 3) for k, v in times1.items() if k != "total"}}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "l2": "inputs larger than L2 (batch %.0f MB)" % (B * frame_bytes / 1e6),
                       "detections_per_batch": n_det, "status": status, "points_per_batch": int(counters["points"]),
                       "clusters_per_batch": int(counters["clusters"]), "quads_per_batch": int(counters["quads"])},
            "clocks": clocks, "e2e": e2e,
            "e2e_modes": {k: {kk: v[kk] for kk in ("value", "h2d_bytes_per_step", "h2d_gbs", "calls_in_flight")} for k, v in e2e_modes.items()},
            "gpu_launches": int(launches), "roofline": roofline, "roofline_pipeline": roofline_pipeline, "roofline_dominant": roofline_dominant,
            "stages_ms_per_step": stage_ms, "stages_note": "stage times from extra steps with CUDA events between the stages (outside the timed region)", "stages_gbs": stage_gbs, "dominant_stage": dominant, "latency_720p": latency,
            "cpu_baseline": cpu_baseline}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
