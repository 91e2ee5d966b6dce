#!/usr/bin/env python3
"""Pretty-prints the JSON lines of tools/gpu_tune.py (markdown tables)."""
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip().startswith("{")]
dev = [r for r in rows if r.get("event") == "device"]
host = [r for r in rows if r.get("event") == "host"]
print("| knobs (device-resident, 256 x 1080p bgr8) | ms/step | frames/s | parity | preprocess | threshold | ccl | cluster | quadfit | decode | finalize |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for r in dev:
    if "error" in r:
        print(f"| {r['tune']} | error: {r['error'][:80]} |")
        continue
    s = r["stages_ms"]
    print(f"| {r['tune']} | {r['ms_per_step']:.2f} | {r['fps']:.0f} | {('ok' if r.get('identical', True) else 'ok (ids exact, corners <= %.1e px)' % (r.get('max_corner_diff') or 0)) if r['parity'] and r['status'] == 0 else 'MISMATCH'} | "
          + " | ".join(f"{s.get(k, 0):.3f}" for k in ("preprocess", "threshold", "ccl", "cluster", "quadfit", "decode", "finalize")) + " |")
for r in rows:
    if r.get("event") == "device_subbatched":
        print(f"\ndevice path in calls of {r.get('sub')} frames: {r.get('ms_per_256', r.get('error'))} ms per 256 frames")
print()
print("| knobs (host entry point) | staging | sub-batch | compute/copy streams | pipelined fetch | ms/step | frames/s | H2D MB/frame | parity |")
print("|---|---|---|---|---|---|---|---|---|")
for r in host:
    if "error" in r:
        print(f"| {r['tune']} | {r.get('sparse')} | {r.get('host_sub')} | {r.get('streams')} | error: {r['error'][:80]} |")
        continue
    print(f"| {r.get('tag', '')} {r['tune']} | {'sparse' if r['sparse'] else 'full copy'} | {r['host_sub']} | {r['streams']}/{r.get('copy_streams', 1)} | {r.get('pipe', 0)}{'+ramp' if r.get('ramp', 0) == 1 else ''} | {r['ms_per_step']:.2f} | {r['fps']:.0f} | "
          f"{r['h2d_bytes'] / 256 / 1e6:.2f} | {'ok' if r['parity'] and r['status'] == 0 else 'MISMATCH'} |")
