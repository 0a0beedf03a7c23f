#!/usr/bin/env python
"""Human-readable digest of one bench.py JSON line.  usage: python tools/bench_summary.py bench.json"""
import json
import signal
import sys

signal.signal(signal.SIGPIPE, signal.SIG_DFL)      # `| head` must not end in a traceback
d = json.load(open(sys.argv[1]))
print({k: d[k] for k in ("value", "frame_pairs_per_sec", "ms_per_step", "gpu_launches")})
print("e2e", d["e2e"]["frame_pairs_per_sec"], "pairs/s", d["e2e"]["ms_per_step"], "ms/step")
print("roofline", d["roofline"])
for k, v in d["kernels"].items():
    print("  %-16s %8.4f ms x %4.1f /step  %s GB/s" % (k, v["ms_per_launch"], v["launches_per_step"],
                                                     None if v["gbps"] is None else round(v["gbps"])))
print(d["pipeline"])
for key in ("api_single_pair", "select", "sequence_api", "sequence_api_affine", "cpu_baseline", "clocks"):
    if d.get(key) is not None:
        v = d[key]
        print(key, {k: x for k, x in v.items() if k not in ("call", "kernel_ms_per_frame", "sample")} if isinstance(v, dict) else v)
