#!/usr/bin/env python
"""Per-launch DRAM traffic of the pipeline kernels from an `ncu --set full` capture of one bench.py step, in the form bench.py
reads for `roofline.traffic` (profiles/ncu_traffic_rNN.json).
usage: python tools/ncu_traffic.py report.ncu-rep <frames per launch> <features per launch> "<source note>" > profiles/ncu_traffic_r02.json"""
import csv
import json
import subprocess
import sys
from collections import defaultdict

rep, frames, feats, note = sys.argv[1], float(sys.argv[2]), float(sys.argv[3]), sys.argv[4]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
hdr, units = rows[0], rows[1]
iK, iR, iW = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = defaultdict(list)
for r in rows[2:]:
    name = r[iK]
    for key in ("stream_level01", "stream_level0_", "stream_smooth0", "stream_down2", "stream_grad", "lk_windowed", "lk_track_rows"):
        if key in name:
            acc[key.rstrip("_")].append(float(r[iR]) * scale[units[iR]] + float(r[iW]) * scale[units[iW]])
            break
per_frame = {k: round(sum(v) / len(v) / frames) for k, v in acc.items() if k.startswith("stream_")}
per_feat = {k: round(sum(v) / len(v) / feats) for k, v in acc.items() if k.startswith("lk_")}
json.dump({"source": note, "source_windowed": note, "per_frame_bytes": per_frame, "per_feature_bytes": per_feat,
           "launches": {k: len(v) for k, v in acc.items()},
           "note": "per-launch averages over the captured launches (stream_down2 runs once per level: its entry is the mean of the "
                   "level-1 and level-2 launches, like the algorithmic bytes bench.py divides by)"}, sys.stdout, indent=1)
print()
