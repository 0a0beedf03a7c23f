import json, sys
d = json.load(open(sys.argv[1]))
print({k: d[k] for k in ("value", "frame_pairs_per_sec", "ms_per_step", "gpu_launches")})
print("e2e", d["e2e"]["frame_pairs_per_sec"], "pairs/s", d["e2e"]["ms_per_step"], "ms/step")
print("roofline", d["roofline"])
for k, v in d["kernels"].items():
    print("  %-16s %8.4f ms x %4.1f /step  %s GB/s" % (k, v["ms_per_launch"], v["launches_per_step"], None if v["gbps"] is None else round(v["gbps"])))
print(d["pipeline"])
print(d.get("api_single_pair")); print(d.get("cpu_baseline")); print(d.get("clocks"))
